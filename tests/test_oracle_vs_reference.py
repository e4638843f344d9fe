"""Live check of the oracles against the unmodified reference (only where /root/reference exists).

On the GPU box the tree is absent and this module is skipped; the committed fixtures
(tests/test_oracle_golden.py) carry the same comparison there.
"""
import numpy as np
import pytest
import torch

from aloception_oss_b200.synthetic import Workload, torch_inputs
from oracle import msda_oracle, msda_torch_port, ref_loader
from tests._util import assert_close

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present")

CASES = [
    Workload("a", 2, ((6, 4), (3, 2), (5, 7)), 7, M=3, P=4, D=5),
    Workload("b", 1, ((1, 1), (2, 3)), 3, M=1, P=1, D=1),
    Workload("c", 3, ((9, 11),), 17, M=4, P=5, D=8),
]


@pytest.mark.parametrize("w", CASES, ids=lambda w: w.name)
@pytest.mark.parametrize("mode", ["unit", "wide", "local"])
def test_oracles_equal_reference_fp64(w, mode):
    x = torch_inputs(w, seed=11, loc_mode=mode, dtype=torch.float64)
    out, gv, gl, ga = ref_loader.reference_fwd_bwd(x["value"], x["shapes"], x["loc"], x["attn"], x["grad_out"])
    o2 = msda_oracle.forward_t(x["value"], x["shapes"], x["loc"], x["attn"], x["start"])
    g2 = msda_oracle.backward_t(x["grad_out"], x["value"], x["shapes"], x["loc"], x["attn"], x["start"])
    o3, *g3 = msda_torch_port.msda_fwd_bwd_port(x["value"], x["shapes"], x["loc"], x["attn"], x["grad_out"])
    for got in (o2, o3):
        assert_close(got.numpy(), out.numpy(), 1e-10, 1e-15, "out")
    for got in (g2, g3):
        for a, b, n in zip(got, (gv, gl, ga), ("grad_value", "grad_loc", "grad_attn")):
            assert_close(a.numpy(), b.numpy(), 1e-10, 1e-15, n)


def test_exact_boundary_locations():
    """Samples exactly on the skip-window edges and on pixel centres (floor discontinuities)."""
    w = Workload("edge", 1, ((4, 4), (2, 2)), 4, M=1, P=4, D=3)
    x = torch_inputs(w, seed=1, dtype=torch.float64)
    vals = torch.tensor([-0.125, 0.0, 0.125, 0.375, 0.5, 0.875, 1.0, 1.125, 1.25, -0.25], dtype=torch.float64)
    idx = torch.arange(x["loc"].numel()) % vals.numel()
    x["loc"] = vals[idx].view_as(x["loc"]).contiguous()
    out = ref_loader.reference_forward(x["value"], x["shapes"], x["loc"], x["attn"])
    o2 = msda_oracle.forward_t(x["value"], x["shapes"], x["loc"], x["attn"], x["start"])
    assert_close(o2.numpy(), out.numpy(), 1e-12, 1e-16)
