"""Our kernels against the REFERENCE'S OWN CUDA kernels on the same GPU, same inputs.

oracle/_ref/alonet_ref_msda.so is the reference extension (alonet/deformable_detr/ops/src) compiled for sm_100a by
oracle/build_ref_cuda.py where /root/reference exists; it travels to the GPU box as a built artefact.  Skipped when absent.
"""
import pytest
import torch

import aloception_oss_b200 as msda
from aloception_oss_b200.synthetic import WORKLOADS, Workload, device_inputs
from oracle import build_ref_cuda
from tests._util import assert_close, near_floor_discontinuity, rms

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not build_ref_cuda.built(), reason="oracle/_ref not built")]

CASES = [
    (WORKLOADS["C2"], "unit"),
    (WORKLOADS["C2"], "wide"),
    (WORKLOADS["C5DEC"], "unit"),
    (Workload("enc_small", 2, ((40, 60), (20, 30), (10, 15), (5, 8)), 3190, M=8, P=4, D=32), "raster"),
    (Workload("d64", 2, ((17, 9), (8, 5)), 50, M=4, P=4, D=64), "wide"),
    (Workload("d30_generic", 1, ((6, 4), (3, 2)), 9, M=2, P=2, D=30), "wide"),   # ops/test.py channel sweep
    (Workload("d71_generic", 1, ((6, 4), (3, 2)), 9, M=2, P=2, D=71), "wide"),
]


@pytest.mark.parametrize("w,mode", CASES, ids=lambda c: c.name if isinstance(c, Workload) else c)
def test_forward_and_backward_match_reference_cuda(w, mode, cuda_device):
    msda.load_ops()
    ref = build_ref_cuda.load_ops()
    x = device_inputs(w, seed=41, device=cuda_device, loc_mode=mode)
    o1 = msda.ms_deform_attn_forward(x["value"], x["shapes"], x["start"], x["loc"], x["attn"])
    o2 = ref.ms_deform_attn_forward(x["value"], x["shapes"], x["start"], x["loc"], x["attn"], 64)
    g1 = msda.ms_deform_attn_backward(x["value"], x["shapes"], x["start"], x["loc"], x["attn"], x["grad_out"])
    g2 = ref.ms_deform_attn_backward(x["value"], x["shapes"], x["start"], x["loc"], x["attn"], x["grad_out"], 64)
    f = lambda t: t.double().cpu().numpy()
    # both are fp32 evaluations with different summation orders (and fp32 atomics on both sides for grad_value)
    assert_close(f(o1), f(o2), 1e-4, 1e-5 * rms(f(o2)), "out")
    assert_close(f(g1[0]), f(g2[0]), 1e-4, 1e-4 * rms(f(g2[0])), "grad_value")
    assert_close(f(g1[2]), f(g2[2]), 1e-4, 1e-4 * rms(f(g2[2])), "grad_attn")
    gl1, gl2 = f(g1[1]), f(g2[1])
    skip = near_floor_discontinuity(f(x["loc"]), x["shapes"].cpu().numpy())
    gl1[skip] = gl2[skip]
    assert_close(gl1, gl2, 1e-4, 1e-4 * rms(gl2), "grad_loc")
