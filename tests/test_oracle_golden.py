"""The CPU oracles (C restatement, torch port) against the golden vectors of the real reference.

tests/golden/*.npz were produced by oracle/make_golden.py from the reference's own
``ms_deform_attn_core_pytorch`` + autograd (ms_deform_attn_func.py:85-190).  This is the test that
PINS the oracle: everything the GPU parity tests compare against is validated here first.
"""
import numpy as np
import pytest
import torch

from oracle import msda_oracle, msda_torch_port
from tests._util import assert_close, assert_close_grad, check_grad_value, golden_names, load_golden

NAMES = golden_names()


def test_fixtures_present():
    assert len(NAMES) >= 10, "golden fixtures missing: run oracle/make_golden.py in the build container"


@pytest.mark.parametrize("name", NAMES)
def test_c_oracle_matches_reference(name):
    w, x, ref = load_golden(name)
    f64 = x["value"].dtype == np.float64
    out = msda_oracle.forward(x["value"], x["shapes"], x["loc"], x["attn"], x["start"])
    gv, gl, ga = msda_oracle.backward(x["grad_out"], x["value"], x["shapes"], x["loc"], x["attn"], x["start"])
    if f64:  # reference op test: fp64 allclose with default tolerances (ops/test.py:60)
        rt, at = 1e-9, 1e-14
        assert_close(out, ref["out"], rt, at, "out")
        assert_close(gl, ref["grad_loc"], rt, at, "grad_loc")
        assert_close(ga, ref["grad_attn"], rt, at, "grad_attn")
        check_grad_value(gv, ref, rt, at)
    else:  # fp32 oracle vs the fp64 evaluation of the same inputs: BASELINE.json's 1e-4 rtol
        assert_close(out, ref["out64"], 1e-4, 1e-8, "out")
        assert_close(out, ref["out"], 1e-4, 1e-8, "out(ref f32)")
        assert_close_grad(gl, ref["grad_loc64"], 1e-4, "grad_loc")
        assert_close_grad(ga, ref["grad_attn64"], 1e-4, "grad_attn")
        check_grad_value(gv, ref, 1e-4)


@pytest.mark.parametrize("name", [n for n in NAMES if not n.startswith("C2")])
def test_torch_port_matches_reference(name):
    w, x, ref = load_golden(name)
    t = {k: torch.from_numpy(v) for k, v in x.items()}
    out, gv, gl, ga = msda_torch_port.msda_fwd_bwd_port(t["value"], t["shapes"], t["loc"], t["attn"], t["grad_out"])
    if x["value"].dtype == np.float64:
        rt, at = 1e-9, 1e-14
        assert_close(out.numpy(), ref["out"], rt, at, "out")
        assert_close(gl.numpy(), ref["grad_loc"], rt, at, "grad_loc")
        assert_close(ga.numpy(), ref["grad_attn"], rt, at, "grad_attn")
        check_grad_value(gv.numpy(), ref, rt, at)
    else:
        assert_close(out.numpy(), ref["out"], 1e-5, 1e-9, "out")
        assert_close_grad(gl.numpy(), ref["grad_loc64"], 1e-4, "grad_loc")
        assert_close_grad(ga.numpy(), ref["grad_attn64"], 1e-4, "grad_attn")
        check_grad_value(gv.numpy(), ref, 1e-4)


def test_port_grid_sample_variant_agrees():
    w, x, ref = load_golden("ragged_wide_f64")
    t = {k: torch.from_numpy(v) for k, v in x.items()}
    a = msda_torch_port.msda_core_port(t["value"], t["shapes"], t["loc"], t["attn"])
    b = msda_torch_port.msda_core_port(t["value"], t["shapes"], t["loc"], t["attn"], use_grid_sample=True)
    assert_close(a.numpy(), b.numpy(), 1e-9, 1e-14)


def test_c_oracle_f64_of_f32_inputs_is_the_headroom_reference():
    """out64 in the fixtures is the reference run in float64 on the float32 draws."""
    w, x, ref = load_golden("small4lvl_wide_f32")
    out = msda_oracle.forward(x["value"].astype(np.float64), x["shapes"], x["loc"].astype(np.float64),
                              x["attn"].astype(np.float64), x["start"])
    assert_close(out, ref["out64"], 2e-7, 1e-10)


def test_oracle_drops_samples_with_wild_locations():
    """NaN / Inf / huge sampling locations fail the reference's window test (ms_deform_im2col_cuda.cuh:285-291) and are
    skipped: same result as the same call with those samples given zero weight on a harmless location."""
    import torch

    from aloception_oss_b200.synthetic import Workload, torch_inputs

    w = Workload("wild", 1, ((5, 6), (3, 3)), 7, M=2, P=3, D=4)
    x = {k: (v.double().numpy() if v.is_floating_point() else v.numpy()) for k, v in torch_inputs(w, seed=5, loc_mode="wide").items()}
    mask = np.zeros(x["loc"].shape[:-1], dtype=bool)
    mask.reshape(-1)[::3] = True
    wild = np.array([np.nan, np.inf, -np.inf, 1e30, -1e30, 3e9, -7.5])
    x["loc"][mask] = wild[np.arange(mask.sum()) % wild.size][:, None]
    y = {k: v.copy() for k, v in x.items()}
    y["attn"][mask] = 0.0
    y["loc"][mask] = 0.5
    got = msda_oracle.forward(x["value"], x["shapes"], x["loc"], x["attn"], x["start"])
    want = msda_oracle.forward(y["value"], y["shapes"], y["loc"], y["attn"], y["start"])
    assert np.isfinite(got).all() and np.array_equal(got, want)
    g = msda_oracle.backward(x["grad_out"], x["value"], x["shapes"], x["loc"], x["attn"], x["start"])
    gw = msda_oracle.backward(y["grad_out"], y["value"], y["shapes"], y["loc"], y["attn"], y["start"])
    assert np.array_equal(g[0], gw[0]) and not g[2][mask].any() and not g[1][mask].any()
