"""GPU parity of the TILE-BINNED backward (aloception_oss_b200/csrc/msda_bwd_tile.cuh; knob ``bwd_tile_mode``) against the C
oracle (restatement of ms_deform_im2col_cuda.cuh:87-159) and against the unit-ordered backward kernel.

The tile kernel is a schedule, not a new function: for ANY sampling locations it must return the reference's gradients
(fp32 tolerance 1e-4 + the rms term of tests/_util, because only the order of the fp32 additions differs).  Cases: raster
queries with local offsets (what it is built for), uniformly random and out-of-range locations (records mostly outside the
window -> "direct" destinations, shrunken windows), Lq != S (clipped tiles, linear tail tiles), one-pixel levels, NaN / Inf
locations, full-size encoder calls.
"""
import numpy as np
import pytest
import torch

import aloception_oss_b200 as msda
from aloception_oss_b200 import _capi
from aloception_oss_b200.synthetic import WORKLOADS, Workload, device_inputs, torch_inputs
from oracle import msda_oracle
from tests._util import assert_close, assert_close_grad, assert_close_grad_loc, check_grad_value, rms

pytestmark = pytest.mark.gpu

KNOBS = ("bwd_tile_mode", "bwd_tile_ctas", "no_pdl", "force_generic")


@pytest.fixture(autouse=True)
def _ops(cuda_device):
    msda.load_ops()
    for k in KNOBS:
        _capi.set_tuning(k, 0)
    yield
    for k in KNOBS:
        _capi.set_tuning(k, 0)


def bwd(x):
    return msda.ms_deform_attn_backward(x["value"], x["shapes"], x["start"], x["loc"], x["attn"], x["grad_out"])


def oracle_bwd(x):
    n = {k: (v.detach().double().cpu().numpy() if v.is_floating_point() else v.cpu().numpy()) for k, v in x.items()}
    return msda_oracle.backward(n["grad_out"], n["value"], n["shapes"], n["loc"], n["attn"], n["start"])


def check_against_oracle(x, got):
    want = oracle_bwd(x)
    gv, gl, ga = (g.double().cpu().numpy() for g in got)
    check_grad_value(gv, {"grad_value": want[0]}, 1e-4)
    assert_close_grad_loc(gl, want[1], x["loc"].cpu().numpy(), x["shapes"].cpu().numpy(), 1e-4)
    assert_close_grad(ga, want[2], 1e-4, "grad_attn")


PYR = ((20, 27), (10, 14), (5, 7), (3, 4))
PYR_S = sum(h * w for h, w in PYR)  # 727
L5 = ((17, 19), (9, 10), (5, 5), (3, 3), (2, 2))
CASES = [
    # name, workload, loc_mode
    ("raster", Workload("t_raster", 2, PYR, PYR_S), "raster"),
    ("unit", Workload("t_unit", 2, PYR, PYR_S), "unit"),
    ("wide", Workload("t_wide", 2, PYR, PYR_S), "wide"),
    ("local", Workload("t_local", 2, PYR, PYR_S), "local"),
    ("few_queries", Workload("t_few", 2, PYR, 37), "wide"),        # Lq < S: clipped tiles
    ("tail", Workload("t_tail", 1, PYR, PYR_S + 300), "wide"),     # Lq > S: linear tail tiles
    ("big_level", Workload("t_big", 1, ((70, 90), (4, 5)), 900, M=3), "unit"),  # window shrinks (70 x 90 > 2048 destinations)
    ("thin", Workload("t_thin", 2, ((1, 40), (33, 1), (1, 1)), 74, M=2), "wide"),  # one-pixel-wide levels
    ("one_level", Workload("t_l1", 1, ((16, 16),), 256, M=8), "raster"),
    ("five_levels", Workload("t_l5", 1, L5, sum(h * w for h, w in L5), M=4), "raster"),
]


def make_inputs(w, mode, dev, seed=7):
    if mode == "raster":
        return device_inputs(w, seed=seed, device=dev, loc_mode="raster")
    return {k: v.to(dev) for k, v in torch_inputs(w, seed=seed, loc_mode=mode).items()}


@pytest.mark.parametrize("name,w,mode", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("no_pdl", [0, 1])
def test_tile_backward_vs_oracle(name, w, mode, no_pdl, cuda_device):
    x = make_inputs(w, mode, cuda_device)
    _capi.set_tuning("no_pdl", no_pdl)
    _capi.set_tuning("bwd_tile_mode", 2)
    n0 = _capi.kernel_launch_count()
    got = bwd(x)
    torch.cuda.synchronize()
    assert _capi.kernel_launch_count() == n0 + 2  # zero-fill + tile kernel
    check_against_oracle(x, got)
    # and the unit-ordered kernel on the same inputs: grad_loc / grad_attn agree to fp32 rounding of the channel sums
    _capi.set_tuning("bwd_tile_mode", 1)
    base = bwd(x)
    for a, b, nm in zip(got, base, ("grad_value", "grad_loc", "grad_attn")):
        a, b = a.double().cpu().numpy(), b.double().cpu().numpy()
        if nm == "grad_loc":
            assert_close_grad_loc(a, b, x["loc"].cpu().numpy(), x["shapes"].cpu().numpy(), 1e-4, nm)
        else:
            assert_close(a, b, 1e-4, 1e-4 * rms(b), nm)


def test_tile_backward_non_finite_locations(cuda_device):
    """NaN / Inf / huge sampling locations are dropped like in the reference: exactly zero gradients for those samples."""
    w = Workload("t_nan", 2, PYR, PYR_S)
    x = device_inputs(w, seed=9, device=cuda_device, loc_mode="raster")
    loc = x["loc"].clone()
    flat = loc.view(-1, 2)
    bad = torch.tensor([float("nan"), float("inf"), -float("inf"), 1e30, -1e30], device=cuda_device)
    idx = torch.arange(0, flat.shape[0], 97, device=cuda_device)
    flat[idx, 0] = bad[idx % 5]
    flat[idx + 1, 1] = bad[(idx + 2) % 5]
    x = dict(x, loc=loc)
    _capi.set_tuning("bwd_tile_mode", 2)
    gv, gl, ga = bwd(x)
    torch.cuda.synchronize()
    assert torch.isfinite(gv).all() and torch.isfinite(gl).all() and torch.isfinite(ga).all()
    glf, gaf = gl.view(-1, 2), ga.view(-1)
    for i in (idx, idx + 1):
        assert (glf[i] == 0).all() and (gaf[i] == 0).all()
    _capi.set_tuning("bwd_tile_mode", 1)
    base = bwd(x)
    assert torch.allclose(gv, base[0], rtol=1e-4, atol=1e-7)
    assert torch.allclose(ga, base[2], rtol=1e-4, atol=1e-7)


def test_tile_backward_full_size_encoder_call_vs_oracle(cuda_device):
    """Full ENC size (N = 2, Lq = S = 13 294, raster queries): the forced tile kernel and the unit-ordered kernel (the
    default: the tile schedule is opt-in, see msda_capi.cu) both give the oracle's gradients."""
    w = WORKLOADS["ENC"]
    x = device_inputs(w, seed=11, device=cuda_device, loc_mode="raster")
    n0 = _capi.kernel_launch_count()
    base = bwd(x)  # auto = unit-ordered kernel
    _capi.set_tuning("bwd_tile_mode", 2)
    forced = bwd(x)
    again = bwd(x)
    torch.cuda.synchronize()
    assert _capi.kernel_launch_count() == n0 + 6
    # grad_loc / grad_attn do not depend on the order of any atomics: bit-reproducible run to run
    assert torch.equal(forced[1], again[1]) and torch.equal(forced[2], again[2])
    check_against_oracle(x, forced)
    check_against_oracle(x, base)


@pytest.mark.parametrize("name,mode", [("C5ENC", "raster"), ("C5ENC", "unit"), ("C4ENC", "raster")])
def test_tile_backward_full_size_vs_unit_kernel(name, mode, cuda_device):
    w = WORKLOADS[name]
    x = device_inputs(w, seed=12, device=cuda_device, loc_mode=mode)
    _capi.set_tuning("bwd_tile_mode", 2)
    got = bwd(x)
    _capi.set_tuning("bwd_tile_mode", 1)
    base = bwd(x)
    torch.cuda.synchronize()
    for a, b, nm in zip(got, base, ("grad_value", "grad_loc", "grad_attn")):
        d = (a.double() - b.double()).abs()
        scale = b.double().pow(2).mean().sqrt().item()
        if nm == "grad_loc":  # floor() discontinuity: both kernels use the same fused multiply-add, so no mask is needed
            assert (d <= 1e-4 * b.double().abs() + 2e-4 * scale).all(), (nm, d.max().item(), scale)
        else:
            assert (d <= 1e-4 * b.double().abs() + 1e-4 * scale).all(), (nm, d.max().item(), scale)
    # adjointness at full size: <value, grad_value> == <attn, grad_attn> (forward is linear in both)
    a = (x["value"].double() * got[0].double()).sum()
    c = (x["attn"].double() * got[2].double()).sum()
    assert abs(a - c) <= 1e-4 * max(abs(a), abs(c)) + 1e-9, (a, c)
