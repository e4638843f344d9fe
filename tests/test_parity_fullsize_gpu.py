"""GPU parity at the FULL BASELINE.json shapes against the C oracle (restatement of ms_deform_im2col_cuda.cuh:33-159,237-299,
pinned to the reference's golden vectors by tests/test_oracle_golden.py): forward and all three gradients, fp32 (rtol 1e-4)
and bf16 storage (rtol 1e-2, oracle evaluated on the same rounded inputs).

  ENC    N=2, square pyramid, Lq = S = 13 294, raster queries with +-4 px offsets (encoder self-attention)
  C5ENC  N=2, 800x1333 pyramid, Lq = S = 22 223, raster                       (BASELINE configs[4], encoder call)
  C5DEC  N=2, 800x1333 pyramid, Lq = 300                                      (configs[4], decoder call)
  C4DEC  N=32, 800x1333 pyramid, Lq = 300: the GPU runs the whole batch, the oracle checks images 0, 13 and 31 -- every
         image is independent, so each is compared with the oracle run on that image alone.

Also here: the samples that tests/_util.assert_close_grad_loc masks (pixel coordinate within 2e-5 of an integer, where
floor() decides between two pixel pairs and grad_loc is discontinuous) must equal ONE of the two legitimate one-sided
values -- a wrong tap pair there would otherwise pass unnoticed.
"""
import numpy as np
import pytest
import torch

import aloception_oss_b200 as msda
from aloception_oss_b200 import _capi
from aloception_oss_b200.synthetic import WORKLOADS, Workload, device_inputs, level_tensors
from oracle import msda_oracle
from tests._util import assert_close, assert_close_grad, assert_close_grad_loc, check_grad_value, near_floor_discontinuity, rms

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _ops(cuda_device):
    msda.load_ops()
    yield


def gpu_fwd_bwd(x):
    out = msda.ms_deform_attn_forward(x["value"], x["shapes"], x["start"], x["loc"], x["attn"])
    gv, gl, ga = msda.ms_deform_attn_backward(x["value"], x["shapes"], x["start"], x["loc"], x["attn"], x["grad_out"])
    torch.cuda.synchronize()
    return out, gv, gl, ga


def oracle64_of(x, images=None):
    """Oracle in float64 on (a subset of the images of) x; returns numpy arrays."""
    sel = (lambda t: t) if images is None else (lambda t: t[images])
    n = {k: sel(x[k]).detach().double().cpu().numpy() for k in ("value", "loc", "attn", "grad_out")}
    shapes, start = x["shapes"].cpu().numpy(), x["start"].cpu().numpy()
    out = msda_oracle.forward(n["value"], shapes, n["loc"], n["attn"], start)
    gv, gl, ga = msda_oracle.backward(n["grad_out"], n["value"], shapes, n["loc"], n["attn"], start)
    return out, gv, gl, ga


FULL = [("ENC", "raster", None), ("C5ENC", "raster", None), ("C5DEC", "unit", None), ("C4DEC", "unit", [0, 13, 31])]


@pytest.mark.parametrize("name,mode,images", FULL, ids=[f[0] for f in FULL])
def test_full_size_fp32_vs_oracle(name, mode, images, cuda_device):
    w = WORKLOADS[name]
    x = device_inputs(w, seed=29, device=cuda_device, loc_mode=mode)
    got = [g.double().cpu() for g in gpu_fwd_bwd(x)]
    if images is not None:
        got = [g[images] for g in got]
    got = [g.numpy() for g in got]
    want = oracle64_of(x, images)
    loc = (x["loc"] if images is None else x["loc"][images]).cpu().numpy()
    assert_close(got[0], want[0], 1e-4, 1e-7 * rms(want[0]), f"{name} out")
    check_grad_value(got[1], {"grad_value": want[1]}, 1e-4, what=f"{name} grad_value")
    assert_close_grad_loc(got[2], want[2], loc, x["shapes"].cpu().numpy(), 1e-4, what=f"{name} grad_loc")
    assert_close_grad(got[3], want[3], 1e-4, f"{name} grad_attn")


@pytest.mark.parametrize("name,mode,images", FULL, ids=[f[0] for f in FULL])
def test_full_size_bf16_vs_oracle(name, mode, images, cuda_device):
    """bf16 storage, fp32 arithmetic: the oracle sees the same bf16-rounded inputs; tolerance 1e-2 (north_star)."""
    w = WORKLOADS[name]
    x = device_inputs(w, seed=30, device=cuda_device, dtype=torch.bfloat16, loc_mode=mode)
    got = [g.double().cpu() for g in gpu_fwd_bwd(x)]
    if images is not None:
        got = [g[images] for g in got]
    got = [g.numpy() for g in got]
    want = oracle64_of(x, images)
    # bf16 locations quantise the pixel coordinate to ~0.4 px on a 100-px level, so many samples sit exactly ON an
    # integer coordinate: grad_loc is compared away from the floor() discontinuities (the one-sided test below covers them)
    loc = (x["loc"] if images is None else x["loc"][images]).double().cpu().numpy()
    skip = near_floor_discontinuity(loc, x["shapes"].cpu().numpy(), eps=1e-3)
    for g, wnt, nm in zip(got, want, ("out", "grad_value", "grad_loc", "grad_attn")):
        g = g.copy()
        if nm == "grad_loc":
            g[skip] = wnt[skip]
        assert_close(g, wnt, 1e-2, 1e-2 * rms(wnt), f"{name} bf16 {nm}")


def _one_sided_variants(loc, shapes, eps):
    """For every coordinate within eps of an integer pixel coordinate: the location nudged to either side (x4 sign combos)."""
    wh = shapes[:, ::-1].astype(np.float64).reshape(1, 1, 1, -1, 1, 2)
    pix = loc * wh - 0.5
    near = np.abs(pix - np.round(pix)) < eps
    out = []
    for sx in (-1.0, 1.0):
        for sy in (-1.0, 1.0):
            sgn = np.stack([np.full(pix.shape[:-1], sx), np.full(pix.shape[:-1], sy)], -1)
            p2 = np.where(near, np.round(pix) + sgn * 4 * eps, pix)
            out.append((p2 + 0.5) / wh)
    return near.any(-1), out


@pytest.mark.parametrize("fused_tile", [1, 2], ids=["unit_kernel", "tile_kernel"])
def test_grad_loc_on_floor_discontinuities_is_one_of_the_two_one_sided_values(fused_tile, cuda_device):
    """Sampling locations placed (to fp32 rounding) ON integer pixel coordinates: every such sample's grad_loc must equal the
    oracle's value for the location nudged just below or just above the integer -- the only two legitimate answers."""
    w = Workload("edge", 2, ((20, 27), (10, 14), (5, 7), (3, 4)), 727, M=8, P=4, D=32)
    rng = np.random.default_rng(5)
    shapes, start = level_tensors(w.levels)
    wh = shapes[:, ::-1].astype(np.float64).reshape(1, 1, 1, -1, 1, 2)
    k = rng.integers(-1, 28, size=(w.N, w.Lq, w.M, w.L, w.P, 2)).astype(np.float64)
    k = np.minimum(k, wh)  # integer pixel coordinates in [-1, size]
    on_int = rng.random(k.shape) < 0.7
    pix = np.where(on_int, k, k + rng.random(k.shape))
    loc = ((pix + 0.5) / wh).astype(np.float32)
    value = (rng.random((w.N, w.S, w.M, w.D), dtype=np.float32))
    attn = rng.random((w.N, w.Lq, w.M, w.L, w.P), dtype=np.float32) + 0.1
    attn /= attn.sum((-1, -2), keepdims=True)
    go = rng.random((w.N, w.Lq, w.M * w.D), dtype=np.float32) - 0.5
    dev = cuda_device
    x = dict(value=torch.from_numpy(value).to(dev), loc=torch.from_numpy(loc).to(dev), attn=torch.from_numpy(attn).to(dev),
             grad_out=torch.from_numpy(go).to(dev), shapes=torch.from_numpy(shapes).to(dev), start=torch.from_numpy(start).to(dev))
    _capi.set_tuning("bwd_tile_mode", fused_tile)
    try:
        gl = msda.ms_deform_attn_backward(x["value"], x["shapes"], x["start"], x["loc"], x["attn"], x["grad_out"])[1]
        torch.cuda.synchronize()
    finally:
        _capi.set_tuning("bwd_tile_mode", 0)
    gl = gl.double().cpu().numpy()
    eps = 2e-5
    loc64 = loc.astype(np.float64)
    near, variants = _one_sided_variants(loc64, shapes, eps)
    assert near.mean() > 0.5  # the case is about these samples
    v64, a64, g64 = value.astype(np.float64), attn.astype(np.float64), go.astype(np.float64)
    sides = [msda_oracle.backward(g64, v64, shapes, lv, a64, start)[1] for lv in variants]
    exact = msda_oracle.backward(g64, v64, shapes, loc64, a64, start)[1]
    scale = rms(exact)
    tol = lambda ref: 1e-3 * np.abs(ref) + 1e-3 * scale  # the nudge (8e-5 px) moves the interpolation weights by ~1e-4
    ok_far = np.abs(gl - exact) <= 1e-4 * np.abs(exact) + 1e-4 * scale
    assert ok_far[~near].all(), "grad_loc away from the discontinuities"
    ok_near = np.zeros(gl.shape, dtype=bool)
    for sd in sides:
        ok_near |= np.abs(gl - sd) <= tol(sd)
    # per SAMPLE: one variant has to explain both components (the same floor decision yields x and y)
    ok_sample = np.zeros(near.shape, dtype=bool)
    for sd in sides:
        ok_sample |= (np.abs(gl - sd) <= tol(sd)).all(-1)
    bad = near & ~ok_sample
    assert not bad.any(), f"{bad.sum()} of {near.sum()} on-edge samples match neither one-sided gradient"
