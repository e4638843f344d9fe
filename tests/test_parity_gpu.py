"""GPU parity: the sm_100a operator (through the C ABI, via the reference-shaped Python surface) against the
CPU oracle and the committed golden vectors of the reference.

Tolerances (BASELINE.json north_star): fp32 1e-4 rtol, bf16/f16 1e-2 rtol; fp64 ~1e-10.  Gradient tensors add an
absolute term of rtol * rms(reference) -- see tests/_util.assert_close_grad for why a purely relative bound
cannot hold for fp32 channel sums.
"""
import numpy as np
import pytest
import torch

import aloception_oss_b200 as msda
from aloception_oss_b200 import _capi
from aloception_oss_b200.synthetic import WORKLOADS, Workload, device_inputs, torch_inputs
from oracle import msda_oracle
from tests._util import (assert_close, assert_close_grad, assert_close_grad_loc, check_grad_value, golden_names,
                         load_golden, rms)

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _ops(cuda_device):
    msda.load_ops()
    for k in ("force_generic", "fwd_unroll", "bwd_unroll", "warps_per_block", "no_pdl", "head_major", "smem_records", "patch_mode", "patch_px",
              "patch_py", "patch_ctas", "staged_mode", "staged_kb", "staged_warps", "staged_variant", "zero_mode", "zero_ctas",
              "zero_threads", "zero_chunk_kb", "spec_mode", "bwd_tile_mode", "bwd_tile_ctas", "bwd_two_pass", "fwd_pair_mode", "fwd_pair_px",
              "fwd_pair_py", "fwd_pair_ctas", "fwd_win_mode", "fwd_win_ctas"):
        _capi.set_tuning(k, 0)
    yield


def run_op(x, dev, dtype=None, need_grad=True):
    """x: dict of CPU tensors -> (out, gv, gl, ga) as float64 numpy, computed on the GPU."""
    f = lambda t: (t.to(dtype) if dtype is not None else t).to(dev)
    value, loc, attn, go = f(x["value"]), f(x["loc"]), f(x["attn"]), f(x["grad_out"])
    shapes, start = x["shapes"].to(dev), x["start"].to(dev)
    if need_grad:
        value.requires_grad_(True)
        loc.requires_grad_(True)
        attn.requires_grad_(True)
    out = msda.MSDeformAttnFunction.apply(value, shapes, start, loc, attn, 64)
    res = [out.detach().double().cpu().numpy()]
    if need_grad:
        out.backward(go)
        res += [t.grad.double().cpu().numpy() for t in (value, loc, attn)]
    torch.cuda.synchronize()
    return res


def oracle64(x):
    n = {k: (v.double().numpy() if v.is_floating_point() else v.numpy()) for k, v in x.items()}
    out = msda_oracle.forward(n["value"], n["shapes"], n["loc"], n["attn"], n["start"])
    gv, gl, ga = msda_oracle.backward(n["grad_out"], n["value"], n["shapes"], n["loc"], n["attn"], n["start"])
    return out, gv, gl, ga


# ---------------------------------------------------------------------------------------------------------
# golden vectors of the reference
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_names())
def test_golden(name, cuda_device):
    w, x, ref = load_golden(name)
    xt = {k: torch.from_numpy(v) for k, v in x.items()}
    out, gv, gl, ga = run_op(xt, cuda_device)
    if x["value"].dtype == np.float64:  # the reference's fp64 check (ops/test.py:36-66) + gradcheck cases
        assert_close(out, ref["out"], 1e-10, 1e-15, "out")
        assert_close(gl, ref["grad_loc"], 1e-9, 1e-14, "grad_loc")
        assert_close(ga, ref["grad_attn"], 1e-9, 1e-14, "grad_attn")
        check_grad_value(gv, ref, 1e-9, 1e-14)
    else:
        assert_close(out, ref["out64"], 1e-4, 1e-4 * 1e-3 * rms(ref["out64"]), "out")
        assert_close_grad_loc(gl, ref["grad_loc64"], x["loc"], x["shapes"], 1e-4)
        assert_close_grad(ga, ref["grad_attn64"], 1e-4, "grad_attn")
        check_grad_value(gv, ref, 1e-4)


# ---------------------------------------------------------------------------------------------------------
# oracle comparisons over shapes / dtypes / kernels
# ---------------------------------------------------------------------------------------------------------
SHAPES = [
    Workload("d32", 2, ((12, 16), (6, 8), (3, 4), (2, 2)), 37, M=8, P=4, D=32),
    Workload("d32_p3_l3", 1, ((9, 5), (4, 3), (2, 1)), 19, M=5, P=3, D=32),     # L*P = 9: ragged lane groups
    Workload("d64", 2, ((7, 9), (3, 5)), 11, M=4, P=4, D=64),
    Workload("d16", 2, ((7, 9), (3, 5)), 11, M=4, P=2, D=16),
    Workload("d128", 1, ((7, 9), (3, 5)), 5, M=2, P=4, D=128),
    Workload("d30", 1, ((6, 4), (3, 2)), 2, M=2, P=2, D=30),                     # generic kernel (ops/test.py:130)
    Workload("d71", 1, ((6, 4), (3, 2)), 2, M=2, P=2, D=71),
    Workload("d1", 2, ((1, 1), (2, 3)), 3, M=1, P=1, D=1),
    Workload("l1", 1, ((16, 16),), 33, M=8, P=4, D=32),
    Workload("p17", 1, ((8, 8), (4, 4)), 9, M=2, P=17, D=32),
]


@pytest.mark.parametrize("w", SHAPES, ids=lambda w: w.name)
@pytest.mark.parametrize("mode", ["unit", "wide"])
def test_fp32_vs_oracle(w, mode, cuda_device):
    x = torch_inputs(w, seed=13, loc_mode=mode)
    got = run_op(x, cuda_device)
    want = oracle64(x)
    assert_close(got[0], want[0], 1e-4, 1e-7 * rms(want[0]), "out")
    check_grad_value(got[1], {"grad_value": want[1]}, 1e-4)
    assert_close_grad_loc(got[2], want[2], x["loc"].numpy(), x["shapes"].numpy(), 1e-4)
    assert_close_grad(got[3], want[3], 1e-4, "grad_attn")


@pytest.mark.parametrize("w", SHAPES, ids=lambda w: w.name)
def test_fp64_vs_oracle(w, cuda_device):
    x = torch_inputs(w, seed=14, loc_mode="wide", dtype=torch.float64)
    got = run_op(x, cuda_device)
    want = oracle64(x)
    for g, wnt, n in zip(got, want, ("out", "grad_value", "grad_loc", "grad_attn")):
        assert_close(g, wnt, 1e-9, 1e-13, n)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
@pytest.mark.parametrize("w", SHAPES[:6], ids=lambda w: w.name)
def test_16bit_vs_oracle(w, dtype, cuda_device):
    """16-bit storage, fp32 arithmetic: compare with the oracle evaluated on the SAME rounded inputs."""
    x = torch_inputs(w, seed=15, loc_mode="unit")
    xr = {k: (v.to(dtype).float() if v.is_floating_point() else v) for k, v in x.items()}
    got = run_op(xr, cuda_device, dtype=dtype)
    want = oracle64(xr)
    assert_close(got[0], want[0], 1e-2, 1e-2 * rms(want[0]), "out")
    assert_close(got[1], want[1], 1e-2, 1e-2 * rms(want[1]), "grad_value")
    assert_close(got[2], want[2], 1e-2, 1e-2 * rms(want[2]), "grad_loc")
    assert_close(got[3], want[3], 1e-2, 1e-2 * rms(want[3]), "grad_attn")


def _random_workloads(n, seed):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        L = int(rng.integers(1, 6))
        levels = tuple((int(rng.integers(1, 23)), int(rng.integers(1, 23))) for _ in range(L))
        D = int(rng.choice([1, 3, 8, 16, 24, 32, 32, 32, 40, 64, 96, 128]))
        M = int(rng.choice([1, 2, 3, 5, 8, 8]))
        P = int(rng.choice([1, 2, 3, 4, 4, 7, 9]))
        out.append(Workload(f"rnd{i}_L{L}_M{M}_P{P}_D{D}", int(rng.integers(1, 4)), levels, int(rng.integers(1, 70)), M=M, P=P, D=D))
    return out


@pytest.mark.parametrize("w", _random_workloads(48, 20261017), ids=lambda w: w.name)
def test_random_shapes_vs_oracle(w, cuda_device):
    """Shape fuzz: random pyramids (levels down to 1 x 1), head / point / channel counts on and off the vector kernels' grid,
    out-of-range locations -- fp32 forward and all three gradients against the fp64 oracle; bf16 forward where it applies."""
    import zlib
    x = torch_inputs(w, seed=zlib.crc32(w.name.encode()) % 1000, loc_mode="wide")
    got = run_op(x, cuda_device)
    want = oracle64(x)
    assert_close(got[0], want[0], 1e-4, 1e-7 * max(rms(want[0]), 1e-30), "out")
    check_grad_value(got[1], {"grad_value": want[1]}, 1e-4)
    assert_close_grad_loc(got[2], want[2], x["loc"].numpy(), x["shapes"].numpy(), 1e-4)
    assert_close_grad(got[3], want[3], 1e-4, "grad_attn")
    xr = {k: (v.bfloat16().float() if v.is_floating_point() else v) for k, v in x.items()}
    got16 = run_op(xr, cuda_device, dtype=torch.bfloat16, need_grad=False)
    want16 = oracle64(xr)
    assert_close(got16[0], want16[0], 1e-2, 1e-2 * max(rms(want16[0]), 1e-30), "out (bf16)")


@pytest.mark.parametrize("which", ["loc", "attn", "both"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
@pytest.mark.parametrize("w", SHAPES[:6], ids=lambda w: w.name)
def test_mixed_precision_fp32_locations_and_weights(w, dtype, which, cuda_device):
    """MSDA_LOC_F32 / MSDA_ATTN_F32: what torch.autocast hands the operator -- 16-bit value / grad_output, fp32 sampling_loc
    and / or attn_weight.  The fp32 tensors are read unrounded (oracle on value / grad_out rounded, loc / attn exact) and
    their gradients come back in fp32."""
    x = torch_inputs(w, seed=16, loc_mode="wide")
    keep = {"loc": ("loc",), "attn": ("attn",), "both": ("loc", "attn")}[which]
    xr = {k: (v.to(dtype).float() if v.is_floating_point() and k not in keep else v) for k, v in x.items()}
    dev = cuda_device
    value, go = xr["value"].to(dtype).to(dev).requires_grad_(True), xr["grad_out"].to(dtype).to(dev)
    loc = (xr["loc"] if "loc" in keep else xr["loc"].to(dtype)).to(dev).requires_grad_(True)
    attn = (xr["attn"] if "attn" in keep else xr["attn"].to(dtype)).to(dev).requires_grad_(True)
    out = msda.MSDeformAttnFunction.apply(value, xr["shapes"].to(dev), xr["start"].to(dev), loc, attn, 64)
    out.backward(go)
    assert out.dtype == dtype and value.grad.dtype == dtype
    assert loc.grad.dtype == loc.dtype and attn.grad.dtype == attn.dtype
    want = oracle64(xr)
    got = [t.detach().double().cpu().numpy() for t in (out, value.grad, loc.grad, attn.grad)]
    assert_close(got[0], want[0], 1e-2, 1e-2 * rms(want[0]), "out")
    assert_close(got[1], want[1], 1e-2, 1e-2 * rms(want[1]), "grad_value")
    # the fp32 gradients are not rounded to 16 bits: only value / grad_output's rounding (already in `want`) and fp32 sums remain
    if "loc" in keep:
        assert_close_grad_loc(got[2], want[2], xr["loc"].numpy(), xr["shapes"].numpy(), 1e-4)
    else:
        assert_close(got[2], want[2], 1e-2, 1e-2 * rms(want[2]), "grad_loc")
    assert_close_grad(got[3], want[3], 1e-4 if "attn" in keep else 1e-2, "grad_attn")
    # forward only, through the functional entry and the registered op
    with torch.no_grad():
        o2 = msda.ms_deform_attn_forward(value, xr["shapes"].to(dev), xr["start"].to(dev), loc, attn)
        o3 = torch.ops.alonet_custom.ms_deform_attn_forward(value, xr["shapes"].to(dev), xr["start"].to(dev), loc, attn, 64)
    assert torch.equal(o2, out) and torch.equal(o3, out)


@pytest.mark.parametrize("det", [False, True], ids=["reds", "deterministic"])
def test_mixed_precision_full_size_call_and_buffers(det, cuda_device):
    """An encoder-sized mixed-precision call (bf16 value, fp32 loc / attn) against the all-fp32 operator on the same rounded
    value / grad_output: same function up to bf16 rounding of out / grad_value; caller-provided fp32 gradient buffers."""
    w = Workload("enc", 2, ((50, 67), (25, 34), (13, 17), (7, 9)), 4484, M=8, P=4, D=32)
    x = torch_inputs(w, seed=21, loc_mode="wide")
    dev = cuda_device
    shapes, start = x["shapes"].to(dev), x["start"].to(dev)
    v16, go16 = x["value"].bfloat16().to(dev), x["grad_out"].bfloat16().to(dev)
    loc, attn = x["loc"].to(dev), x["attn"].to(dev)
    out = msda.ms_deform_attn_forward(v16, shapes, start, loc, attn)
    ref = msda.ms_deform_attn_forward(v16.float(), shapes, start, loc, attn)
    assert out.dtype == torch.bfloat16
    assert_close(out.double().cpu().numpy(), ref.double().cpu().numpy(), 1e-2, 1e-2 * rms(ref.double().cpu().numpy()), "out")
    grads = [torch.empty_like(v16), torch.full_like(loc, float("nan")), torch.full_like(attn, float("nan"))]
    got = msda.ms_deform_attn_backward(v16, shapes, start, loc, attn, go16, grads=grads, deterministic=det)
    want = msda.ms_deform_attn_backward(v16.float(), shapes, start, loc, attn, go16.float(), deterministic=det)
    assert got[1].data_ptr() == grads[1].data_ptr() and got[1].dtype == torch.float32 and got[2].dtype == torch.float32
    n = lambda t: t.double().cpu().numpy()
    assert_close(n(got[0]), n(want[0]), 1e-2, 1e-2 * rms(n(want[0])), "grad_value")
    assert_close_grad_loc(n(got[1]), n(want[1]), x["loc"].numpy(), x["shapes"].numpy(), 1e-4)
    assert_close_grad(n(got[2]), n(want[2]), 1e-4, "grad_attn")
    if det:
        again = msda.ms_deform_attn_backward(v16, shapes, start, loc, attn, go16, deterministic=True)
        assert all(torch.equal(a, b) for a, b in zip(got, again))


def test_mixed_precision_rejections(cuda_device):
    dev = cuda_device
    x = torch_inputs(SHAPES[0], seed=3)
    shapes, start = x["shapes"].to(dev), x["start"].to(dev)
    # fp32 value with 16-bit locations is not a mixed mode
    with pytest.raises(RuntimeError, match="sampling_loc has dtype"):
        msda.ms_deform_attn_forward(x["value"].to(dev), shapes, start, x["loc"].bfloat16().to(dev), x["attn"].to(dev))
    # float64 next to 16-bit value neither
    with pytest.raises(RuntimeError, match="attn_weight has dtype"):
        msda.ms_deform_attn_forward(x["value"].bfloat16().to(dev), shapes, start, x["loc"].to(dev), x["attn"].double().to(dev))
    # grad_output must have value's dtype
    with pytest.raises(RuntimeError, match="grad_output"):
        msda.ms_deform_attn_backward(x["value"].bfloat16().to(dev), shapes, start, x["loc"].to(dev), x["attn"].to(dev),
                                     x["grad_out"].to(dev))


@pytest.mark.parametrize("w", [SHAPES[0], SHAPES[1], SHAPES[9]], ids=lambda w: w.name)
@pytest.mark.parametrize("no_pdl", [0, 1])
@pytest.mark.parametrize("knob,val", [("force_generic", 1), ("fwd_unroll", 2), ("fwd_unroll", 4), ("bwd_unroll", 2),
                                      ("bwd_unroll", 4), ("warps_per_block", 3), ("warps_per_block", 8), ("head_major", 2),
                                      ("smem_records", 1), ("smem_records", 2), ("patch_mode", 2),
                                      ("staged_mode", 2), ("zero_mode", 2), ("zero_threads", 64), ("spec_mode", 1),
                                      ("bwd_two_pass", 2)])
def test_kernel_variants_agree(knob, val, no_pdl, w, cuda_device):
    x = torch_inputs(w, seed=16, loc_mode="wide")
    want = oracle64(x)
    _capi.set_tuning("no_pdl", no_pdl)
    _capi.set_tuning(knob, val)
    got = run_op(x, cuda_device)
    assert_close(got[0], want[0], 1e-4, 1e-7 * rms(want[0]), "out")
    check_grad_value(got[1], {"grad_value": want[1]}, 1e-4)
    assert_close_grad_loc(got[2], want[2], x["loc"].numpy(), x["shapes"].numpy(), 1e-4)
    assert_close_grad(got[3], want[3], 1e-4, "grad_attn")


@pytest.mark.parametrize("lq_delta", [0, 7, -5], ids=["Lq=S", "Lq>S", "Lq<S"])
@pytest.mark.parametrize("px,py", [(8, 16), (8, 8), (3, 2), (16, 4)])
@pytest.mark.parametrize("dtype", [None, torch.bfloat16], ids=["f32", "bf16"])
def test_patch_ordered_forward_is_the_same_function(lq_delta, px, py, dtype, cuda_device):
    """The patch-ordered persistent forward (encoder scheduling) visits every query exactly once for any level shapes and
    query count, and returns bit-identical results to the unit-ordered kernel (same per-unit arithmetic)."""
    levels = ((13, 21), (7, 11), (4, 6), (2, 3))
    S = sum(h * w for h, w in levels)
    w = Workload("enc_small", 2, levels, S + lq_delta, M=8, P=4, D=32)
    x = torch_inputs(w, seed=23, loc_mode="wide")
    _capi.set_tuning("patch_mode", 1)
    base = run_op(x, cuda_device, dtype=dtype, need_grad=False)[0]
    _capi.set_tuning("patch_mode", 2)
    _capi.set_tuning("patch_px", px)
    _capi.set_tuning("patch_py", py)
    got = run_op(x, cuda_device, dtype=dtype, need_grad=False)[0]
    assert np.array_equal(got, base)
    if dtype is None:
        want = oracle64(x)[0]
        assert_close(got, want, 1e-4, 1e-7 * rms(want), "out")


@pytest.mark.parametrize("levels,lq,M,P", [
    (((20, 27), (10, 14), (5, 7), (3, 4)), 700, 8, 4),    # out-of-range samples on every level
    (((9, 1), (1, 7), (2, 2), (1, 1)), 300, 8, 4),         # levels narrower than 2 pixels: the flagged path must take over
    (((16, 16), (8, 8), (2, 3)), 200, 8, 17),              # L*P = 51: two passes
    (((2, 2), (16, 16), (1, 9)), 150, 8, 12),              # L*P = 36: the degenerate level only shows up in the second pass
    (((16, 16), (8, 8)), 100, 5, 4),                       # run-time head count
])
@pytest.mark.parametrize("poison", [None, float("nan"), float("inf")], ids=["finite", "nan", "inf"])
@pytest.mark.parametrize("dtype", [None, torch.bfloat16, torch.float16], ids=["f32", "bf16", "f16"])
@pytest.mark.parametrize("patch", [1, 2], ids=["unit-ordered", "patch-ordered"])
def test_speculative_gather_is_the_same_function(levels, lq, M, P, poison, dtype, patch, cuda_device):
    """The speculative regular-window forward gather (taps outside a level get a zero WEIGHT on a clamped in-range address
    instead of a zero-line ADDRESS) returns bit-identical results to the flagged path -- also when value holds NaN / Inf next
    to out-of-range samples (0 * Inf would differ: such units are detected by their non-finite sums and redone)."""
    if patch == 2 and P * len(levels) > 32:
        pytest.skip("patch-ordered kernel needs L*P <= 32")
    w = Workload("spec_small", 2, levels, lq, M=M, P=P, D=32)
    x = torch_inputs(w, seed=31, loc_mode="wide")
    if poison is not None:
        x["value"][:, ::29] = poison  # sparse: about half of the units stay finite, and clamped zero-weight taps do hit poisoned pixels
    _capi.set_tuning("patch_mode", 1)
    _capi.set_tuning("spec_mode", 1)
    base = run_op(x, cuda_device, dtype=dtype, need_grad=False)[0]
    _capi.set_tuning("patch_mode", patch)
    _capi.set_tuning("spec_mode", 2)
    got = run_op(x, cuda_device, dtype=dtype, need_grad=False)[0]
    assert np.array_equal(got, base, equal_nan=True)
    if poison is not None and lq >= 300:
        assert np.isfinite(base).any() and not np.isfinite(base).all()


@pytest.mark.parametrize("levels,lq,M,P", [
    (((13, 21), (7, 11), (4, 6), (2, 3)), 392, 8, 4),   # pixel-aligned queries (Lq == S), L*P = 16: the encoder case
    (((13, 21), (7, 11), (4, 6), (2, 3)), 450, 8, 4),   # more queries than pixels: the SM-affine order's plain-order tail
    (((13, 21), (7, 11), (4, 6), (2, 3)), 300, 8, 4),   # fewer queries than pixels: its level grids overshoot Lq
    (((16, 16), (8, 8)), 320, 5, 4),                       # odd run-time head count: the last pair of a query has one head
    (((16, 16), (8, 8), (4, 4)), 336, 8, 3),               # L*P = 9: ragged rounds, idle lanes in both halves
    (((9, 11),), 99, 2, 2),                                # one level, one pair
    (((12, 1), (5, 7)), 47, 4, 4),                         # a level narrower than 2 pixels: no regular window, flagged path
])
@pytest.mark.parametrize("poison", [None, float("nan")], ids=["finite", "nan"])
@pytest.mark.parametrize("dtype", [None, torch.bfloat16, torch.float16, "mixed"], ids=["f32", "bf16", "f16", "bf16+f32loc"])
@pytest.mark.parametrize("mode,px,py", [(2, 0, 0), (3, 3, 3), (3, 2, 1), (3, 1, 4)], ids=["static", "affine8x8", "affine4x2", "affine2x16"])
def test_paired_forward_is_the_same_function(levels, lq, M, P, poison, dtype, mode, px, py, cuda_device):
    """The paired forward (one warp = two heads of a query; knob fwd_pair_mode = 2) and its SM-affine patch order (= 3, scheduling
    words from msda_forward_ws' workspace) are pure re-schedulings of the unit-ordered forward: bit-identical outputs, every
    (image, query, head) written exactly once whatever the level shapes, including units that fall back to the flagged path."""
    w = Workload("pair_small", 3, levels, lq, M=M, P=P, D=32)
    x = torch_inputs(w, seed=37, loc_mode="wide")
    if poison is not None:
        x["value"][:, ::29] = poison
    dev = cuda_device
    vt = torch.bfloat16 if dtype == "mixed" else dtype
    f = lambda t, d: (t.to(d) if d is not None else t).to(dev)
    lt = None if dtype == "mixed" else dtype
    value, loc, attn = f(x["value"], vt), f(x["loc"], lt), f(x["attn"], lt)
    shapes, start = x["shapes"].to(dev), x["start"].to(dev)
    _capi.set_tuning("spec_mode", 2)
    _capi.set_tuning("fwd_pair_mode", 1)
    base = msda.ms_deform_attn_forward(value, shapes, start, loc, attn)
    n0 = _capi.kernel_launch_count()
    _capi.set_tuning("fwd_pair_mode", mode)
    _capi.set_tuning("fwd_pair_px", px)
    _capi.set_tuning("fwd_pair_py", py)
    out = torch.full_like(base, float("inf"))  # poisoned buffer: an unwritten row would show
    got = msda.ms_deform_attn_forward(value, shapes, start, loc, attn, out=out)
    torch.cuda.synchronize()
    for k in ("fwd_pair_mode", "fwd_pair_px", "fwd_pair_py", "spec_mode"):
        _capi.set_tuning(k, 0)
    assert _capi.kernel_launch_count() == n0 + 1
    assert torch.equal(torch.nan_to_num(got.float(), nan=12345.0), torch.nan_to_num(base.float(), nan=12345.0))
    if poison is not None:
        assert torch.isfinite(base.float()).any() and not torch.isfinite(base.float()).all()


@pytest.mark.parametrize("levels,lq,M,P", [
    (((13, 21), (7, 11), (4, 6), (2, 3)), 392, 8, 4),   # pixel-aligned queries (Lq == S), L*P = 16: the encoder case
    (((13, 21), (7, 11), (4, 6), (2, 3)), 450, 8, 4),   # more queries than pixels: plain-order tail
    (((13, 21), (7, 11), (4, 6), (2, 3)), 300, 8, 4),   # fewer queries than pixels: the level grids overshoot Lq
    (((40, 40), (20, 20)), 2000, 3, 4),                    # run-time head count; boxes of the fine level overflow the window budget
    (((16, 16), (8, 8), (4, 4)), 336, 8, 3),               # L*P = 9
    (((12, 1), (5, 7)), 47, 4, 4),                         # a level narrower than 2 pixels: flagged path
])
@pytest.mark.parametrize("poison", [None, float("nan")], ids=["finite", "nan"])
@pytest.mark.parametrize("mode", ["raster-ish", "wide"])
def test_windowed_forward_is_the_same_function(levels, lq, M, P, poison, mode, cuda_device):
    """The windowed forward (msda_fwd_win.cuh, knob fwd_win_mode = 2: per-level boxes of `value` staged in shared memory per
    query tile) is a pure re-scheduling of the unit-ordered forward: bit-identical outputs for local and non-local sampling
    locations, staged and unstaged levels, every (image, query, head) written exactly once."""
    w = Workload("win_small", 3, levels, lq, M=M, P=P, D=32)
    x = torch_inputs(w, seed=41, loc_mode="wide")
    if mode == "raster-ish":  # locations near the query's own pixel (as far as the query is a pixel), +-3 px
        S = sum(h * wd for h, wd in levels)
        ref = []
        for h, wd in levels:
            ys, xs = torch.meshgrid((torch.arange(h) + 0.5) / h, (torch.arange(wd) + 0.5) / wd, indexing="ij")
            ref.append(torch.stack([xs.reshape(-1), ys.reshape(-1)], -1))
        ref = torch.cat(ref, 0)
        ref = torch.cat([ref, torch.rand(max(0, lq - S), 2)], 0)[:lq]
        wh = x["shapes"].flip(-1).float().view(1, 1, 1, len(levels), 1, 2)
        x["loc"] = (ref.view(1, lq, 1, 1, 1, 2) + (torch.rand(x["loc"].shape) - 0.5) * 6.0 / wh).contiguous()
    if poison is not None:
        x["value"][:, ::29] = poison
    dev = cuda_device
    value, loc, attn = x["value"].to(dev), x["loc"].to(dev), x["attn"].to(dev)
    shapes, start = x["shapes"].to(dev), x["start"].to(dev)
    _capi.set_tuning("fwd_win_mode", 1)
    base = msda.ms_deform_attn_forward(value, shapes, start, loc, attn)
    _capi.set_tuning("fwd_win_mode", 2)
    out = torch.full_like(base, float("inf"))
    got = msda.ms_deform_attn_forward(value, shapes, start, loc, attn, out=out)
    torch.cuda.synchronize()
    _capi.set_tuning("fwd_win_mode", 0)
    assert torch.equal(torch.nan_to_num(got, nan=12345.0), torch.nan_to_num(base, nan=12345.0))
    if poison is not None:
        assert torch.isfinite(base).any() and not torch.isfinite(base).all()


def test_paired_forward_through_the_registered_op_and_under_graph_capture(cuda_device):
    """fwd_pair_mode = 3 through torch.ops (C++ shim or Python registration: both query msda_forward_workspace_bytes and pass a
    caching-allocator buffer to msda_forward_ws) and inside a CUDA graph (the scheduling words are zeroed by a memset node)."""
    w = WORKLOADS["ENC"]
    x = device_inputs(w, seed=9, device=cuda_device, loc_mode="raster")
    args = (x["value"], x["shapes"], x["start"], x["loc"], x["attn"], 64)
    base = torch.ops.alonet_custom.ms_deform_attn_forward(*args)
    _capi.set_tuning("fwd_pair_mode", 3)
    try:
        dims = _capi.MsdaDims(w.N, w.S, w.M, w.D, w.L, w.Lq, w.P)
        import ctypes
        assert _capi.lib().msda_forward_workspace_bytes(ctypes.byref(dims), _capi.F32) > 0
        got = torch.ops.alonet_custom.ms_deform_attn_forward(*args)
        assert torch.equal(got, base)
        g = torch.cuda.CUDAGraph()
        static = [None]
        with torch.cuda.graph(g):
            static[0] = torch.ops.alonet_custom.ms_deform_attn_forward(*args)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        assert torch.equal(static[0], base)
    finally:
        _capi.set_tuning("fwd_pair_mode", 0)
    import ctypes
    assert _capi.lib().msda_forward_workspace_bytes(ctypes.byref(dims), _capi.F32) == 0  # default: no schedule needs scratch


@pytest.mark.parametrize("levels,lq,M,P", [
    (((13, 21), (7, 11), (4, 6), (2, 3)), 450, 8, 4),   # L*P = 16: unrolled instantiation; all but level 0 fit 4 KB
    (((9, 11),), 333, 8, 4),                               # one level: the whole image is the tile (or nothing fits)
    (((16, 16), (8, 8), (4, 4)), 500, 8, 3),               # ragged rounds (L*P = 9)
    (((16, 16), (8, 8)), 100, 5, 4),                       # run-time head count
    (((6, 5), (12, 10), (3, 2)), 200, 8, 4),               # levels not sorted by size: still a contiguous suffix
])
@pytest.mark.parametrize("kb", [0, 1, 4, 16])
@pytest.mark.parametrize("variant,warps", [(0, 0), (1, 0), (0, 5)])
@pytest.mark.parametrize("dtype", [None, torch.bfloat16], ids=["f32", "bf16"])
def test_tma_staged_forward_is_the_same_function(levels, lq, M, P, kb, variant, warps, dtype, cuda_device):
    """The TMA-staged persistent forward (coarse levels of one (image, head) in shared memory) returns bit-identical
    results to the unit-ordered kernel whatever part of the pyramid fits its tile (kb = tile budget; 0 = all it can get),
    including out-of-range samples (zero row) and rounds that mix staged and unstaged levels."""
    w = Workload("staged_small", 2, levels, lq, M=M, P=P, D=32)
    x = torch_inputs(w, seed=29, loc_mode="wide")
    _capi.set_tuning("staged_mode", 1)
    base = run_op(x, cuda_device, dtype=dtype, need_grad=False)[0]
    _capi.set_tuning("staged_mode", 2)
    _capi.set_tuning("staged_kb", kb)
    _capi.set_tuning("staged_variant", variant)
    _capi.set_tuning("staged_warps", warps)
    n0 = _capi.kernel_launch_count()
    got = run_op(x, cuda_device, dtype=dtype, need_grad=False)[0]
    assert _capi.kernel_launch_count() == n0 + 1
    assert np.array_equal(got, base)
    if dtype is None:
        want = oracle64(x)[0]
        assert_close(got, want, 1e-4, 1e-7 * rms(want), "out")


def test_tma_staged_forward_full_size_encoder_call(cuda_device):
    """Encoder shape (Lq = S = 13 294, N = 2): staged forward == unit-ordered forward, bit for bit."""
    w = WORKLOADS["ENC"]
    s = device_inputs(w, seed=41, device=cuda_device, loc_mode="raster")
    f = lambda: msda.ms_deform_attn_forward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"])
    _capi.set_tuning("staged_mode", 1)
    base = f()
    _capi.set_tuning("staged_mode", 2)
    got = f()
    torch.cuda.synchronize()
    assert torch.equal(got, base)


@pytest.mark.parametrize("channels", [30, 32, 64, 71])
def test_gradient_numerical_like_the_reference_op_test(channels, cuda_device):
    """The reference's own gradient test (alonet/deformable_detr/ops/test.py:88-131): torch.autograd.gradcheck through
    MSDeformAttnFunction in float64 at N, M = 1, 2; Lq, L, P = 2, 2, 2; levels (6,4), (3,2); D in {30, 32, 64, 71}."""
    N, M, Lq, L, P = 1, 2, 2, 2, 2
    shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.int32, device=cuda_device)
    start = torch.cat((shapes.new_zeros((1,)), (shapes[:, 0] * shapes[:, 1]).cumsum(0)[:-1].to(torch.int32)))
    S = int((shapes[:, 0] * shapes[:, 1]).sum())
    torch.manual_seed(3)
    value = (torch.rand(N, S, M, channels, device=cuda_device) * 0.01).double().requires_grad_(True)
    loc = torch.rand(N, Lq, M, L, P, 2, device=cuda_device).double().requires_grad_(True)
    attn = torch.rand(N, Lq, M, L, P, device=cuda_device) + 1e-5
    attn = (attn / attn.sum(-1, keepdim=True).sum(-2, keepdim=True)).double().requires_grad_(True)
    assert torch.autograd.gradcheck(msda.MSDeformAttnFunction.apply, (value, shapes, start, loc, attn, 2))


# ---------------------------------------------------------------------------------------------------------
# edge cases
# ---------------------------------------------------------------------------------------------------------
def test_exact_window_edges_and_pixel_centres(cuda_device):
    """Locations exactly on the skip-window borders, pixel centres and integer coordinates."""
    w = Workload("edge", 1, ((4, 4), (2, 2)), 4, M=1, P=4, D=32)
    x = torch_inputs(w, seed=1)
    vals = torch.tensor([-0.125, 0.0, 0.125, 0.375, 0.5, 0.875, 1.0, 1.125, 1.25, -0.25])
    idx = torch.arange(x["loc"].numel()) % vals.numel()
    x["loc"] = vals[idx].view_as(x["loc"]).contiguous()
    got = run_op(x, cuda_device, need_grad=False)
    want = oracle64(x)
    assert_close(got[0], want[0], 1e-5, 1e-9, "out")


def test_all_samples_outside_gives_zero(cuda_device):
    w = Workload("outside", 1, ((5, 5),), 3, M=2, P=2, D=32)
    x = torch_inputs(w, seed=2)
    x["loc"] = torch.full_like(x["loc"], 7.5)
    out, gv, gl, ga = run_op(x, cuda_device)
    assert not out.any() and not gv.any() and not gl.any() and not ga.any()


@pytest.mark.parametrize("spec", [1, 2], ids=["flagged", "speculative"])
def test_wild_locations_contribute_nothing(spec, cuda_device):
    """NaN, +-Inf and astronomically large sampling locations fall outside every level: such samples contribute exactly
    zero to out / grad_value / grad_attn (the reference's window test is false for them), on both forward gather paths."""
    w = Workload("wild", 2, ((9, 11), (5, 6), (3, 3)), 40, M=8, P=4, D=32)
    x = torch_inputs(w, seed=5, loc_mode="wide")
    wild = torch.tensor([float("nan"), float("inf"), float("-inf"), 1e30, -1e30, 3e9, -7.5])
    loc = x["loc"]
    mask = torch.zeros(loc.shape[:-1], dtype=torch.bool)
    mask.view(-1)[::3] = True  # every third sample is wild
    loc[mask] = wild[torch.arange(int(mask.sum())) % wild.numel()].unsqueeze(-1).expand(-1, 2)
    x["loc"] = loc.contiguous()
    _capi.set_tuning("spec_mode", spec)
    out, gv, gl, ga = run_op(x, cuda_device)
    # reference semantics: drop the wild samples (zero attention weight on a harmless location)
    y = {k: v.clone() for k, v in x.items()}
    y["attn"][mask] = 0.0
    y["loc"][mask] = 0.5
    want = oracle64(y)
    assert np.isfinite(out).all() and np.isfinite(gv).all()
    assert_close(out, want[0], 1e-4, 1e-7 * rms(want[0]), "out")
    check_grad_value(gv, {"grad_value": want[1]}, 1e-4)
    assert not ga[mask.numpy()].any()


@pytest.mark.parametrize("field,val", [("N", 0), ("Lq", 0)])
def test_empty_inputs(field, val, cuda_device):
    base = dict(N=2, Lq=5)
    base[field] = val
    w = Workload("empty", base["N"], ((3, 3),), base["Lq"], M=2, P=2, D=32)
    x = torch_inputs(w, seed=2)
    out, gv, gl, ga = run_op(x, cuda_device)
    assert out.shape == (w.N, w.Lq, w.M * w.D)
    assert gv.shape == (w.N, w.S, w.M, w.D) and not gv.any()


def test_unaligned_views_use_generic_path(cuda_device):
    """A value view that starts 4 bytes into its storage cannot use 128-bit loads; result must not change."""
    w = SHAPES[0]
    x = torch_inputs(w, seed=17)
    want = oracle64(x)
    dev = cuda_device
    buf = torch.empty(x["value"].numel() + 1, device=dev)
    v = buf[1:].view_as(x["value"])
    v.copy_(x["value"])
    assert v.is_contiguous() and v.data_ptr() % 16 != 0
    out = msda.ms_deform_attn_forward(v, x["shapes"].to(dev), x["start"].to(dev), x["loc"].to(dev), x["attn"].to(dev))
    assert_close(out.double().cpu().numpy(), want[0], 1e-4, 1e-7 * rms(want[0]))


def test_int64_level_tensors_accepted(cuda_device):
    w = SHAPES[0]
    x = torch_inputs(w, seed=18)
    dev = cuda_device
    a = msda.ms_deform_attn_forward(x["value"].to(dev), x["shapes"].to(dev), x["start"].to(dev), x["loc"].to(dev), x["attn"].to(dev))
    b = msda.ms_deform_attn_forward(x["value"].to(dev), x["shapes"].long().to(dev), x["start"].long().to(dev), x["loc"].to(dev), x["attn"].to(dev))
    assert torch.equal(a, b)


def test_error_behaviour_matches_reference(cuda_device):
    w = SHAPES[0]
    x = {k: v.to(cuda_device) for k, v in torch_inputs(w, seed=19).items()}
    with pytest.raises(RuntimeError, match="value tensor has to be contiguous"):
        msda.ms_deform_attn_forward(x["value"].transpose(2, 3), x["shapes"], x["start"], x["loc"], x["attn"])
    with pytest.raises(RuntimeError, match="sampling_loc must be a CUDA tensor"):
        msda.ms_deform_attn_forward(x["value"], x["shapes"], x["start"], x["loc"].cpu(), x["attn"])
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        torch.ops.alonet_custom.ms_deform_attn_forward(x["value"].cpu(), x["shapes"].cpu(), x["start"].cpu(), x["loc"].cpu(), x["attn"].cpu(), 64)
    with pytest.raises(RuntimeError, match="must divide im2col_step"):
        v3 = torch.cat([x["value"], x["value"][:1]]); l3 = torch.cat([x["loc"], x["loc"][:1]]); a3 = torch.cat([x["attn"], x["attn"][:1]])
        msda.ms_deform_attn_forward(v3, x["shapes"], x["start"], l3, a3, im2col_step=2)


# ---------------------------------------------------------------------------------------------------------
# full-size configurations: size-independent properties
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["C2", "C5DEC", "C5ENC"])
def test_full_size_properties(name, cuda_device):
    w = WORKLOADS[name]
    x = device_inputs(w, seed=5, device=cuda_device, loc_mode="unit")
    fwd = lambda v, a=None: msda.ms_deform_attn_forward(v, x["shapes"], x["start"], x["loc"], x["attn"] if a is None else a)
    out = fwd(x["value"])
    # (1) linearity in value
    v2 = torch.rand_like(x["value"])
    lhs = fwd(2.0 * x["value"] + 0.5 * v2)
    rhs = 2.0 * out + 0.5 * fwd(v2)
    assert torch.allclose(lhs, rhs, rtol=1e-4, atol=1e-6)
    # (2) partition of unity: constant value, locations well inside every level -> out = sum of weights = 1
    loc_in = x["loc"] * 0.8 + 0.1
    ones = torch.ones_like(x["value"])
    o1 = msda.ms_deform_attn_forward(ones, x["shapes"], x["start"], loc_in, x["attn"])
    assert torch.allclose(o1, torch.ones_like(o1), rtol=1e-5, atol=1e-5)
    # (3) adjointness: <out(value), g> == <value, grad_value(g)>  (forward is linear in value)
    gv, gl, ga = msda.ms_deform_attn_backward(x["value"], x["shapes"], x["start"], x["loc"], x["attn"], x["grad_out"])
    a = (out.double() * x["grad_out"].double()).sum()
    b = (x["value"].double() * gv.double()).sum()
    assert abs(a - b) <= 1e-4 * max(abs(a), abs(b)) + 1e-9, (a, b)
    # (4) <attn, grad_attn> == <out, g> as well (forward is linear in the weights)
    c = (x["attn"].double() * ga.double()).sum()
    assert abs(a - c) <= 1e-4 * max(abs(a), abs(c)) + 1e-9, (a, c)
    # (5) batch independence (the sharding property): the first image alone gives the same rows
    o_first = msda.ms_deform_attn_forward(x["value"][:1].contiguous(), x["shapes"], x["start"], x["loc"][:1].contiguous(), x["attn"][:1].contiguous())
    assert torch.equal(o_first, out[:1])


def test_c2_against_oracle_full(cuda_device):
    """BASELINE.json configs[1]/[2] at full size against the C oracle (runs in ~1 s on the host)."""
    w = WORKLOADS["C2"]
    x = torch_inputs(w, seed=23, loc_mode="wide")
    got = run_op(x, cuda_device)
    want = oracle64(x)
    assert_close(got[0], want[0], 1e-4, 1e-7 * rms(want[0]), "out")
    check_grad_value(got[1], {"grad_value": want[1]}, 1e-4)
    assert_close_grad_loc(got[2], want[2], x["loc"].numpy(), x["shapes"].numpy(), 1e-4)
    assert_close_grad(got[3], want[3], 1e-4, "grad_attn")


def test_module_matches_its_tracing_path(cuda_device):
    """MSDeformAttn through the CUDA op == the same module through the traceable pure-PyTorch graph."""
    torch.manual_seed(0)
    dev = cuda_device
    mod = msda.MSDeformAttn(256, 4, 8, 4).to(dev)
    with torch.no_grad():  # make offsets/weights input-dependent (they are zero-initialised)
        mod.sampling_offsets.weight.normal_(0, 0.02)
        mod.attention_weights.weight.normal_(0, 0.1)
    levels = ((16, 20), (8, 10), (4, 5), (2, 3))
    S = sum(h * w for h, w in levels)
    shapes = torch.tensor(levels, dtype=torch.int32, device=dev)
    start = torch.cat((shapes.new_zeros((1,)), (shapes[:, 0] * shapes[:, 1]).cumsum(0)[:-1])).to(torch.int32)
    N, Lq = 2, 50
    q = torch.randn(N, Lq, 256, device=dev)
    src = torch.randn(N, S, 256, device=dev)
    ref2 = torch.rand(N, Lq, 4, 2, device=dev)
    ref4 = torch.cat([ref2, torch.rand(N, Lq, 4, 2, device=dev) * 0.3], -1)
    mask = torch.zeros(N, S, dtype=torch.bool, device=dev)
    mask[:, -7:] = True
    for refp in (ref2, ref4):
        a = mod(q, refp, src, shapes, start, mask)
        b = mod(q, refp, src, shapes, start, mask, is_tracing=None)
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-5), (a - b).abs().max()


def test_more_than_2e31_elements_uses_64bit_indexing(cuda_device):
    """value with > 2^31 elements: the reference indexes with 32-bit `int` (ms_deform_im2col_cuda.cuh:255-271) and
    overflows; here images are addressed with 64-bit offsets.  The LAST image must give the same rows as that image alone."""
    free, _ = torch.cuda.mem_get_info()
    if free < 40e9:
        pytest.skip("needs ~25 GB of device memory")
    w = Workload("big", 384, ((100, 167), (50, 84), (25, 42), (13, 21)), 3, M=8, P=4, D=32)  # 384*22223*256 = 2.18e9 > 2^31
    # (384 = 6 x 64: the reference API demands batch % min(batch, im2col_step) == 0, ms_deform_attn_cuda.cu:50-52)
    assert w.N * w.S * w.M * w.D > 2**31
    dev = cuda_device
    g = torch.Generator(device=dev).manual_seed(1)
    value = torch.rand((w.N, w.S, w.M, w.D), device=dev, generator=g)
    shapes = torch.tensor(w.levels, dtype=torch.int32, device=dev)
    start = torch.cat((shapes.new_zeros((1,)), (shapes[:, 0] * shapes[:, 1]).cumsum(0)[:-1])).to(torch.int32)
    loc = torch.rand((w.N, w.Lq, w.M, w.L, w.P, 2), device=dev, generator=g)
    attn = torch.rand((w.N, w.Lq, w.M, w.L, w.P), device=dev, generator=g)
    go = torch.rand((w.N, w.Lq, w.M * w.D), device=dev, generator=g)
    out = msda.ms_deform_attn_forward(value, shapes, start, loc, attn)
    last = msda.ms_deform_attn_forward(value[-1:].contiguous(), shapes, start, loc[-1:].contiguous(), attn[-1:].contiguous())
    assert torch.equal(out[-1:], last)
    gv, gl, ga = msda.ms_deform_attn_backward(value, shapes, start, loc, attn, go)
    gv1, gl1, ga1 = msda.ms_deform_attn_backward(value[-1:].contiguous(), shapes, start, loc[-1:].contiguous(), attn[-1:].contiguous(), go[-1:].contiguous())
    assert torch.equal(gl[-1:], gl1) and torch.equal(ga[-1:], ga1)
    assert torch.allclose(gv[-1:], gv1, rtol=1e-5, atol=1e-6)   # atomics: order of accumulation may differ
    assert not gv[:-1].isnan().any()


# ------------------------------------------------------------------ early zero-fill (MSDA_BWD_PREZEROED)

@pytest.mark.parametrize("dtype", [torch.float32, torch.float64, torch.bfloat16])
def test_backward_with_early_zero_fill(cuda_device, dtype):
    """begin_backward_zero_fill on a side stream + prezeroed= backward gives the gradients of the in-order path."""
    from aloception_oss_b200.synthetic import Workload, torch_inputs

    w = Workload("ezf", 2, ((20, 27), (10, 14), (5, 7), (3, 4)), 50, M=8, P=4, D=32)
    x = {k: v.to(cuda_device) for k, v in torch_inputs(w, seed=31, loc_mode="wide", dtype=dtype).items()}
    args = (x["value"], x["shapes"], x["start"], x["loc"], x["attn"], x["grad_out"])
    want = msda.ms_deform_attn_backward(*args)
    poison = torch.full((1 << 22,), float("nan"), device=cuda_device)  # the next allocations reuse dirty memory
    del poison
    h = msda.begin_backward_zero_fill(x["value"])
    out = msda.ms_deform_attn_forward(*args[:5])
    got = msda.ms_deform_attn_backward(*args, prezeroed=h)
    torch.cuda.synchronize()
    assert torch.equal(out, msda.ms_deform_attn_forward(*args[:5]))
    tol = dict(rtol=2e-2, atol=1e-4) if dtype == torch.bfloat16 else dict(rtol=1e-4, atol=1e-7)
    for a, b in zip(got, want):
        assert a.dtype == dtype and torch.allclose(a.double(), b.double(), **tol)
    with pytest.raises(RuntimeError, match="different value tensor"):
        msda.ms_deform_attn_backward(*args, prezeroed=msda.begin_backward_zero_fill(x["value"][:1]))


def test_autograd_function_uses_the_early_zero_fill(cuda_device):
    from aloception_oss_b200 import functions
    from aloception_oss_b200.synthetic import Workload, torch_inputs

    w = Workload("ezf2", 2, ((12, 16), (6, 8)), 33, M=8, P=4, D=32)
    x = {k: v.to(cuda_device) for k, v in torch_inputs(w, seed=32, loc_mode="wide").items()}
    grads = {}
    default = functions.EARLY_ZERO_FILL
    for flag in (True, False):
        functions.EARLY_ZERO_FILL = flag
        try:
            v, loc, attn = (x[k].clone().requires_grad_(True) for k in ("value", "loc", "attn"))
            n0 = _capi.kernel_launch_count()
            out = msda.MSDeformAttnFunction.apply(v, x["shapes"], x["start"], loc, attn, 64)
            if flag:
                assert _capi.kernel_launch_count() == n0 + 2  # zero-fill (side stream) + forward
            out.backward(x["grad_out"])
            torch.cuda.synchronize()
            assert _capi.kernel_launch_count() == n0 + 3
            grads[flag] = (v.grad, loc.grad, attn.grad)
        finally:
            functions.EARLY_ZERO_FILL = default
    for a, b in zip(grads[True], grads[False]):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-7)
    # forward-only under no_grad launches nothing extra
    n0 = _capi.kernel_launch_count()
    with torch.no_grad():
        msda.MSDeformAttnFunction.apply(x["value"], x["shapes"], x["start"], x["loc"], x["attn"], 64)
    assert _capi.kernel_launch_count() == n0 + 1


def test_module_traces_to_the_custom_op_node(cuda_device):
    """torch.jit.trace of the module (default ``fused=True``) must record ``alonet_custom::ms_deform_attn_forward`` -- the
    node the reference's TensorRT exporter replaces by its plugin (alonet/torch2trt/trt_exporter.py:41) -- not an opaque
    ctypes call; eager execution keeps the fused kernels."""
    torch.manual_seed(0)
    dev = cuda_device
    mod = msda.MSDeformAttn(256, 4, 8, 4).to(dev).eval()
    levels = ((16, 20), (8, 10), (4, 5), (2, 3))
    S = sum(h * w for h, w in levels)
    shapes = torch.tensor(levels, dtype=torch.int32, device=dev)
    start = torch.cat((shapes.new_zeros((1,)), (shapes[:, 0] * shapes[:, 1]).cumsum(0)[:-1])).to(torch.int32)
    q, src, ref = torch.randn(2, 50, 256, device=dev), torch.randn(2, S, 256, device=dev), torch.rand(2, 50, 4, 2, device=dev)
    class Wrap(torch.nn.Module):  # (tracing a closure over a module cannot embed its parameters)
        def __init__(self):
            super().__init__()
            self.mod = mod

        def forward(self, a, b, c):
            return self.mod(a, b, c, shapes, start)

    with torch.no_grad():
        eager = mod(q, ref, src, shapes, start)
        traced = torch.jit.trace(Wrap(), (q, ref, src), check_trace=False)
        g = str(traced.inlined_graph)
        # torch.jit.trace records a Python autograd Function as prim::PythonOp (the ONNX exporter inlines its forward into the
        # custom-op node); either way the sampling is IN the graph, fed by the traced tensors -- not a baked-in constant
        assert "alonet_custom::ms_deform_attn_forward" in g or "MSDeformAttnFunction" in g, g[-3000:]
        assert "MSDeformAttnFusedFunction" not in g
        assert torch.allclose(traced(q, ref, src), eager, rtol=1e-4, atol=1e-5)
        q2 = torch.randn_like(q)  # the traced graph follows its inputs
        assert torch.allclose(traced(q2, ref, src), mod(q2, ref, src, shapes, start), rtol=1e-4, atol=1e-5)


def test_forward_unroll_knob_is_clamped_for_16bit_rows_of_16_channels(cuda_device):
    """16-bit D = 16: a row is 2 lanes, the warp holds 16 lane groups; ``fwd_unroll = 4`` would address 64 samples per round."""
    w = Workload("d16h", 2, ((7, 9), (3, 5)), 23, M=4, P=4, D=16)
    x = torch_inputs(w, seed=61, loc_mode="wide")
    xr = {k: (v.to(torch.bfloat16).float() if v.is_floating_point() else v) for k, v in x.items()}
    want = oracle64(xr)[0]
    for u in (0, 2, 4):
        _capi.set_tuning("fwd_unroll", u)
        got = run_op(xr, cuda_device, dtype=torch.bfloat16, need_grad=False)[0]
        assert_close(got, want, 1e-2, 1e-2 * rms(want), f"out (fwd_unroll={u})")


def test_module_under_torch_compile_matches_eager(cuda_device):
    """torch.compile (Dynamo + AOT autograd, eager backend) of the mirror module: while the graph is recorded the module
    takes the unfused sequence through ``torch.ops.alonet_custom.*`` (fake tensors via the Meta kernels); forward and
    gradients must equal eager execution."""
    torch.manual_seed(0)
    dev = cuda_device
    mod = msda.MSDeformAttn(256, 4, 8, 4).to(dev)
    with torch.no_grad():
        mod.sampling_offsets.weight.normal_(0, 0.02)
        mod.attention_weights.weight.normal_(0, 0.1)
    levels = ((16, 20), (8, 10), (4, 5), (2, 3))
    S = sum(h * w for h, w in levels)
    shapes = torch.tensor(levels, dtype=torch.int32, device=dev)
    start = torch.cat((shapes.new_zeros((1,)), (shapes[:, 0] * shapes[:, 1]).cumsum(0)[:-1])).to(torch.int32)
    q = torch.randn(2, 50, 256, device=dev, requires_grad=True)
    src = torch.randn(2, S, 256, device=dev, requires_grad=True)
    ref = torch.rand(2, 50, 4, 2, device=dev)
    go = torch.randn(2, 50, 256, device=dev)
    want = mod(q, ref, src, shapes, start)
    gq, gs = torch.autograd.grad(want, (q, src), go)
    try:
        cmod = torch.compile(mod, backend="aot_eager")
        got = cmod(q, ref, src, shapes, start)
        cq, cs = torch.autograd.grad(got, (q, src), go)
    except Exception as e:  # a Dynamo limitation is not an operator bug: report it without failing the suite
        pytest.xfail(f"torch.compile could not trace the module here: {type(e).__name__}: {str(e)[:200]}")
    assert torch.allclose(got, want, rtol=1e-4, atol=1e-5)
    assert torch.allclose(cq, gq, rtol=1e-4, atol=1e-5) and torch.allclose(cs, gs, rtol=1e-4, atol=1e-5)


def test_concurrent_streams_and_threads(cuda_device):
    """The library keeps per-call state (mixed-precision bits, scratch pointer, pre-zeroed flag) in thread-local variables and
    nothing per stream: two Python threads, each on its own CUDA stream, one calling the fp32 operator and the other the
    mixed-precision (bf16 value, fp32 locations) one with the opt-in SM-affine schedule, must both get their single-threaded
    results."""
    import threading

    dev = cuda_device
    w = Workload("mt", 2, ((24, 31), (12, 16), (6, 8), (3, 4)), 1068, M=8, P=4, D=32)
    x = torch_inputs(w, seed=77, loc_mode="wide")
    shapes, start = x["shapes"].to(dev), x["start"].to(dev)
    v32, loc, attn, go = x["value"].to(dev), x["loc"].to(dev), x["attn"].to(dev), x["grad_out"].to(dev)
    v16, go16 = v32.bfloat16(), go.bfloat16()
    want32 = msda.ms_deform_attn_forward(v32, shapes, start, loc, attn)
    want16 = msda.ms_deform_attn_forward(v16, shapes, start, loc, attn)
    wantb = msda.ms_deform_attn_backward(v16, shapes, start, loc, attn, go16)
    torch.cuda.synchronize()
    errors = []

    def worker(kind):
        try:
            stream = torch.cuda.Stream(device=dev)
            with torch.cuda.stream(stream):
                for _ in range(40):
                    if kind == 0:
                        got = msda.ms_deform_attn_forward(v32, shapes, start, loc, attn)
                        ok = torch.equal(got, want32)
                    else:
                        got = msda.ms_deform_attn_forward(v16, shapes, start, loc, attn)
                        gb = msda.ms_deform_attn_backward(v16, shapes, start, loc, attn, go16)
                        ok = torch.equal(got, want16) and torch.equal(gb[1], wantb[1]) and gb[1].dtype == torch.float32
                    stream.synchronize()
                    if not ok:
                        errors.append(kind)
                        return
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    _capi.set_tuning("fwd_pair_mode", 3)
    try:
        ts = [threading.Thread(target=worker, args=(k,)) for k in (0, 1, 0, 1)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
    finally:
        _capi.set_tuning("fwd_pair_mode", 0)
    torch.cuda.synchronize()
    assert not errors, errors
