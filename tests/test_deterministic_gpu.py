"""MSDA_BWD_DETERMINISTIC: bit-reproducible grad_value through 64-bit fixed-point accumulation (include/msda_b200.h).

The reference scatters grad_value with fp32 atomicAdd (ms_deform_im2col_cuda.cuh:125-152): its result depends on the order in
which the additions land.  The deterministic mode must (1) return the SAME BITS whatever the launch geometry / scheduling,
(2) agree with the oracle to the fp32 tolerance, (3) leave grad_loc / grad_attn untouched, (4) refuse what it does not serve.
"""
import numpy as np
import pytest
import torch

import aloception_oss_b200 as msda
from aloception_oss_b200 import _capi, functions
from aloception_oss_b200.synthetic import WORKLOADS, Workload, device_inputs, torch_inputs
from oracle import msda_oracle
from tests._util import assert_close, check_grad_value, rms

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _ops(cuda_device):
    msda.load_ops()
    for k in ("warps_per_block", "head_major", "bwd_tile_mode"):
        _capi.set_tuning(k, 0)
    yield
    for k in ("warps_per_block", "head_major", "bwd_tile_mode"):
        _capi.set_tuning(k, 0)


def bwd(x, **kw):
    return msda.ms_deform_attn_backward(x["value"], x["shapes"], x["start"], x["loc"], x["attn"], x["grad_out"], **kw)


SHAPES = [
    (Workload("det_d32", 2, ((20, 27), (10, 14), (5, 7), (3, 4)), 300, M=8, P=4, D=32), "unit"),
    (Workload("det_hot", 2, ((3, 4), (2, 2)), 4000, M=8, P=4, D=32), "wide"),   # ~2600 contributions per grad_value element
    (Workload("det_d64", 1, ((7, 9), (3, 5)), 211, M=4, P=4, D=64), "wide"),
    (Workload("det_d16", 2, ((7, 9), (3, 5)), 97, M=4, P=2, D=16), "unit"),
]


@pytest.mark.parametrize("w,mode", SHAPES, ids=[s[0].name for s in SHAPES])
def test_deterministic_backward_is_bit_reproducible_and_correct(w, mode, cuda_device):
    x = {k: v.to(cuda_device) for k, v in torch_inputs(w, seed=51, loc_mode=mode).items()}
    ref = None
    for wpb, hm in ((0, 0), (1, 0), (3, 0), (4, 2), (2, 2)):  # different CTA shapes / unit orders -> different red arrival orders
        _capi.set_tuning("warps_per_block", wpb)
        _capi.set_tuning("head_major", hm)
        for _ in range(2):
            gv, gl, ga = bwd(x, deterministic=True)
            torch.cuda.synchronize()
            if ref is None:
                ref = (gv.clone(), gl.clone(), ga.clone())
            assert torch.equal(gv, ref[0]), "grad_value differs between two deterministic runs"
            assert torch.equal(gl, ref[1]) and torch.equal(ga, ref[2])
    _capi.set_tuning("warps_per_block", 0)
    _capi.set_tuning("head_major", 0)
    n = {k: (v.double().cpu().numpy() if v.is_floating_point() else v.cpu().numpy()) for k, v in x.items()}
    want = msda_oracle.backward(n["grad_out"], n["value"], n["shapes"], n["loc"], n["attn"], n["start"])
    check_grad_value(ref[0].double().cpu().numpy(), {"grad_value": want[0]}, 1e-4)
    base = bwd(x, deterministic=False)
    assert torch.equal(base[1], ref[1]) and torch.equal(base[2], ref[2])  # same kernel arithmetic for grad_loc / grad_attn
    assert_close(base[0].double().cpu().numpy(), ref[0].double().cpu().numpy(), 1e-4, 1e-4 * rms(want[0]), "det vs default")


def test_deterministic_backward_is_at_least_as_accurate_as_the_default(cuda_device):
    """Fixed point with 2^-37 of the largest contribution per addition beats fp32 accumulation on a hot element."""
    w, mode = SHAPES[1]
    x = {k: v.to(cuda_device) for k, v in torch_inputs(w, seed=52, loc_mode=mode).items()}
    n = {k: (v.double().cpu().numpy() if v.is_floating_point() else v.cpu().numpy()) for k, v in x.items()}
    want = msda_oracle.backward(n["grad_out"], n["value"], n["shapes"], n["loc"], n["attn"], n["start"])[0]
    err_det = np.abs(bwd(x, deterministic=True)[0].double().cpu().numpy() - want).max()
    err_def = np.abs(bwd(x, deterministic=False)[0].double().cpu().numpy() - want).max()
    assert err_det <= 1.5 * err_def + 1e-7 * np.abs(want).max(), (err_det, err_def)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
def test_deterministic_backward_16bit(dtype, cuda_device):
    w, mode = SHAPES[0]
    x = {k: v.to(cuda_device) for k, v in torch_inputs(w, seed=53, loc_mode=mode, dtype=dtype).items()}
    a = bwd(x, deterministic=True)
    _capi.set_tuning("warps_per_block", 3)
    b = bwd(x, deterministic=True)
    torch.cuda.synchronize()
    assert all(torch.equal(p, q) for p, q in zip(a, b))
    base = bwd(x, deterministic=False)
    assert torch.allclose(a[0].float(), base[0].float(), rtol=2e-2, atol=1e-2 * base[0].float().abs().max().item())


def test_deterministic_full_size_and_scale_edge_cases(cuda_device):
    """C2 at full size twice; all-zero grad_out; a huge and a tiny grad_out scale (the fixed-point shift follows the data)."""
    x = device_inputs(WORKLOADS["C2"], seed=54, device=cuda_device, loc_mode="unit")
    a = bwd(x, deterministic=True)[0].clone()
    b = bwd(x, deterministic=True)[0]
    assert torch.equal(a, b)
    for scale in (0.0, 1e-30, 3e25):
        xs = dict(x, grad_out=x["grad_out"] * scale)
        gv = bwd(xs, deterministic=True)[0]
        base = bwd(xs, deterministic=False)[0]
        assert torch.isfinite(gv).all()
        if scale == 0.0:
            assert (gv == 0).all()
        else:
            assert torch.allclose(gv, base, rtol=1e-4, atol=1e-5 * base.abs().max().item())


def test_deterministic_mode_follows_torch_and_refuses_what_it_does_not_serve(cuda_device):
    w, mode = SHAPES[0]
    x = {k: v.to(cuda_device) for k, v in torch_inputs(w, seed=55, loc_mode=mode).items()}
    want = bwd(x, deterministic=True)[0]
    torch.use_deterministic_algorithms(True)
    try:
        assert functions.deterministic_requested()
        v, loc, attn = (x[k].clone().requires_grad_(True) for k in ("value", "loc", "attn"))
        n0 = _capi.kernel_launch_count()
        out = msda.MSDeformAttnFunction.apply(v, x["shapes"], x["start"], loc, attn, 64)
        out.backward(x["grad_out"])
        torch.cuda.synchronize()
        assert _capi.kernel_launch_count() == n0 + 5  # forward, absmax, zero-fill, scatter, conversion
        assert torch.equal(v.grad, want)
        # float64 / odd channel counts keep working (with reds) under an IMPLICIT request ...
        x64 = {k: (t.double() if t.is_floating_point() else t) for k, t in x.items()}
        bwd(x64)
    finally:
        torch.use_deterministic_algorithms(False)
    # ... and raise for an explicit one
    with pytest.raises(RuntimeError, match="MSDA_BWD_DETERMINISTIC"):
        bwd(x64, deterministic=True)
    w30 = Workload("d30", 1, ((6, 4), (3, 2)), 2, M=2, P=2, D=30)
    x30 = {k: v.to(cuda_device) for k, v in torch_inputs(w30, seed=56).items()}
    with pytest.raises(RuntimeError, match="MSDA_BWD_DETERMINISTIC"):
        bwd(x30, deterministic=True)
    with pytest.raises(RuntimeError, match="prezeroed"):
        bwd(x, deterministic=True, prezeroed=msda.begin_backward_zero_fill(x["value"]))


def test_module_is_bit_reproducible_under_torch_deterministic_mode(cuda_device):
    """With torch.use_deterministic_algorithms the mirror module leaves its fused path (whose backward uses fp32 atomics) for
    the plain operator with the fixed-point backward: parameter and input gradients are identical run to run; without the
    mode the same module takes the fused kernels."""
    torch.manual_seed(0)
    dev = cuda_device
    mod = msda.MSDeformAttn(256, 4, 8, 4).to(dev)
    with torch.no_grad():
        mod.sampling_offsets.weight.normal_(0, 0.02)
        mod.attention_weights.weight.normal_(0, 0.1)
    levels = ((20, 27), (10, 14), (5, 7), (3, 4))
    S = sum(h * w for h, w in levels)
    shapes = torch.tensor(levels, dtype=torch.int32, device=dev)
    start = torch.cat((shapes.new_zeros((1,)), (shapes[:, 0] * shapes[:, 1]).cumsum(0)[:-1])).to(torch.int32)
    q = torch.randn(2, S, 256, device=dev)
    src = torch.randn(2, S, 256, device=dev)
    ref = torch.rand(2, S, 4, 2, device=dev)
    go = torch.randn(2, S, 256, device=dev)

    def grads():
        src_ = src.clone().requires_grad_(True)
        mod.zero_grad(set_to_none=True)
        out = mod(q, ref, src_, shapes, start)
        out.backward(go)
        return [src_.grad.clone()] + [p.grad.clone() for p in mod.parameters()]

    n0 = _capi.kernel_launch_count()
    base = grads()  # default: fused forward + fused backward (+ zero-fill) = 3 launches
    assert _capi.kernel_launch_count() == n0 + 3
    torch.use_deterministic_algorithms(True, warn_only=True)  # (warn_only: cuBLAS needs CUBLAS_WORKSPACE_CONFIG to be strict)
    try:
        n0 = _capi.kernel_launch_count()
        a = grads()
        assert _capi.kernel_launch_count() == n0 + 5  # forward, absmax, zero-fill, scatter, conversion
        b = grads()
    finally:
        torch.use_deterministic_algorithms(False)
    # the operator's contribution is bit-reproducible: grad of the module input that feeds `value` goes through value_proj's
    # GEMM (cuBLAS, deterministic for fixed shapes and no split-K reductions changes) -- compare everything bit for bit
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    for x, y in zip(a, base):
        assert torch.allclose(x, y, rtol=1e-3, atol=1e-4 * y.abs().max().item())
