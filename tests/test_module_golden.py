"""MSDeformAttn module against golden vectors of the reference MODULE (tests/golden_module/, produced by
oracle/make_golden_module.py from the unmodified reference class).

CPU: our module mirror on its tracing (pure-PyTorch) branch in float64 -- pins the mirror and the fixtures.
GPU: the same module through the CUDA operator, unfused (reference op sequence) and FUSED (softmax + location arithmetic
in the kernels), in float32.
"""
import os

import numpy as np
import pytest
import torch

import aloception_oss_b200 as msda
from aloception_oss_b200.synthetic import MODULE_CASES, module_case
from tests._util import assert_close, rms

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_module")


def run_module(name, device, dtype, fused, tracing):
    cfg, state, x = module_case(name)
    mod = msda.MSDeformAttn(cfg["d_model"], cfg["n_levels"], cfg["n_heads"], cfg["n_points"], fused=fused)
    mod.load_state_dict({k: torch.from_numpy(v) for k, v in state.items()})
    mod = mod.to(device=device, dtype=dtype)
    t = lambda a: torch.from_numpy(a).to(device=device, dtype=dtype)
    q, ref, src = t(x["query"]).requires_grad_(True), t(x["reference_points"]).requires_grad_(True), t(x["input_flatten"]).requires_grad_(True)
    mask = None if x["mask"] is None else torch.from_numpy(x["mask"]).to(device)
    shapes, start = torch.from_numpy(x["shapes"]).to(device), torch.from_numpy(x["start"]).to(device)
    kw = {"is_tracing": None} if tracing else {}
    out = mod(q, ref, src, shapes, start, mask, **kw)
    out.backward(t(x["grad_out"]))
    res = {"out": out, "g_query": q.grad, "g_ref": ref.grad, "g_src": src.grad}
    for k, p in mod.named_parameters():
        res["gp_" + k] = p.grad
    return {k: v.detach().double().cpu().numpy() for k, v in res.items()}


def compare(got, name, rtol):
    ref = np.load(os.path.join(GOLD, name + ".npz"))
    for k in ref.files:
        if k == "name":
            continue
        want = ref[k].astype(np.float64)
        assert_close(got[k], want, rtol, rtol * rms(want), f"{name}:{k}")


def test_fixtures_present():
    assert sorted(os.path.splitext(f)[0] for f in os.listdir(GOLD)) == sorted(MODULE_CASES)


@pytest.mark.parametrize("name", sorted(MODULE_CASES))
def test_module_mirror_tracing_branch_matches_reference_module(name):
    compare(run_module(name, "cpu", torch.float64, fused=False, tracing=True), name, 2e-6)  # fixtures are stored in fp32


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [False, True], ids=["unfused", "fused"])
@pytest.mark.parametrize("name", sorted(MODULE_CASES))
def test_module_on_gpu_matches_reference_module(name, fused, cuda_device):
    msda.load_ops()
    if fused:
        cfg, state, x = module_case(name)
        dims_ok = cfg["n_levels"] * cfg["n_points"] <= 32
        assert dims_ok
    compare(run_module(name, cuda_device, torch.float32, fused=fused, tracing=False), name, 2e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("ref_dim", [2, 4])
@pytest.mark.parametrize("dtype,rtol", [(torch.float32, 1e-4), (torch.bfloat16, 2e-2)], ids=["f32", "bf16"])
def test_fused_function_equals_unfused_composition(ref_dim, dtype, rtol, cuda_device):
    """MSDeformAttnFusedFunction == softmax + location arithmetic (eager torch) + MSDeformAttnFunction, fwd and all grads,
    on a COCO-shaped call (decoder: Lq=300 on the 800x1333 pyramid)."""
    msda.load_ops()
    dev = cuda_device
    g = torch.Generator(device=dev).manual_seed(5)
    levels = ((100, 167), (50, 84), (25, 42), (13, 21))
    N, Lq, M, D, L, P = 2, 300, 8, 32, 4, 4
    S = sum(h * w for h, w in levels)
    shapes = torch.tensor(levels, dtype=torch.int32, device=dev)
    start = torch.cat((shapes.new_zeros((1,)), (shapes[:, 0] * shapes[:, 1]).cumsum(0)[:-1])).to(torch.int32)
    rnd = lambda *s: torch.rand(*s, device=dev, generator=g)
    value = ((rnd(N, S, M, D) - 0.5)).to(dtype)
    ref = rnd(N, Lq, L, 2)
    if ref_dim == 4:
        ref = torch.cat([ref, rnd(N, Lq, L, 2) * 0.3], -1)
    ref = ref.to(dtype)
    off = ((rnd(N, Lq, M, L, P, 2) - 0.5) * 8).to(dtype)
    logits = ((rnd(N, Lq, M, L * P) - 0.5) * 4).to(dtype)
    go = (rnd(N, Lq, M * D) - 0.5).to(dtype)

    def leaves():
        return [t.detach().clone().requires_grad_(True) for t in (value, ref, off, logits)]

    v1, r1, o1, l1 = leaves()
    out1 = msda.MSDeformAttnFusedFunction.apply(v1, shapes, start, r1, o1, l1)
    out1.backward(go)
    # composition in float32 on the SAME (possibly bf16-rounded) inputs = what the fused kernel computes internally
    v2, r2, o2, l2 = [t.float().detach().requires_grad_(True) for t in (value, ref, off, logits)]
    attn = torch.softmax(l2, -1).view(N, Lq, M, L, P)
    if ref_dim == 2:
        norm = torch.stack([shapes[..., 1], shapes[..., 0]], -1)
        loc = r2[:, :, None, :, None, :] + o2 / norm[None, None, None, :, None, :]
    else:
        loc = r2[:, :, None, :, None, :2] + o2 / P * r2[:, :, None, :, None, 2:] * 0.5
    out2 = msda.MSDeformAttnFunction.apply(v2, shapes, start, loc.contiguous(), attn.contiguous(), 64)
    out2.backward(go.float())
    f = lambda t: t.detach().double().cpu().numpy()
    assert_close(f(out1), f(out2), rtol, rtol * rms(f(out2)), "out")
    for a, b, n in ((v1, v2, "grad_value"), (r1, r2, "grad_ref"), (o1, o2, "grad_offsets"), (l1, l2, "grad_logits")):
        assert_close(f(a.grad), f(b.grad), rtol, rtol * rms(f(b.grad)), n)


@pytest.mark.gpu
@pytest.mark.parametrize("ref_dim", [2, 4])
def test_fused_function_keeps_fp32_reference_points_next_to_bf16_tensors(ref_dim, cuda_device):
    """torch.autocast leaves the reference points in fp32 while value / offsets / logits are bf16.  With MSDA_FUSED_REF_F32
    the kernel starts the location arithmetic from the EXACT points: the result must agree with the fp32 composition on the
    exact points, and be clearly closer to it than what rounding the points to bf16 (1/256 of the image = 0.65 px on the
    167-px level) gives."""
    msda.load_ops()
    dev = cuda_device
    g = torch.Generator(device=dev).manual_seed(6)
    levels = ((100, 167), (50, 84), (25, 42), (13, 21))
    N, Lq, M, D, L, P = 2, 300, 8, 32, 4, 4
    S = sum(h * w for h, w in levels)
    shapes = torch.tensor(levels, dtype=torch.int32, device=dev)
    start = torch.cat((shapes.new_zeros((1,)), (shapes[:, 0] * shapes[:, 1]).cumsum(0)[:-1])).to(torch.int32)
    rnd = lambda *s: torch.rand(*s, device=dev, generator=g)
    value = (rnd(N, S, M, D) - 0.5).to(torch.bfloat16)
    ref = rnd(N, Lq, L, 2) * 0.9 + 0.05
    if ref_dim == 4:
        ref = torch.cat([ref, rnd(N, Lq, L, 2) * 0.3], -1)
    off = ((rnd(N, Lq, M, L, P, 2) - 0.5) * 8).to(torch.bfloat16)
    logits = ((rnd(N, Lq, M, L * P) - 0.5) * 4).to(torch.bfloat16)
    go = (rnd(N, Lq, M * D) - 0.5).to(torch.bfloat16)

    def run(ref_in):
        v, r, o, l = [t.detach().clone().requires_grad_(True) for t in (value, ref_in, off, logits)]
        out = msda.MSDeformAttnFusedFunction.apply(v, shapes, start, r, o, l)
        out.backward(go)
        return out, v.grad, r.grad, o.grad, l.grad

    got = run(ref)                        # fp32 points, bf16 everything else
    old = run(ref.to(torch.bfloat16))     # what the module did before: points rounded to bf16
    assert got[2].dtype == torch.float32 and old[2].dtype == torch.bfloat16
    # fp32 composition on the exact points (value / offsets / logits as stored, i.e. bf16-rounded)
    v2, r2, o2, l2 = [t.float().detach().requires_grad_(True) for t in (value, ref, off, logits)]
    attn = torch.softmax(l2, -1).view(N, Lq, M, L, P)
    if ref_dim == 2:
        norm = torch.stack([shapes[..., 1], shapes[..., 0]], -1)
        loc = r2[:, :, None, :, None, :] + o2 / norm[None, None, None, :, None, :]
    else:
        loc = r2[:, :, None, :, None, :2] + o2 / P * r2[:, :, None, :, None, 2:] * 0.5
    out2 = msda.MSDeformAttnFunction.apply(v2, shapes, start, loc.contiguous(), attn.contiguous(), 64)
    out2.backward(go.float())
    f = lambda t: t.detach().double().cpu().numpy()
    want = [f(out2), f(v2.grad), f(r2.grad), f(o2.grad), f(l2.grad)]
    names = ("out", "grad_value", "grad_ref", "grad_offsets", "grad_logits")
    for a, w, n in zip(got, want, names):
        assert_close(f(a), w, 2e-2, 2e-2 * rms(w), n)
    err_new = np.sqrt(((f(got[0]) - want[0]) ** 2).mean())
    err_old = np.sqrt(((f(old[0]) - want[0]) ** 2).mean())
    assert err_new < 0.25 * err_old, (err_new, err_old)


@pytest.mark.gpu
def test_module_under_autocast_tracks_the_fp32_module(cuda_device):
    """MSDeformAttn under torch.autocast(bf16) (fused path, fp32 reference points kept) vs the same module in fp32."""
    msda.load_ops()
    dev = cuda_device
    torch.manual_seed(0)
    mod = msda.MSDeformAttn(256, 4, 8, 4).to(dev)
    with torch.no_grad():
        mod.sampling_offsets.weight.normal_(0, 0.02)
        mod.attention_weights.weight.normal_(0, 0.1)
    levels = ((100, 167), (50, 84), (25, 42), (13, 21))
    S = sum(h * w for h, w in levels)
    shapes = torch.tensor(levels, dtype=torch.int32, device=dev)
    start = torch.cat((shapes.new_zeros((1,)), (shapes[:, 0] * shapes[:, 1]).cumsum(0)[:-1])).to(torch.int32)
    q, src, ref = torch.randn(2, 300, 256, device=dev), torch.randn(2, S, 256, device=dev), torch.rand(2, 300, 4, 2, device=dev)
    with torch.no_grad():
        want = mod(q, ref, src, shapes, start)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            got = mod(q, ref, src, shapes, start)
            old = mod(q, ref.to(torch.bfloat16), src, shapes, start)
    e_new = (got.float() - want).pow(2).mean().sqrt().item()
    e_old = (old.float() - want).pow(2).mean().sqrt().item()
    scale = want.pow(2).mean().sqrt().item()
    assert e_new < 3e-2 * scale, (e_new, scale)
    assert e_new < e_old, (e_new, e_old)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
def test_unfused_module_under_autocast_keeps_fp32_locations(dtype, cuda_device):
    """fused=False reproduces the reference's op sequence; under autocast the plain operator then gets 16-bit value next to
    fp32 sampling locations / weights (MSDA_LOC_F32 | MSDA_ATTN_F32) and must track the fp32 module as well as the fused path
    does (both read the exact locations)."""
    msda.load_ops()
    dev = cuda_device
    torch.manual_seed(0)
    mod = msda.MSDeformAttn(256, 4, 8, 4, fused=False).to(dev)
    with torch.no_grad():
        mod.sampling_offsets.weight.normal_(0, 0.02)
        mod.attention_weights.weight.normal_(0, 0.1)
    fused = msda.MSDeformAttn(256, 4, 8, 4, fused=True).to(dev)
    fused.load_state_dict(mod.state_dict())
    levels = ((100, 167), (50, 84), (25, 42), (13, 21))
    S = sum(h * w for h, w in levels)
    shapes = torch.tensor(levels, dtype=torch.int32, device=dev)
    start = torch.cat((shapes.new_zeros((1,)), (shapes[:, 0] * shapes[:, 1]).cumsum(0)[:-1])).to(torch.int32)
    q, src, ref = torch.randn(2, 300, 256, device=dev), torch.randn(2, S, 256, device=dev), torch.rand(2, 300, 4, 2, device=dev)

    def run(m, cast):
        m.zero_grad()
        qq = q.clone().requires_grad_(True)
        if cast:
            with torch.autocast("cuda", dtype=dtype):
                out = m(qq, ref, src, shapes, start)
        else:
            out = m(qq, ref, src, shapes, start)
        out.float().square().sum().backward()
        return out.detach().float(), qq.grad, m.sampling_offsets.weight.grad.clone()

    want, got, fus = run(mod, False), run(mod, True), run(fused, True)
    for name, g, f, w in zip(("out", "grad_query", "grad_W_offsets"), got, fus, want):
        scale = w.pow(2).mean().sqrt().item()
        e_g = (g.float() - w).pow(2).mean().sqrt().item()
        e_f = (f.float() - w).pow(2).mean().sqrt().item()
        # (gradients through the piecewise-constant bilinear derivative: see tests/test_ref_module_autocast_gpu.py)
        assert e_g < (5e-2 if name == "out" else 4e-1) * scale, (name, e_g, scale)
        assert e_g < 2.0 * e_f + 1e-3 * scale, (name, e_g, e_f)


def _fused_fuzz_cases(n, seed):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        L = int(rng.integers(1, 5))
        P = int(rng.choice([1, 2, 3, 4, 4, 8]))
        while L * P > 32:
            P //= 2
        levels = tuple((int(rng.integers(2, 31)), int(rng.integers(2, 31))) for _ in range(L))
        out.append((i, int(rng.integers(1, 4)), levels, int(rng.integers(1, 90)), int(rng.choice([1, 2, 5, 8, 8])), P,
                    int(rng.choice([16, 32, 32, 64, 128])), int(rng.choice([2, 4]))))
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("case", _fused_fuzz_cases(32, 7), ids=lambda c: "fz%d_N%d_L%d_Lq%d_M%d_P%d_D%d_rd%d" % (c[0], c[1], len(c[2]), c[3], c[4], c[5], c[6], c[7]))
def test_fused_function_fuzz_against_the_unfused_composition(case, cuda_device):
    """Shape fuzz of the fused operator (softmax + location arithmetic inside the kernels, both reference-point forms) against
    eager softmax / location arithmetic + the plain operator: forward and the four gradients, fp32."""
    msda.load_ops()
    _, N, levels, Lq, M, P, D, ref_dim = case
    dev = cuda_device
    g = torch.Generator(device=dev).manual_seed(100 + case[0])
    L = len(levels)
    S = sum(h * w for h, w in levels)
    shapes = torch.tensor(levels, dtype=torch.int32, device=dev)
    start = torch.cat((shapes.new_zeros((1,)), (shapes[:, 0] * shapes[:, 1]).cumsum(0)[:-1])).to(torch.int32)
    rnd = lambda *s: torch.rand(*s, device=dev, generator=g)
    value = rnd(N, S, M, D) - 0.5
    ref = rnd(N, Lq, L, 2) * 1.2 - 0.1
    if ref_dim == 4:
        ref = torch.cat([ref, rnd(N, Lq, L, 2) * 0.3], -1)
    off = (rnd(N, Lq, M, L, P, 2) - 0.5) * 8
    logits = (rnd(N, Lq, M, L * P) - 0.5) * 4
    go = rnd(N, Lq, M * D) - 0.5
    if not msda.fused_supported(value, shapes, ref, off, logits):
        pytest.skip("shape outside the fused kernels")
    v1, r1, o1, l1 = [t.detach().clone().requires_grad_(True) for t in (value, ref, off, logits)]
    out1 = msda.MSDeformAttnFusedFunction.apply(v1, shapes, start, r1, o1, l1)
    out1.backward(go)
    v2, r2, o2, l2 = [t.detach().clone().requires_grad_(True) for t in (value, ref, off, logits)]
    attn = torch.softmax(l2, -1).view(N, Lq, M, L, P)
    if ref_dim == 2:
        norm = torch.stack([shapes[..., 1], shapes[..., 0]], -1)
        loc = r2[:, :, None, :, None, :] + o2 / norm[None, None, None, :, None, :]
    else:
        loc = r2[:, :, None, :, None, :2] + o2 / P * r2[:, :, None, :, None, 2:] * 0.5
    out2 = msda.MSDeformAttnFunction.apply(v2, shapes, start, loc.contiguous(), attn.contiguous(), 64)
    out2.backward(go)
    f = lambda t: t.detach().double().cpu().numpy()
    assert_close(f(out1), f(out2), 1e-4, 1e-4 * rms(f(out2)), "out")
    for a, b, n in ((v1, v2, "grad_value"), (r1, r2, "grad_ref"), (o1, o2, "grad_offsets"), (l1, l2, "grad_logits")):
        assert_close(f(a.grad), f(b.grad), 1e-4, 2e-4 * max(rms(f(b.grad)), 1e-30), n)
