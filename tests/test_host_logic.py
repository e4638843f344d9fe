"""Host-side mirror of the reference surface: registration, error behaviour, module contract, sharding (CPU only)."""
import os

import pytest
import torch

import aloception_oss_b200 as msda
from aloception_oss_b200 import integration, sharding, synthetic
from oracle import ref_loader


def _cpu_inputs():
    w = synthetic.Workload("t", 2, ((3, 4), (2, 2)), 5, M=2, P=2, D=4)
    return w, synthetic.torch_inputs(w, seed=1)


def test_ops_registered_with_reference_schemas():
    msda.load_MultiScaleDeformableAttention()
    msda.load_ops()  # idempotent
    fwd = torch.ops.alonet_custom.ms_deform_attn_forward.default._schema
    bwd = torch.ops.alonet_custom.ms_deform_attn_backward.default._schema
    assert [a.name for a in fwd.arguments] == ["value", "spatial_shapes", "level_start_index", "sampling_loc", "attn_weight", "im2col_step"]
    assert len(bwd.arguments) == 7 and str(bwd.returns[0].type) == "List[Tensor]"


def test_cpu_tensors_raise_like_the_reference():
    w, x = _cpu_inputs()
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        msda.MSDeformAttnFunction.apply(x["value"], x["shapes"], x["start"], x["loc"], x["attn"], 64)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        msda.ms_deform_attn_forward(x["value"], x["shapes"], x["start"], x["loc"], x["attn"])
    with pytest.raises(RuntimeError, match="has to be contiguous"):
        msda.ms_deform_attn_forward(x["value"].transpose(0, 1), x["shapes"], x["start"], x["loc"], x["attn"])


def test_meta_kernels_give_reference_shapes():
    w, x = _cpu_inputs()
    m = {k: v.to("meta") for k, v in x.items()}
    out = torch.ops.alonet_custom.ms_deform_attn_forward(m["value"], m["shapes"], m["start"], m["loc"], m["attn"], 64)
    assert out.shape == (w.N, w.Lq, w.M * w.D)
    gv, gl, ga = torch.ops.alonet_custom.ms_deform_attn_backward(m["value"], m["shapes"], m["start"], m["loc"], m["attn"], out, 64)
    assert gv.shape == m["value"].shape and gl.shape == m["loc"].shape and ga.shape == m["attn"].shape


def test_module_contract():
    mod = msda.MSDeformAttn(64, 3, 4, 2)
    assert sorted(mod.state_dict()) == sorted(
        f"{n}.{p}" for n in ("sampling_offsets", "attention_weights", "value_proj", "output_proj") for p in ("weight", "bias"))
    assert mod.im2col_step == 64
    # compass initialisation of the offset bias (reference ms_deform_attn.py:70-82): head 0 points along +x, scaled by point index
    b = mod.sampling_offsets.bias.view(4, 3, 2, 2)
    assert torch.allclose(b[0, :, 0], torch.tensor([1.0, 0.0]).expand(3, 2)) and torch.allclose(b[0, :, 1], torch.tensor([2.0, 0.0]).expand(3, 2))
    with pytest.raises(ValueError):
        msda.MSDeformAttn(65, 3, 4, 2)
    # tracing branch runs on CPU and is differentiable
    levels = ((4, 5), (2, 3), (1, 2))
    S = sum(h * w for h, w in levels)
    shapes = torch.tensor(levels, dtype=torch.int32)
    start = torch.tensor([0, 20, 26], dtype=torch.int32)
    q = torch.randn(2, 7, 64, requires_grad=True)
    out = mod(q, torch.rand(2, 7, 3, 2), torch.randn(2, S, 64), shapes, start, None, is_tracing=None)
    assert out.shape == (2, 7, 64)
    out.sum().backward()
    with pytest.raises(ValueError, match="Last dim of reference_points"):
        mod(q, torch.rand(2, 7, 3, 3), torch.randn(2, S, 64), shapes, start, None, is_tracing=None)


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present")
def test_module_mirror_matches_reference_module_on_the_tracing_path():
    """Same weights, same inputs: our MSDeformAttn mirror == the reference MSDeformAttn (pure-PyTorch branch)."""
    functions, modules = integration.import_reference_ops(os.path.join(ref_loader.REFERENCE_ROOT, "alonet"))
    assert functions.load_MultiScaleDeformableAttention is msda.load_MultiScaleDeformableAttention
    torch.manual_seed(0)
    ref_mod = modules.MSDeformAttn(64, 3, 4, 2)          # the UNMODIFIED reference class; its __init__ calls our loader
    ours = msda.MSDeformAttn(64, 3, 4, 2)
    ours.load_state_dict(ref_mod.state_dict())
    with torch.no_grad():
        for m in (ref_mod, ours):
            m.sampling_offsets.weight.copy_(torch.linspace(-0.02, 0.02, m.sampling_offsets.weight.numel()).view_as(m.sampling_offsets.weight))
            m.attention_weights.weight.copy_(torch.linspace(-0.1, 0.1, m.attention_weights.weight.numel()).view_as(m.attention_weights.weight))
    levels = ((4, 5), (2, 3), (1, 2))
    S = sum(h * w for h, w in levels)
    shapes = torch.tensor(levels, dtype=torch.int32)
    start = torch.tensor([0, 20, 26], dtype=torch.int32)
    q, src = torch.randn(2, 7, 64), torch.randn(2, S, 64)
    mask = torch.zeros(2, S, dtype=torch.bool)
    mask[:, -3:] = True
    for refp in (torch.rand(2, 7, 3, 2), torch.cat([torch.rand(2, 7, 3, 2), torch.rand(2, 7, 3, 2) * 0.2], -1)):
        a = ref_mod(q, refp, src, shapes, start, mask, is_tracing=None)
        b = ours(q, refp, src, shapes, start, mask, is_tracing=None)
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
    # the unmodified reference autograd Function resolves to OUR registered op (CPU tensors -> our CPU kernel raises)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        functions.MSDeformAttnFunction.apply(src.view(2, S, 4, 16), shapes, start, torch.rand(2, 7, 4, 3, 2, 2), torch.rand(2, 7, 4, 3, 2), 64)


def test_shard_ranges_partition_the_batch():
    for n in (0, 1, 7, 32):
        for world in (1, 2, 3, 8):
            ranges = [sharding.shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


def test_workload_accounting_matches_survey():
    c2 = synthetic.WORKLOADS["C2"]
    assert c2.S == 13294 and c2.samples == 76800
    assert c2.algorithmic_bytes(4, False) == 28762160 and c2.algorithmic_bytes(4, True) == 56909872  # 28.76 / 56.91 MB
    assert synthetic.WORKLOADS["C4DEC"].S == 22223
    shapes, start = synthetic.level_tensors(c2.levels)
    assert start.tolist() == [0, 10000, 12500, 13125]


def test_export_graph_is_gather_based_like_the_reference():
    """The ``is_tracing`` branch must trace to split / gather / arithmetic nodes only -- the reference avoids ``grid_sample``
    on purpose (ms_deform_attn_func.py:99-102: its TensorRT tool chain has no GridSample) -- and compute the same numbers as
    ``F.grid_sample`` (kept behind ``use_grid_sample=True``)."""
    import torch

    from aloception_oss_b200.functions import ms_deform_attn_core_pytorch
    from aloception_oss_b200.synthetic import Workload, torch_inputs

    w = Workload("exp", 2, ((6, 4), (3, 2), (1, 5)), 7, M=3, P=3, D=5)
    x = torch_inputs(w, seed=3, loc_mode="wide", dtype=torch.float64)
    shapes = [(int(h), int(wd)) for h, wd in x["shapes"]]
    fn = lambda v, l, a: ms_deform_attn_core_pytorch(v, shapes, l, a)
    traced = torch.jit.trace(fn, (x["value"], x["loc"], x["attn"]))
    g = str(traced.graph)
    assert "grid_sampler" not in g and "aten::gather" in g and "aten::pad" not in g and "constant_pad" not in g
    a = fn(x["value"], x["loc"], x["attn"])
    b = ms_deform_attn_core_pytorch(x["value"], shapes, x["loc"], x["attn"], use_grid_sample=True)
    assert torch.allclose(a, b, rtol=1e-12, atol=1e-15)
    assert torch.allclose(traced(x["value"], x["loc"], x["attn"]), a, rtol=0, atol=0)


def test_level_size_check_is_keyed_on_the_live_tensor_not_on_its_address():
    """A shapes tensor that passed ``sum(H*W) == Len_in`` is remembered, but a NEW tensor is always checked -- also when the
    allocator hands it the address (and version counter) of a freed, validated one."""
    import pytest
    import torch

    from aloception_oss_b200 import modules

    m = modules.MSDeformAttn.__new__(modules.MSDeformAttn)  # the check needs no parameters (and no CUDA library)
    good = torch.tensor([[4, 5], [2, 3]], dtype=torch.int32)
    m._check_level_sizes(good, 26)
    assert modules._VALIDATED_LEVELS[id(good)][0]() is good
    m._check_level_sizes(good, 26)            # cached
    with pytest.raises(AssertionError):
        m._check_level_sizes(good, 27)        # same tensor, other Len_in: checked again
    addr, key = good.data_ptr(), id(good)
    del good
    assert key not in modules._VALIDATED_LEVELS  # evicted with the tensor
    for _ in range(8):  # whatever address the next tensors get (often the same block), inconsistent shapes raise
        bad = torch.tensor([[4, 5], [2, 4]], dtype=torch.int32)
        with pytest.raises(AssertionError):
            m._check_level_sizes(bad, 26)
        del bad
    good2 = torch.tensor([[4, 5], [2, 3]], dtype=torch.int32)
    good2[1, 1] = 4                           # in-place edit bumps the version counter
    with pytest.raises(AssertionError):
        m._check_level_sizes(good2, 26)
