"""GPU parity of the sm_100a `sort_vertices` (libsortv_b200.so, through the reference-shaped Python surface) against the CPU
oracle and -- when oracle/_ref/sort_vertices_ref.so travelled to the box -- the reference's own CUDA kernel.  Integer
result: the bar is bit-exact."""
import os

import numpy as np
import pytest
import torch

from aloception_oss_b200 import rotated_iou
from oracle import build_ref_sortv, sortv_oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden_sortv")


def recipe(b, n, m, seed, p_valid=0.2, max_valid=None, snap=False):
    """cuda_op/cuda_ext.py:33-41: uniform vertices around their mean, random mask.  `snap` quantises the coordinates so that
    exact ties, equal vertices and y == 0 (the comparator's undefined corner) occur."""
    g = torch.Generator().manual_seed(seed)
    v = torch.rand(b, n, m, 2, generator=g)
    if snap:
        v = torch.round(v * 4) / 4
    mask = torch.rand(b, n, m, generator=g) < p_valid
    if max_valid is not None:  # the reference kernel writes out of bounds above 8 valid vertices
        order = torch.rand(b, n, m, generator=g).argsort(-1)
        rank = torch.empty_like(order)
        rank.scatter_(-1, order, torch.arange(m).expand(b, n, m))
        keep = torch.zeros_like(mask)
        csum = torch.zeros(b, n, dtype=torch.long)
        for r in range(m):
            sel = (rank == r) & mask
            take = sel.any(-1) & (csum < max_valid)
            keep |= sel & take[..., None]
            csum += take.long()
        mask = keep
    nv = mask.sum(-1).int()
    if snap:
        v = v - 0.5
    else:
        v = v - v.mean(dim=2, keepdim=True)
    return v.contiguous(), mask.contiguous(), nv.contiguous()


def ours(v, m, nv, dev):
    n0 = rotated_iou.kernel_launch_count()
    idx = rotated_iou.sort_v(v.to(dev), m.to(dev), nv.to(dev))
    torch.cuda.synchronize()
    if v.shape[0] * v.shape[1]:
        assert rotated_iou.kernel_launch_count() == n0 + 1
    assert idx.dtype == torch.int32 and tuple(idx.shape) == (v.shape[0], v.shape[1], 9)
    return idx.cpu().numpy()


@pytest.mark.parametrize("name", ["known_answers", "random_pairs"])
def test_golden_fixtures(cuda_device, name):
    r = np.load(os.path.join(GOLD, name + ".npz"))
    idx = ours(torch.from_numpy(r["vertices_norm"]), torch.from_numpy(r["mask"]), torch.from_numpy(r["num_valid"]), cuda_device)
    assert (idx == r["idx_oracle"]).all()
    if "idx_ref_cuda" in r.files:
        assert (idx == r["idx_ref_cuda"]).all()
    # callers either side of the op: sort_indices (mean / normalisation) and the shoelace area
    v = torch.from_numpy(r["vertices"]).to(cuda_device)
    m = torch.from_numpy(r["mask"]).to(cuda_device)
    idx2 = rotated_iou.sort_indices(v, m)
    assert idx2.dtype == torch.int64
    area, sel = rotated_iou.calculate_area(idx2, v)
    assert tuple(sel.shape) == (v.shape[0], v.shape[1], 9, 2)
    assert np.abs(area.cpu().numpy() - r["area_pipeline"]).max() < 1e-5


@pytest.mark.parametrize("b,n,m,seed,snap", [
    (8, 1024, 24, 0, False),      # cuda_ext.py:36-38, the reference's own demo shape
    (8, 1024, 24, 1, True),       # ties / duplicates / y == 0
    (1, 50000, 24, 2, False),     # b = 1: the reference would use ONE CTA
    (3, 777, 24, 3, True),        # ragged tail CTA
    (2, 300, 16, 4, True),        # register path M = 16
    (2, 300, 32, 5, True),        # register path M = 32
    (2, 300, 12, 6, True),        # generic path
    (2, 130, 40, 7, False),       # generic path, m > 32
    (2, 130, 9, 8, True),         # smallest m
    (1, 1, 24, 9, False),
])
def test_against_oracle(cuda_device, b, n, m, seed, snap):
    v, mask, nv = recipe(b, n, m, seed, p_valid=0.25, snap=snap)
    idx = ours(v, mask, nv, cuda_device)
    want = sortv_oracle.sort_vertices(v.numpy(), mask.numpy(), nv.numpy())
    assert (idx == want).all(), f"{(idx != want).any(-1).sum()} polygons differ"


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("b,n,seed,snap,p_valid", [
    (8, 1024, 30, False, 0.2), (8, 1024, 31, True, 0.2), (3, 777, 32, True, 0.3), (1, 128, 33, True, 0.25),
    (2, 4099, 34, True, 0.6),  # most polygons above 8 valid candidates: the tile kernel's direct scan
    (1, 127, 35, True, 0.05), (1, 129, 36, False, 0.9),
])
def test_schedules_return_identical_indices(cuda_device, variant, b, n, seed, snap, p_valid):
    """TMA tile kernel (default) / register kernel / generic kernel (include/sortv_b200.h sortv_set_variant)."""
    v, mask, nv = recipe(b, n, 24, seed, p_valid=p_valid, snap=snap)
    rotated_iou.set_variant(variant)
    try:
        idx = ours(v, mask, nv, cuda_device)
    finally:
        rotated_iou.set_variant(0)
    want = sortv_oracle.sort_vertices(v.numpy(), mask.numpy(), nv.numpy())
    assert (idx == want).all(), f"variant {variant}: {(idx != want).any(-1).sum()} polygons differ"


def test_non_finite_vertices(cuda_device):
    """NaN / Inf coordinates flow through the comparator exactly as in the scalar restatement."""
    v, mask, nv = recipe(2, 512, 24, 40, snap=True)
    g = torch.Generator().manual_seed(41)
    r = torch.rand(v.shape, generator=g)
    v = torch.where(r < 0.02, torch.full_like(v, float("nan")), v)
    v = torch.where((r >= 0.02) & (r < 0.04), torch.full_like(v, float("inf")), v)
    v = torch.where((r >= 0.04) & (r < 0.05), torch.full_like(v, -float("inf")), v)
    for variant in (0, 1, 2, 3, 4):
        rotated_iou.set_variant(variant)
        try:
            idx = ours(v, mask, nv, cuda_device)
        finally:
            rotated_iou.set_variant(0)
        assert (idx == sortv_oracle.sort_vertices(v.numpy(), mask.numpy(), nv.numpy())).all(), variant


def test_unaligned_views_take_the_generic_path(cuda_device):
    v, mask, nv = recipe(1, 257, 24, 12, snap=True)
    vd = torch.zeros(v.numel() + 1, device=cuda_device)[1:].view_as(v).copy_(v)  # 4-byte aligned only
    assert vd.data_ptr() % 16 != 0 and vd.is_contiguous()
    idx = rotated_iou.sort_v(vd, mask.to(cuda_device), nv.to(cuda_device)).cpu().numpy()
    assert (idx == sortv_oracle.sort_vertices(v.numpy(), mask.numpy(), nv.numpy())).all()


def test_empty_and_degenerate(cuda_device):
    v, mask, nv = recipe(0, 5, 24, 0)
    assert ours(v, mask, nv, cuda_device).shape == (0, 5, 9)
    v, mask, nv = recipe(2, 0, 24, 0)
    assert ours(v, mask, nv, cuda_device).shape == (2, 0, 9)
    # num_valid inconsistent with the mask (caller error): follows num_valid like the reference
    v, mask, nv = recipe(2, 64, 24, 3, snap=True, max_valid=8)
    nv2 = torch.clamp(nv + 1, max=8).int()
    idx = ours(v, mask, nv2, cuda_device)
    assert (idx == sortv_oracle.sort_vertices(v.numpy(), mask.numpy(), nv2.numpy())).all()
    # all candidates valid / none valid
    for fill in (True, False):
        mk = torch.full((1, 33, 24), fill)
        nvf = mk.sum(-1).int()
        idx = ours(v[:1, :33].contiguous(), mk, nvf, cuda_device)
        assert (idx == sortv_oracle.sort_vertices(v[:1, :33].numpy(), mk.numpy(), nvf.numpy())).all()


def test_error_surface(cuda_device):
    v, mask, nv = recipe(2, 8, 24, 0)
    vd, md, nd = v.to(cuda_device), mask.to(cuda_device), nv.to(cuda_device)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        rotated_iou.sort_v(vd, mask, nd)
    with pytest.raises(RuntimeError, match="float tensor"):
        rotated_iou.sort_v(vd.double(), md, nd)
    with pytest.raises(RuntimeError, match="bool tensor"):
        rotated_iou.sort_v(vd, md.int(), nd)
    with pytest.raises(RuntimeError, match="int tensor"):
        rotated_iou.sort_v(vd, md, nd.long())
    with pytest.raises(RuntimeError, match="contiguous"):
        rotated_iou.sort_v(vd.transpose(0, 1), md, nd)
    with pytest.raises(RuntimeError, match="candidates"):
        rotated_iou.sort_v(vd[:, :, :8].contiguous(), md[:, :, :8].contiguous(), nd)
    idx = rotated_iou.sort_v(vd.requires_grad_(True), md, nd)
    assert not idx.requires_grad


def test_side_stream(cuda_device):
    v, mask, nv = recipe(4, 512, 24, 21, snap=True)
    s = torch.cuda.Stream(device=cuda_device)
    vd, md, nd = v.to(cuda_device), mask.to(cuda_device), nv.to(cuda_device)
    torch.cuda.synchronize()
    with torch.cuda.stream(s):
        idx = rotated_iou.sort_v(vd, md, nd)
    s.synchronize()
    assert (idx.cpu().numpy() == sortv_oracle.sort_vertices(v.numpy(), mask.numpy(), nv.numpy())).all()


@pytest.mark.skipif(not build_ref_sortv.built(), reason="oracle/_ref/sort_vertices_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("b,n,seed,snap", [(8, 1024, 0, False), (8, 1024, 1, True), (64, 300, 2, True), (1, 4096, 3, False)])
def test_against_the_reference_cuda_kernel(cuda_device, b, n, seed, snap):
    """Bit-exact against the reference's own kernel compiled for sm_100a (<= 8 valid vertices: above that the reference
    writes past its row)."""
    v, mask, nv = recipe(b, n, 24, seed, p_valid=0.3, max_valid=8, snap=snap)
    assert int(nv.max()) <= 8
    idx = ours(v, mask, nv, cuda_device)
    ref = build_ref_sortv.reference_sort_vertices(v.to(cuda_device), mask.to(cuda_device), nv.to(cuda_device)).cpu().numpy()
    assert (idx == ref).all(), f"{(idx != ref).any(-1).sum()} polygons differ from the reference kernel"
    assert (sortv_oracle.sort_vertices(v.numpy(), mask.numpy(), nv.numpy()) == ref).all()
