"""World-size-2 gloo run of the multi-GPU host logic on CPU: metadata broadcast, batch sharding, output gather.

The operator itself needs a GPU, so each rank evaluates its shard with the CPU oracle port here; what is under test is
that sharding + gather reproduces the unsharded result bit for bit (the batch-independence the N>1 bench relies on)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from aloception_oss_b200 import sharding, synthetic
from oracle.msda_torch_port import msda_core_port


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = synthetic.Workload("g", 5, ((4, 6), (2, 3)), 9, M=2, P=3, D=4)  # 5 images over 2 ranks: uneven shards
        x = synthetic.torch_inputs(w, seed=9)
        shapes = x["shapes"].clone() if rank == 0 else torch.zeros_like(x["shapes"])
        start = x["start"].clone() if rank == 0 else torch.zeros_like(x["start"])
        sharding.broadcast_level_metadata(shapes, start, src=0)
        assert torch.equal(shapes, x["shapes"]) and torch.equal(start, x["start"])
        v, loc, a = sharding.shard_batch(x["value"], x["loc"], x["attn"], rank, world)
        assert v.is_contiguous() and loc.is_contiguous() and a.is_contiguous()
        local = msda_core_port(v, shapes, loc, a)
        full = sharding.gather_outputs(local, w.N)
        ref = msda_core_port(x["value"], x["shapes"], x["loc"], x["attn"])
        ret[rank] = bool(torch.equal(full, ref))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_batch_sharding_world2_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def _sortv_worker(rank, world, port, ret):
    import numpy as np

    from oracle import sortv_oracle

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ok = True
        for b, n in ((5, 37), (1, 101)):  # uneven batch shards; b < world: the polygon axis is split instead
            g = torch.Generator().manual_seed(17)
            v = torch.round(torch.rand(b, n, 24, 2, generator=g) * 4) / 4 - 0.5
            mask = torch.rand(b, n, 24, generator=g) < 0.25
            nv = mask.sum(-1).int()
            ref = torch.from_numpy(sortv_oracle.sort_vertices(v.numpy(), mask.numpy(), nv.numpy()))
            sv, sm, sn, (lo, hi) = sharding.shard_polygons(v, mask, nv, rank, world)
            assert sv.is_contiguous() and sm.is_contiguous() and sn.is_contiguous()
            local = torch.from_numpy(sortv_oracle.sort_vertices(sv.numpy(), sm.numpy(), sn.numpy()))
            if b >= world:
                full = sharding.gather_outputs(local, b)
            else:
                full = sharding.gather_outputs(local[0], b * n).reshape(b, n, 9)
            ok = ok and bool(torch.equal(full, ref))
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_polygon_sharding_world2_gloo():
    """sort_vertices shards by polygon ranges with no collective; the gather reproduces the unsharded indices."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_sortv_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
