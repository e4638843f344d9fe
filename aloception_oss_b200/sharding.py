"""Batch sharding of the operator across the GPUs of one node.

Every ``(n, q, m)`` output depends only on image ``n`` (reference: the kernels index ``value`` by ``b_col`` only,
ms_deform_im2col_cuda.cuh:255-271), so the path shards by contiguous image ranges with NO collective on the data
path.  The only replicated state is the level metadata (``spatial_shapes``, ``level_start_index``; 12 bytes per
level), broadcast once.  One process per GPU, ``torch.distributed`` for the plumbing (NCCL on GPUs, gloo in the CPU
tests).
"""
from __future__ import annotations

from typing import Tuple

import torch


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced ``[begin, end)`` of ``n_items`` images for ``rank`` (earlier ranks take the remainder)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_batch(value, sampling_locations, attention_weights, rank: int, world: int):
    """This rank's image slice of the three per-image tensors (views; call ``.contiguous()`` is not needed: dim 0 slices
    of contiguous tensors are contiguous)."""
    b, e = shard_range(value.shape[0], rank, world)
    return value[b:e], sampling_locations[b:e], attention_weights[b:e]


def shard_polygons(vertices, mask, num_valid, rank: int, world: int):
    """`sort_vertices` (rotated_iou): polygons are independent (sort_vert_kernel.cu:53: one loop iteration per polygon), so the
    (b, n) grid shards like the images above.  Slices dim 0 when b >= world, else flattens to (1, b * n, ...) and slices
    the polygon axis; returns contiguous (b', n', m, 2), (b', n', m), (b', n') for this rank and the ``[begin, end)`` range
    in units of the sliced axis."""
    b, n = vertices.shape[0], vertices.shape[1]
    if b >= world:
        lo, hi = shard_range(b, rank, world)
        return vertices[lo:hi], mask[lo:hi], num_valid[lo:hi], (lo, hi)
    lo, hi = shard_range(b * n, rank, world)
    m = vertices.shape[2]
    return (vertices.reshape(1, b * n, m, 2)[:, lo:hi].contiguous(), mask.reshape(1, b * n, m)[:, lo:hi].contiguous(),
            num_valid.reshape(1, b * n)[:, lo:hi].contiguous(), (lo, hi))


def broadcast_level_metadata(spatial_shapes, level_start_index, src: int = 0, group=None):
    """One broadcast of the replicated metadata from ``src``; returns the (in-place updated) tensors."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(spatial_shapes, src=src, group=group)
        dist.broadcast(level_start_index, src=src, group=group)
    return spatial_shapes, level_start_index


def gather_outputs(local_out, n_total: int, group=None):
    """All-gather of the per-rank output slices back into batch order (utility for evaluation code; NOT on the hot path
    -- the benchmark never calls it)."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    max_n = max(e - b for b, e in sizes)
    pad = local_out.new_zeros((max_n,) + tuple(local_out.shape[1:]))
    pad[: local_out.shape[0]] = local_out
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][: e - b] for r, (b, e) in enumerate(sizes)], 0)
