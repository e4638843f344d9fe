"""Deterministic synthetic inputs for the multi-scale deformable attention operator.

The recipe follows the reference's op test (alonet/deformable_detr/ops/test.py:38-41):
``value = rand * 0.01``, ``sampling_locations = rand`` (uniform in [0, 1]),
``attention_weights = rand + 1e-5`` normalised over (levels, points).  Two extra location
distributions exercise what the reference test does not: ``"wide"`` draws from [-0.2, 1.2] so a
share of the samples falls outside the level (zero padding, skip window), ``"local"`` puts the
points a few pixels around a per-query reference point, like a freshly initialised
``MSDeformAttn`` does in the encoder (ms_deform_attn.py:70-82).

Inputs come from numpy's PCG64 so that the committed golden fixtures (tests/golden/) can be
re-derived bit-for-bit on any machine; ``device_inputs`` is the fast on-device variant the
benchmark uses for shapes that are too large to ship through the host.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Tuple

import numpy as np

# The COCO pyramids of BASELINE.json's configs ("Sizes").
SQUARE_PYRAMID = [(100, 100), (50, 50), (25, 25), (13, 13)]  # configs[1], configs[2]
COCO_800x1333_PYRAMID = [(100, 167), (50, 84), (25, 42), (13, 21)]  # configs[3], configs[4]


@dataclass(frozen=True)
class Workload:
    """One operator call: batch N, level shapes, Lq queries, M heads, P points, D channels/head."""

    name: str
    N: int
    levels: Tuple[Tuple[int, int], ...]
    Lq: int
    M: int = 8
    P: int = 4
    D: int = 32

    @property
    def L(self) -> int:
        return len(self.levels)

    @property
    def S(self) -> int:
        return int(sum(h * w for h, w in self.levels))

    @property
    def samples(self) -> int:
        """Sampling points (n, q, m, l, p) per call -- the unit of BASELINE.json's metric."""
        return self.N * self.Lq * self.M * self.L * self.P

    def algorithmic_bytes(self, elt: int = 4, backward: bool = False) -> int:
        """every tensor counted once per pass."""
        nsmd = self.N * self.S * self.M * self.D
        nqmlp = self.N * self.Lq * self.M * self.L * self.P
        nqmd = self.N * self.Lq * self.M * self.D
        if backward:
            return elt * (2 * nsmd + 6 * nqmlp + nqmd) + 12 * self.L
        return elt * (nsmd + 3 * nqmlp + nqmd) + 12 * self.L

    def with_batch(self, n: int) -> "Workload":
        return Workload(self.name, n, self.levels, self.Lq, self.M, self.P, self.D)


def _sq(levels):
    return tuple((int(h), int(w)) for h, w in levels)


WORKLOADS = {
    # reference op test shape (ops/test.py:26-29)
    "optest": Workload("optest", 1, _sq([(6, 4), (3, 2)]), 2, M=2, P=2, D=2),
    # BASELINE.json configs[0]
    "C1": Workload("C1", 1, _sq([(64, 64)]), 100),
    # configs[1] / configs[2]: decoder cross-attention shape, square pyramid
    "C2": Workload("C2", 2, _sq(SQUARE_PYRAMID), 300),
    # encoder self-attention shape on the same pyramid (Lq = S)
    "ENC": Workload("ENC", 2, _sq(SQUARE_PYRAMID), 13294),
    # configs[3]: one decoder / encoder call of DeformableDETR-R50 at 800x1333, B=32 on one GPU
    "C4DEC": Workload("C4DEC", 32, _sq(COCO_800x1333_PYRAMID), 300),
    "C4ENC": Workload("C4ENC", 4, _sq(COCO_800x1333_PYRAMID), 22223),
    # configs[4]: per-GPU batch 2 of the training step
    "C5DEC": Workload("C5DEC", 2, _sq(COCO_800x1333_PYRAMID), 300),
    "C5ENC": Workload("C5ENC", 2, _sq(COCO_800x1333_PYRAMID), 22223),
}


def level_tensors(levels):
    """(spatial_shapes int32 (L,2), level_start_index int32 (L,)) as numpy arrays."""
    shapes = np.asarray(levels, dtype=np.int32).reshape(-1, 2)
    hw = shapes[:, 0].astype(np.int64) * shapes[:, 1]
    start = np.concatenate([[0], np.cumsum(hw)[:-1]]).astype(np.int32)
    return shapes, start


def host_inputs(w: Workload, seed: int = 3, loc_mode: str = "unit", dtype=np.float32):
    """numpy inputs: dict(value, shapes, start, loc, attn, grad_out).  Draw order is fixed."""
    rng = np.random.default_rng(seed)
    shapes, start = level_tensors(w.levels)
    value = rng.random((w.N, w.S, w.M, w.D), dtype=np.float32) * np.float32(0.01)
    u = rng.random((w.N, w.Lq, w.M, w.L, w.P, 2), dtype=np.float32)
    if loc_mode == "unit":
        loc = u
    elif loc_mode == "wide":
        loc = u * np.float32(1.4) - np.float32(0.2)
    elif loc_mode == "local":
        ref = rng.random((w.N, w.Lq, 1, 1, 1, 2), dtype=np.float32)
        wh = shapes[:, ::-1].astype(np.float32).reshape(1, 1, 1, w.L, 1, 2)
        loc = ref + (u - np.float32(0.5)) * np.float32(8.0) / wh  # +-4 pixels in each level
    else:
        raise ValueError(loc_mode)
    attn = rng.random((w.N, w.Lq, w.M, w.L, w.P), dtype=np.float32) + np.float32(1e-5)
    attn /= attn.sum(axis=(-1, -2), keepdims=True)
    grad_out = rng.random((w.N, w.Lq, w.M * w.D), dtype=np.float32) - np.float32(0.5)
    return dict(
        value=value.astype(dtype),
        shapes=shapes,
        start=start,
        loc=loc.astype(dtype),
        attn=attn.astype(dtype),
        grad_out=grad_out.astype(dtype),
    )


def torch_inputs(w: Workload, seed: int = 3, loc_mode: str = "unit", dtype=None, device="cpu"):
    """Same draws as ``host_inputs`` as torch tensors on ``device`` (dtype: torch dtype or None=f32)."""
    import torch

    h = host_inputs(w, seed, loc_mode)
    out = {}
    for k, v in h.items():
        t = torch.from_numpy(v)
        if k not in ("shapes", "start") and dtype is not None:
            t = t.to(dtype)
        out[k] = t.to(device)
    return out


def device_inputs(w: Workload, seed: int, device, dtype=None, loc_mode: str = "unit"):
    """Inputs drawn ON the device with torch's generator (not bit-compatible with host_inputs)."""
    import torch

    dtype = dtype or torch.float32
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    shapes_np, start_np = level_tensors(w.levels)
    shapes = torch.from_numpy(shapes_np).to(device)
    start = torch.from_numpy(start_np).to(device)
    value = torch.rand((w.N, w.S, w.M, w.D), device=device, generator=g) * 0.01
    u = torch.rand((w.N, w.Lq, w.M, w.L, w.P, 2), device=device, generator=g)
    if loc_mode == "unit":
        loc = u
    elif loc_mode == "wide":
        loc = u * 1.4 - 0.2
    elif loc_mode == "local":
        ref = torch.rand((w.N, w.Lq, 1, 1, 1, 2), device=device, generator=g)
        wh = shapes.flip(-1).float().view(1, 1, 1, w.L, 1, 2)
        loc = ref + (u - 0.5) * 8.0 / wh
    elif loc_mode == "raster":
        # encoder self-attention: query i IS pixel i of the pyramid; reference point = its centre
        refs = []
        for (h, wd) in w.levels:
            ys, xs = torch.meshgrid(
                (torch.arange(h, device=device) + 0.5) / h, (torch.arange(wd, device=device) + 0.5) / wd, indexing="ij"
            )
            refs.append(torch.stack([xs.reshape(-1), ys.reshape(-1)], -1))
        ref = torch.cat(refs, 0)
        if ref.shape[0] != w.Lq:
            raise ValueError("raster locations need Lq == S")
        wh = shapes.flip(-1).float().view(1, 1, 1, w.L, 1, 2)
        loc = ref.view(1, w.Lq, 1, 1, 1, 2) + (u - 0.5) * 8.0 / wh
    else:
        raise ValueError(loc_mode)
    attn = torch.rand((w.N, w.Lq, w.M, w.L, w.P), device=device, generator=g) + 1e-5
    attn = attn / attn.sum((-1, -2), keepdim=True)
    grad_out = torch.rand((w.N, w.Lq, w.M * w.D), device=device, generator=g) - 0.5
    return dict(
        value=value.to(dtype).contiguous(),
        shapes=shapes,
        start=start,
        loc=loc.to(dtype).contiguous(),
        attn=attn.to(dtype).contiguous(),
        grad_out=grad_out.to(dtype).contiguous(),
    )


# ------------------------------------------------------------------------------------------------------------
# module-level cases (MSDeformAttn.forward incl. projections) -- shared by oracle/make_golden_module.py and the tests
# ------------------------------------------------------------------------------------------------------------
MODULE_CASES = {
    # name: (d_model, n_levels, n_heads, n_points, levels, N, Lq, ref_dim, with_mask)
    "mod_ref2_d256": (256, 4, 8, 4, ((12, 16), (6, 8), (3, 4), (2, 2)), 2, 37, 2, True),
    "mod_ref4_d128": (128, 3, 4, 2, ((9, 5), (4, 3), (2, 1)), 2, 19, 4, False),
}


def module_case(name: str, seed: int = 11):
    """Deterministic weights + inputs of one MSDeformAttn module case, as numpy float32 arrays.

    Returns (cfg dict, state dict, inputs dict).  Weights are drawn here (not taken from ``_reset_parameters``) so that
    offsets and attention logits depend on the query and every parameter receives a non-trivial gradient."""
    d_model, L, M, P, levels, N, Lq, ref_dim, with_mask = MODULE_CASES[name]
    rng = np.random.default_rng(seed)
    S = int(sum(h * w for h, w in levels))

    def nrm(shape, scale):
        return (rng.standard_normal(shape) * scale).astype(np.float32)

    state = {
        "sampling_offsets.weight": nrm((M * L * P * 2, d_model), 0.03),
        "sampling_offsets.bias": nrm((M * L * P * 2,), 1.5),
        "attention_weights.weight": nrm((M * L * P, d_model), 0.1),
        "attention_weights.bias": nrm((M * L * P,), 0.1),
        "value_proj.weight": nrm((d_model, d_model), 0.06),
        "value_proj.bias": nrm((d_model,), 0.01),
        "output_proj.weight": nrm((d_model, d_model), 0.06),
        "output_proj.bias": nrm((d_model,), 0.01),
    }
    shapes, start = level_tensors(levels)
    ref_xy = rng.random((N, Lq, L, 2), dtype=np.float32)
    if ref_dim == 4:
        ref = np.concatenate([ref_xy, rng.random((N, Lq, L, 2), dtype=np.float32) * np.float32(0.4)], -1)
    else:
        ref = ref_xy
    mask = np.zeros((N, S), dtype=bool)
    if with_mask:
        mask[:, -5:] = True
        mask[1, 3:9] = True
    inputs = {
        "query": nrm((N, Lq, d_model), 1.0),
        "reference_points": ref,
        "input_flatten": nrm((N, S, d_model), 1.0),
        "shapes": shapes,
        "start": start,
        "mask": mask if with_mask else None,
        "grad_out": nrm((N, Lq, d_model), 1.0),
    }
    cfg = dict(d_model=d_model, n_levels=L, n_heads=M, n_points=P, levels=levels, N=N, Lq=Lq, ref_dim=ref_dim, S=S)
    return cfg, state, inputs
