"""PyTorch restatement of the reference's pure-PyTorch CPU path.

TEST INFRASTRUCTURE ONLY (see oracle/msda_oracle.c).  Used (a) as the differentiable oracle --
autograd through it gives the reference gradients -- and (b) as the ``cpu_baseline`` /
``--impl reference`` arm of bench.py, because the reference's CPU implementation of this path
IS a pure-PyTorch function (there is no C++ CPU kernel: ops/src/cpu/ms_deform_attn_cpu.cpp:17-40
only throws).

Follows alonet/deformable_detr/ops/functions/ms_deform_attn_func.py:
  * ``msda_core_port``       <- ``ms_deform_attn_core_pytorch`` (:85-107): per-level split of value,
    grid = 2*loc-1, per-level bilinear sampling of a (N*M, D, H, W) view, stack over levels, weight by
    attention, sum over L*P, return (N, Lq, M*D).
  * ``_bilinear_zero_pad``   <- ``bilinear_grid_sample`` (:110-190) with align_corners=False:
    x = ((g+1)*W-1)/2, floor, four corner weights, zero padding realised as a 1-pixel zero border,
    clamp of the corner indices into the padded image, four gathers.
Same algorithm and the same tensor-level work (4 gathers per level over (N*M, D, Lq*P)), written
independently; parity with the real reference function is checked in tests/test_oracle_vs_reference.py
(when /root/reference is present) and through tests/golden/.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _bilinear_zero_pad(img: torch.Tensor, grid: torch.Tensor) -> torch.Tensor:
    """img (B, C, H, W), grid (B, Hg, Wg, 2) in [-1, 1] (x, y) -> (B, C, Hg, Wg); align_corners=False."""
    B, C, H, W = img.shape
    _, Hg, Wg, _ = grid.shape
    gx = ((grid[..., 0] + 1) * W - 1) / 2
    gy = ((grid[..., 1] + 1) * H - 1) / 2
    gx = gx.reshape(B, -1)
    gy = gy.reshape(B, -1)
    x_lo = torch.floor(gx).long()
    y_lo = torch.floor(gy).long()
    x_hi = x_lo + 1
    y_hi = y_lo + 1
    w_ll = ((x_hi - gx) * (y_hi - gy)).unsqueeze(1)  # weight of (y_lo, x_lo)
    w_hl = ((x_hi - gx) * (gy - y_lo)).unsqueeze(1)  # weight of (y_hi, x_lo)
    w_lh = ((gx - x_lo) * (y_hi - gy)).unsqueeze(1)  # weight of (y_lo, x_hi)
    w_hh = ((gx - x_lo) * (gy - y_lo)).unsqueeze(1)  # weight of (y_hi, x_hi)

    padded = F.pad(img, (1, 1, 1, 1), mode="constant", value=0.0)  # 1-px zero frame
    PH, PW = H + 2, W + 2
    x_lo = (x_lo + 1).clamp_(0, PW - 1)
    x_hi = (x_hi + 1).clamp_(0, PW - 1)
    y_lo = (y_lo + 1).clamp_(0, PH - 1)
    y_hi = (y_hi + 1).clamp_(0, PH - 1)
    flat = padded.reshape(B, C, PH * PW)

    def take(yy, xx):
        idx = (xx + yy * PW).unsqueeze(1).expand(-1, C, -1)
        return torch.gather(flat, 2, idx)

    res = take(y_lo, x_lo) * w_ll + take(y_hi, x_lo) * w_hl + take(y_lo, x_hi) * w_lh + take(y_hi, x_hi) * w_hh
    return res.reshape(B, C, Hg, Wg)


def msda_core_port(value, spatial_shapes, sampling_locations, attention_weights, use_grid_sample: bool = False):
    """(N,S,M,D), (L,2), (N,Lq,M,L,P,2), (N,Lq,M,L,P) -> (N, Lq, M*D).  Differentiable."""
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = sampling_locations.shape
    hw = [(int(h), int(w)) for h, w in spatial_shapes.tolist()]
    per_level = value.split([h * w for h, w in hw], dim=1)
    grids = 2 * sampling_locations - 1
    sampled = []
    for lvl, (h, w) in enumerate(hw):
        img = per_level[lvl].flatten(2).transpose(1, 2).reshape(N * M, D, h, w)
        g = grids[:, :, :, lvl].transpose(1, 2).flatten(0, 1)  # (N*M, Lq, P, 2)
        if use_grid_sample:
            sampled.append(F.grid_sample(img, g, mode="bilinear", padding_mode="zeros", align_corners=False))
        else:
            sampled.append(_bilinear_zero_pad(img, g))
    aw = attention_weights.transpose(1, 2).reshape(N * M, 1, Lq, L * P)
    out = (torch.stack(sampled, dim=-2).flatten(-2) * aw).sum(-1).view(N, M * D, Lq)
    return out.transpose(1, 2).contiguous()


def msda_fwd_bwd_port(value, spatial_shapes, sampling_locations, attention_weights, grad_output):
    """Forward + autograd backward through the port; returns (out, grad_value, grad_loc, grad_attn)."""
    v = value.detach().clone().requires_grad_(True)
    loc = sampling_locations.detach().clone().requires_grad_(True)
    a = attention_weights.detach().clone().requires_grad_(True)
    out = msda_core_port(v, spatial_shapes, loc, a)
    gv, gl, ga = torch.autograd.grad(out, (v, loc, a), grad_output.reshape_as(out))
    return out.detach(), gv, gl, ga
