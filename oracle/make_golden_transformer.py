"""Golden vectors of the reference's encoder / decoder LAYER LOOP around the operator (run in the build container only).

    python oracle/make_golden_transformer.py

TEST INFRASTRUCTURE.  Imports the UNMODIFIED reference classes ``DeformableTransformerEncoderLayer/Encoder`` and
``DeformableTransformerDecoderLayer/Decoder`` (alonet/deformable_detr/deformable_transformer.py:306-632) through the
namespace stubs of ``aloception_oss_b200.integration.import_reference_ops``, initialises them under a fixed seed (eval
mode, dropout 0), perturbs the zero-initialised attention-logit / offset weights so that they depend on the query, and
evaluates forward (pure-PyTorch operator branch, ``is_tracing``) + autograd backward in float64.  Stored per case: the
state_dict, the inputs, the outputs and the gradients w.r.t. the inputs.  Pins aloception_oss_b200/transformer.py
(SURVEY.md section 8(f) row 3).
"""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from aloception_oss_b200 import integration  # noqa: E402
from aloception_oss_b200.synthetic import level_tensors  # noqa: E402
from oracle import ref_loader  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden_transformer")
LEVELS = ((12, 16), (6, 8), (3, 4), (2, 2))
D_MODEL, D_FFN, HEADS, POINTS, N, LQ = 64, 128, 2, 4, 2, 19


def perturb(module, gen):
    """Give every parameter a query-dependent, non-degenerate value (the reference zero-initialises the offset and
    attention-logit weights)."""
    with torch.no_grad():
        for name, p in module.named_parameters():
            if "sampling_offsets.weight" in name or "attention_weights.weight" in name:
                p.copy_(torch.randn(p.shape, generator=gen, dtype=torch.float64) * 0.05)
            elif "attention_weights.bias" in name:
                p.copy_(torch.randn(p.shape, generator=gen, dtype=torch.float64) * 0.5)
            elif name.endswith("bias") and "sampling_offsets" not in name:
                p.add_(torch.randn(p.shape, generator=gen, dtype=torch.float64) * 0.02)


def main():
    os.makedirs(GOLD, exist_ok=True)
    integration.import_reference_ops(os.path.join(ref_loader.REFERENCE_ROOT, "alonet"))
    ref = importlib.import_module("alonet.deformable_detr.deformable_transformer")
    gen = torch.Generator().manual_seed(1234)
    rnd = lambda *s: torch.randn(*s, generator=gen, dtype=torch.float64)
    shapes_np, start_np = level_tensors(LEVELS)
    shapes, start = torch.from_numpy(shapes_np), torch.from_numpy(start_np)
    S = int(sum(h * w for h, w in LEVELS))
    valid_ratios = 0.75 + 0.25 * torch.rand(N, len(LEVELS), 2, generator=gen, dtype=torch.float64)
    mask = torch.zeros(N, S, dtype=torch.bool)
    mask[1, -7:] = True

    # ---- encoder: 2 layers ----
    torch.manual_seed(7)
    enc = ref.DeformableTransformerEncoder(
        ref.DeformableTransformerEncoderLayer(D_MODEL, D_FFN, 0.0, "relu", len(LEVELS), HEADS, POINTS), 2).double().eval()
    perturb(enc, gen)
    src = rnd(N, S, D_MODEL).requires_grad_(True)
    pos = rnd(N, S, D_MODEL).requires_grad_(True)
    out = enc(src, shapes, start, valid_ratios, pos, mask, is_tracing=None)
    g_out = rnd(*out.shape)
    out.backward(g_out)
    rec = {"src": src.detach(), "pos": pos.detach(), "valid_ratios": valid_ratios, "mask": mask, "out": out.detach(), "grad_out": g_out,
           "g_src": src.grad, "g_pos": pos.grad, "shapes": shapes, "start": start}
    rec.update({"sd_" + k: v for k, v in enc.state_dict().items()})
    save("encoder2", rec)
    memory = out.detach()

    # ---- decoder: 2 layers, intermediates, 2-d reference points; and 4-d reference points with box refinement ----
    for name, ref_dim, refine in (("decoder2_ref2", 2, False), ("decoder2_ref4_refine", 4, True)):
        torch.manual_seed(11)
        dec = ref.DeformableTransformerDecoder(
            ref.DeformableTransformerDecoderLayer(D_MODEL, D_FFN, 0.0, "relu", len(LEVELS), HEADS, POINTS), 2,
            return_intermediate=True).double().eval()
        perturb(dec, gen)
        if refine:
            torch.manual_seed(13)
            dec.bbox_embed = torch.nn.ModuleList([torch.nn.Linear(D_MODEL, 4).double() for _ in range(2)])
        tgt = rnd(N, LQ, D_MODEL).requires_grad_(True)
        query_pos = rnd(N, LQ, D_MODEL).requires_grad_(True)
        mem = memory.clone().requires_grad_(True)
        refpts = torch.rand(N, LQ, ref_dim, generator=gen, dtype=torch.float64) * 0.6 + 0.2
        if ref_dim == 4:
            refpts[..., 2:] = refpts[..., 2:] * 0.3
        res = dec(tgt, refpts, mem, shapes, start, valid_ratios, query_pos, mask, is_tracing=None)
        hs = res["hs"]
        g_hs = rnd(*hs.shape)
        hs.backward(g_hs)
        rec = {"tgt": tgt.detach(), "query_pos": query_pos.detach(), "memory": mem.detach(), "reference_points": refpts,
               "valid_ratios": valid_ratios, "mask": mask, "hs": hs.detach(), "inter_references_out": res["inter_references_out"].detach(),
               "grad_hs": g_hs, "g_tgt": tgt.grad, "g_query_pos": query_pos.grad, "g_memory": mem.grad, "shapes": shapes, "start": start}
        rec.update({"sd_" + k: v for k, v in dec.state_dict().items()})
        save(name, rec)


def save(name, rec):
    arrs = {}
    for k, v in rec.items():
        a = v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)
        arrs[k] = a.astype(np.float32) if a.dtype == np.float64 else a
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"{name}: {len(arrs)} arrays -> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
