#!/usr/bin/env python3
"""Layer-loop benchmark (GPU box): 6-layer deformable encoder and 6-layer decoder of DeformableDETR-R50 at the COCO
800x1333 pyramid, inference, random weights.  Eager vs CUDA-graph replay, fp32 vs bf16 autocast,
fused vs unfused operator; the operator's own share is timed separately with the same tensors.

    python tools/bench_transformer.py [--out gpurun_out/transformer.jsonl] [--batch 2]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import aloception_oss_b200 as msda
from aloception_oss_b200 import transformer as T
from aloception_oss_b200.synthetic import COCO_800x1333_PYRAMID, level_tensors

LEVELS = [tuple(x) for x in COCO_800x1333_PYRAMID]


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/transformer.jsonl")
    ap.add_argument("--batch", type=int, default=2)
    args = ap.parse_args()
    msda.load_ops()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    N, d, S, Lq = args.batch, 256, sum(h * w for h, w in LEVELS), 300
    shapes_np, start_np = level_tensors(LEVELS)
    shapes, start = torch.from_numpy(shapes_np).to(dev), torch.from_numpy(start_np).to(dev)
    vr = torch.ones(N, len(LEVELS), 2, device=dev)
    src, pos = torch.randn(N, S, d, device=dev), torch.randn(N, S, d, device=dev)
    tgt, qpos = torch.randn(N, Lq, d, device=dev), torch.randn(N, Lq, d, device=dev)
    refpts = torch.rand(N, Lq, 2, device=dev)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    f = open(args.out, "a")

    def emit(rec):
        print(json.dumps(rec), flush=True)
        f.write(json.dumps(rec) + "\n")

    for fused in (True, False):
        enc = T.DeformableTransformerEncoder(T.DeformableTransformerEncoderLayer(d, 1024, 0.1, "relu", 4, 8, 4, fused=fused), 6).to(dev).eval()
        dec = T.DeformableTransformerDecoder(T.DeformableTransformerDecoderLayer(d, 1024, 0.1, "relu", 4, 8, 4, fused=fused), 6,
                                             return_intermediate=True).to(dev).eval()
        with torch.no_grad():
            for dtype in (torch.float32, torch.bfloat16):
                ctx = lambda: torch.autocast("cuda", dtype=torch.bfloat16, enabled=dtype == torch.bfloat16)
                run_enc = lambda: enc(src, shapes, start, vr, pos, None, spatial_shapes_host=LEVELS)
                with ctx():
                    memory = run_enc()
                run_dec = lambda: dec(tgt, refpts, memory.float(), shapes, start, vr, qpos, None)["hs"]
                with ctx():
                    t_enc = timed(run_enc)
                    t_dec = timed(run_dec)
                    g_enc = T.GraphedModule(enc, src, shapes, start, vr, pos, None, spatial_shapes_host=LEVELS)
                    g_dec = T.GraphedModule(dec, tgt, refpts, memory.float(), shapes, start, vr, qpos, None)
                tg_enc = timed(lambda: g_enc.graph.replay())
                tg_dec = timed(lambda: g_dec.graph.replay())
                emit(dict(N=N, S=S, fused=fused, dtype=str(dtype).split(".")[-1], encoder6_eager_ms=round(t_enc, 3), encoder6_graph_ms=round(tg_enc, 3),
                          decoder6_eager_ms=round(t_dec, 3), decoder6_graph_ms=round(tg_dec, 3)))
                del g_enc, g_dec
    # the operator's share: 6 encoder-shape + 6 decoder-shape fused forward calls on the same shapes
    for dtype in (torch.float32, torch.bfloat16):
        value = torch.randn(N, S, 8, 32, device=dev, dtype=dtype)
        for name, lq in (("encoder", S), ("decoder", Lq)):
            off = torch.randn(N, lq, 8, 4, 4, 2, device=dev, dtype=dtype)
            logit = torch.randn(N, lq, 8, 16, device=dev, dtype=dtype)
            ref = torch.rand(N, lq, 4, 2, device=dev, dtype=dtype)
            t = timed(lambda: msda.ms_deform_attn_fused_forward(value, shapes, start, ref, off, logit), n=50)
            emit(dict(N=N, S=S, op="msda_fused_forward", shape=name, dtype=str(dtype).split(".")[-1], us_per_call=round(t * 1e3, 1), x6_ms=round(6 * t, 3)))


if __name__ == "__main__":
    main()
