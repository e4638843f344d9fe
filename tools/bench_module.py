#!/usr/bin/env python3
"""Module-level timing: MSDeformAttn.forward (+ backward) with the fused operator vs. the reference op sequence.

    python tools/bench_module.py [--out gpurun_out/bench_module.jsonl]

Shapes: encoder self-attention (Lq = S) and decoder cross-attention (Lq = 300) of DeformableDETR-R50 on the 800x1333
pyramid, batch 2 (BASELINE.json configs[4] per-GPU) -- d_model 256, 8 heads, 4 levels, 4 points.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import aloception_oss_b200 as msda
from aloception_oss_b200.synthetic import COCO_800x1333_PYRAMID


def time_graph(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n * 1e3)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/bench_module.jsonl")
    args = ap.parse_args()
    msda.load_ops()
    dev = torch.device("cuda:0")
    levels = COCO_800x1333_PYRAMID
    S = sum(h * w for h, w in levels)
    shapes = torch.tensor(levels, dtype=torch.int32, device=dev)
    start = torch.cat((shapes.new_zeros((1,)), (shapes[:, 0] * shapes[:, 1]).cumsum(0)[:-1])).to(torch.int32)
    N = 2
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "a") as f:
        for what, Lq in (("encoder Lq=S", S), ("decoder Lq=300", 300)):
            torch.manual_seed(0)
            src = torch.randn(N, S, 256, device=dev)
            q = torch.randn(N, Lq, 256, device=dev, requires_grad=True)
            if Lq == S:  # raster reference points, like DeformableTransformerEncoder.get_reference_points
                refs = []
                for (h, w) in levels:
                    ys, xs = torch.meshgrid((torch.arange(h, device=dev) + 0.5) / h, (torch.arange(w, device=dev) + 0.5) / w, indexing="ij")
                    refs.append(torch.stack([xs.reshape(-1), ys.reshape(-1)], -1))
                ref = torch.cat(refs, 0)[None, :, None, :].expand(N, S, 4, 2).contiguous()
            else:
                ref = torch.rand(N, Lq, 4, 2, device=dev)
            src.requires_grad_(True)
            rec = {"shape": what, "N": N, "S": S, "Lq": Lq}
            for fused in (False, True):
                mod = msda.MSDeformAttn(256, 4, 8, 4, fused=fused).to(dev)
                with torch.no_grad():
                    mod.sampling_offsets.weight.normal_(0, 0.01)
                    mod.attention_weights.weight.normal_(0, 0.05)
                go = torch.randn(N, Lq, 256, device=dev)

                def fwd():
                    with torch.no_grad():
                        return mod(q, ref, src, shapes, start, None)

                def fwd_bwd():
                    out = mod(q, ref, src, shapes, start, None)
                    torch.autograd.grad(out, [q, src] + list(mod.parameters()), go)

                key = "fused" if fused else "unfused"
                rec[key + "_fwd_us"] = round(time_graph(fwd), 1)
                rec[key + "_fwd_bwd_us"] = round(time_graph(fwd_bwd), 1)
            rec["fwd_speedup"] = round(rec["unfused_fwd_us"] / rec["fused_fwd_us"], 3)
            rec["fwd_bwd_speedup"] = round(rec["unfused_fwd_bwd_us"] / rec["fused_fwd_bwd_us"], 3)
            print(json.dumps(rec), flush=True)
            f.write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    main()
