#!/usr/bin/env python3
"""Reference CUDA kernels (rebuilt for sm_100a, oracle/_ref/) vs. this repository's kernels, same inputs, same box.

    python tests/compare_ref_cuda.py [--out gpurun_out/compare_ref.jsonl] [--workloads C2,C5DEC,C4DEC,ENC,C5ENC]

For each workload: checks that both implementations agree (fp32), then times forward and backward of each with
CUDA graphs over rotating input sets.  The reference op is loaded under the torch namespace ``alonet_ref``
(see oracle/build_ref_cuda.py); it is test / bench infrastructure, never part of the product path.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import aloception_oss_b200 as msda
from aloception_oss_b200.synthetic import WORKLOADS, device_inputs
from oracle import build_ref_cuda

L2 = 126 * 1024 * 1024


def time_graph(fn, sets, n=32):
    for i in range(3):
        fn(sets[i % len(sets)])
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(n):
            fn(sets[i % len(sets)])
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
    ap.add_argument("--out", default="gpurun_out/compare_ref.jsonl")
    ap.add_argument("--workloads", default="C2,C5DEC,C4DEC,ENC,C5ENC")
    args = ap.parse_args()
    msda.load_ops()
    ref = build_ref_cuda.load_ops()
    dev = torch.device("cuda:0")
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "a") as f:
        for name in args.workloads.split(","):
            w = WORKLOADS[name]
            mode = "raster" if w.Lq == w.S else "unit"
            sb = w.algorithmic_bytes(4, False) + w.algorithmic_bytes(4, True)
            n_sets = max(2, min(16, int(6 * L2 / sb) + 2))
            sets = [device_inputs(w, seed=5 + i, device=dev, loc_mode=mode) for i in range(n_sets)]
            ours_f = lambda s: msda.ms_deform_attn_forward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"])
            ours_b = lambda s: msda.ms_deform_attn_backward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"], s["grad_out"])
            ref_f = lambda s: ref.ms_deform_attn_forward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"], 64)
            ref_b = lambda s: ref.ms_deform_attn_backward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"], s["grad_out"], 64)
            s0 = sets[0]
            o1, o2 = ours_f(s0), ref_f(s0)
            g1, g2 = ours_b(s0), ref_b(s0)
            errs = {"out": float((o1 - o2).abs().max() / o2.abs().max())}
            for k, a, b in zip(("grad_value", "grad_loc", "grad_attn"), g1, g2):
                errs[k] = float((a - b).abs().max() / b.abs().max())
            # grad_loc away from the floor() discontinuities (tests/_util.near_floor_discontinuity): a sample whose pixel
            # coordinate is within 2e-5 of an integer may legitimately interpolate either neighbouring pixel pair
            wh = s0["shapes"].flip(-1).double().view(1, 1, 1, -1, 1, 2)
            pix = s0["loc"].double() * wh - 0.5
            near = ((pix - pix.round()).abs() < 2e-5).any(-1)
            d = (g1[1] - g2[1]).abs().amax(-1)
            errs["grad_loc_off_discontinuities"] = float(d[~near].max() / g2[1].abs().max())
            errs["samples_on_discontinuities"] = int(near.sum())
            errs["samples"] = int(near.numel())
            rec = dict(workload=name, loc=mode, sets=n_sets, max_rel_to_peak_err=errs,
                       ours_fwd_us=round(time_graph(ours_f, sets), 2), ref_fwd_us=round(time_graph(ref_f, sets), 2),
                       ours_bwd_us=round(time_graph(ours_b, sets), 2), ref_bwd_us=round(time_graph(ref_b, sets), 2))
            rec["fwd_speedup"] = round(rec["ref_fwd_us"] / rec["ours_fwd_us"], 2)
            rec["bwd_speedup"] = round(rec["ref_bwd_us"] / rec["ours_bwd_us"], 2)
            print(json.dumps(rec), flush=True)
            f.write(json.dumps(rec) + "\n")
            del sets
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
