#!/usr/bin/env python3
"""Do the red-bound unit-ordered backward and the issue-bound tile-binned backward overlap when they run CONCURRENTLY?
The batch is split: the first `k` images go through the tile kernel on one stream, the rest through the unit-ordered kernel
on another (images are independent).  Compared with either kernel alone on the whole batch.  JSON lines.

    python tools/exp_hybrid.py [--out gpurun_out/hybrid.jsonl]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import aloception_oss_b200 as msda
from aloception_oss_b200 import _capi
from aloception_oss_b200.synthetic import WORKLOADS, device_inputs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/hybrid.jsonl")
    args = ap.parse_args()
    msda.load_ops()
    dev = torch.device("cuda:0")
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    sa, sb = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def part(s, lo, hi):
        p = {k: (v[lo:hi].contiguous() if k not in ("shapes", "start") else v) for k, v in s.items() if k != "grads"}
        p["grads"] = [torch.empty_like(p["value"]), torch.empty_like(p["loc"]), torch.empty_like(p["attn"])]
        return p

    def call(s, mode):
        _capi.set_tuning("bwd_tile_mode", mode)
        return msda.ms_deform_attn_backward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"], s["grad_out"], grads=s["grads"])

    def timed(fn, iters=30):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e3

    with open(args.out, "a") as f:
        for name, mode in (("ENC", "raster"), ("C4ENC", "raster"), ("C4ENC", "unit")):
            w = WORKLOADS[name]
            s = device_inputs(w, seed=3, device=dev, loc_mode=mode)
            s["grads"] = [torch.empty_like(s["value"]), torch.empty_like(s["loc"]), torch.empty_like(s["attn"])]
            rec = dict(workload=name, loc=mode, N=w.N)
            rec["unit_all_us"] = round(timed(lambda: call(s, 1)), 1)
            rec["tile_all_us"] = round(timed(lambda: call(s, 2)), 1)
            for k in range(1, w.N):
                a, b = part(s, 0, k), part(s, k, w.N)

                def both():
                    cur = torch.cuda.current_stream()
                    sa.wait_stream(cur)
                    sb.wait_stream(cur)
                    with torch.cuda.stream(sa):  # tile kernel first: its persistent CTAs take their place, the unit CTAs fill the rest
                        call(a, 2)
                    with torch.cuda.stream(sb):
                        call(b, 1)
                    cur.wait_stream(sa)
                    cur.wait_stream(sb)

                rec[f"tile_{k}_of_{w.N}_concurrent_us"] = round(timed(both), 1)
                rec[f"tile_{k}_alone_us"] = round(timed(lambda: call(a, 2)), 1)
                rec[f"unit_{w.N - k}_alone_us"] = round(timed(lambda: call(b, 1)), 1)
            _capi.set_tuning("bwd_tile_mode", 0)
            print(json.dumps(rec), flush=True)
            f.write(json.dumps(rec) + "\n")
            del s
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
