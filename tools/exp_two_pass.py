#!/usr/bin/env python3
"""Backward: single pass (scatter of a round right behind its gather) vs two passes (all gathers, fence, all scatters) -- knob
"bwd_two_pass" -- with result buffers that rotate with the input sets (nothing L2-resident from the step before).  JSON lines.

    python tools/exp_two_pass.py [--out gpurun_out/two_pass.jsonl] [--workloads C2,C5DEC,C1,C4DEC,ENC]
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
from tools.sweep import time_graph

L2 = 126 * 1024 * 1024


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/two_pass.jsonl")
    ap.add_argument("--workloads", default="C2,C5DEC,C1,C4DEC,ENC")
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    msda.load_ops()
    dev = torch.device("cuda:0")
    tdt = {"f32": torch.float32, "bf16": torch.bfloat16}[args.dtype]
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "a") as f:
        for name in args.workloads.split(","):
            w = WORKLOADS[name]
            mode = "raster" if w.Lq == w.S else "unit"
            sb = w.algorithmic_bytes(4, True)
            n_sets = max(3, min(24, int(8 * L2 / sb) + 2))
            sets = [device_inputs(w, seed=5 + i, device=dev, dtype=tdt, loc_mode=mode) for i in range(n_sets)]
            for s in sets:
                s["grads"] = [torch.empty_like(s["value"]), torch.empty_like(s["loc"]), torch.empty_like(s["attn"])]
            bwd = lambda s: msda.ms_deform_attn_backward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"], s["grad_out"], grads=s["grads"])
            rec = dict(tag=args.tag, workload=name, dtype=args.dtype, loc=mode, sets=n_sets)
            n = 200 if w.samples < 1e6 else 24
            for tp in (1, 2, 1, 2):
                _capi.set_tuning("bwd_two_pass", tp)
                t = time_graph(bwd, sets, n=n)
                key = "single_us" if tp == 1 else "two_pass_us"
                rec[key] = round(min(rec.get(key, 1e9), t), 2)
            _capi.set_tuning("bwd_two_pass", 0)
            rec["speedup"] = round(rec["single_us"] / rec["two_pass_us"], 3)
            print(json.dumps(rec), flush=True)
            f.write(json.dumps(rec) + "\n")
            del sets
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
