#!/usr/bin/env python3
"""Run the operator a few times under a profiler (GPU box).

    ncu --set full -k regex:msda_fwd -s 4 -c 2 -o gpurun_out/prof python tools/prof_driver.py --workload C2 --what fwd

Rotates through enough distinct input sets that no launch finds its rows in L2.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import aloception_oss_b200 as msda
from aloception_oss_b200 import _capi
from aloception_oss_b200.synthetic import WORKLOADS, device_inputs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--what", default="both", choices=["fwd", "bwd", "both"])
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--iters", type=int, default=8)
    ap.add_argument("--sets", type=int, default=0)
    ap.add_argument("--loc-mode", default="")
    ap.add_argument("--tune", default="", help="comma list knob=value")
    args = ap.parse_args()
    msda.load_ops()
    for kv in filter(None, args.tune.split(",")):
        k, v = kv.split("=")
        _capi.set_tuning(k, int(v))
    dev = torch.device("cuda:0")
    tdt = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}[args.dtype]
    w = WORKLOADS[args.workload]
    mode = args.loc_mode or ("raster" if w.Lq == w.S else "unit")
    n_sets = args.sets or max(2, min(args.iters, int(400e6 / w.algorithmic_bytes(4, False)) + 1))
    sets = [device_inputs(w, seed=31 + i, device=dev, dtype=tdt, loc_mode=mode) for i in range(n_sets)]
    torch.cuda.synchronize()
    for i in range(args.iters):
        s = sets[i % n_sets]
        if args.what in ("fwd", "both"):
            msda.ms_deform_attn_forward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"])
        if args.what in ("bwd", "both"):
            msda.ms_deform_attn_backward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"], s["grad_out"])
    torch.cuda.synchronize()
    print("done", args.workload, args.what, mode, n_sets)


if __name__ == "__main__":
    main()
