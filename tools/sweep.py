#!/usr/bin/env python3
"""Time forward / backward under every tuning-knob combination (GPU box).  Writes JSON lines.

    python tools/sweep.py [--out gpurun_out/sweep.jsonl] [--workloads C2,C4DEC,ENC] [--dtype f32]
"""
import argparse
import itertools
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import aloception_oss_b200 as msda
from aloception_oss_b200 import _capi
from aloception_oss_b200.synthetic import WORKLOADS, device_inputs

L2 = 126 * 1024 * 1024


def time_graph(fn, sets, n=48):
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
    ap.add_argument("--out", default="gpurun_out/sweep.jsonl")
    ap.add_argument("--workloads", default="C2,C4DEC,ENC")
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--unrolls", default="1,2,4")
    ap.add_argument("--wpbs", default="2,4,8")
    ap.add_argument("--no-pdl", default="0")
    ap.add_argument("--head-major", default="1")
    ap.add_argument("--smem-records", default="0")
    ap.add_argument("--patch", default="1:0:0:0", help="comma list of patch_mode:px:py:ctas")
    ap.add_argument("--tag", default="")
    ap.add_argument("--no-generic", action="store_true")
    ap.add_argument("--max-sets", type=int, default=24)
    args = ap.parse_args()
    msda.load_ops()
    dev = torch.device("cuda:0")
    tdt = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}[args.dtype]
    elt = 4 if args.dtype == "f32" else 2
    peak = 6533.8
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "a") as f:
        for name in args.workloads.split(","):
            w = WORKLOADS[name]
            mode = "raster" if w.Lq == w.S else "unit"
            sb = w.algorithmic_bytes(elt, False) + w.algorithmic_bytes(elt, True)
            n_sets = max(2, min(args.max_sets, int(6 * L2 / sb) + 2))  # touched rows are sparse: over-provision
            sets = [device_inputs(w, seed=5 + i, device=dev, dtype=tdt, loc_mode=mode) for i in range(n_sets)]
            fwd = lambda s: msda.ms_deform_attn_forward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"])
            bwd = lambda s: msda.ms_deform_attn_backward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"], s["grad_out"])
            for pm, sr, hm, var, u, wpb in itertools.product(args.patch.split(","), [int(x) for x in args.smem_records.split(",")], [int(x) for x in args.head_major.split(",")], [int(x) for x in args.no_pdl.split(",")], [int(x) for x in args.unrolls.split(",")], [int(x) for x in args.wpbs.split(",")]):
                for k, v in zip(("patch_mode", "patch_px", "patch_py", "patch_ctas"), pm.split(":")):
                    _capi.set_tuning(k, int(v))
                _capi.set_tuning("smem_records", sr)
                _capi.set_tuning("head_major", hm)
                _capi.set_tuning("no_pdl", var)
                _capi.set_tuning("warps_per_block", wpb)
                _capi.set_tuning("fwd_unroll", u)
                _capi.set_tuning("bwd_unroll", u)
                tf = time_graph(fwd, sets)
                tb = time_graph(bwd, sets)
                rec = dict(tag=args.tag, patch=pm, workload=name, dtype=args.dtype, loc=mode, sets=n_sets, no_pdl=var, head_major=hm, smem_records=sr, unroll=u, wpb=wpb, fwd_us=round(tf, 2), bwd_us=round(tb, 2),
                           fwd_frac=round(w.algorithmic_bytes(elt, False) / tf / 1e3 / peak, 4),
                           bwd_frac=round(w.algorithmic_bytes(elt, True) / tb / 1e3 / peak, 4),
                           fwd_gsps=round(w.samples / tf / 1e3, 3), bwd_gsps=round(w.samples / tb / 1e3, 3))
                print(json.dumps(rec), flush=True)
                f.write(json.dumps(rec) + "\n")
            _capi.set_tuning("head_major", 0)
            _capi.set_tuning("smem_records", 0)
            _capi.set_tuning("patch_mode", 0)
            if args.no_generic:
                del sets
                torch.cuda.empty_cache()
                continue
            _capi.set_tuning("force_generic", 1)
            tf = time_graph(fwd, sets)
            tb = time_graph(bwd, sets)
            _capi.set_tuning("force_generic", 0)
            rec = dict(workload=name, dtype=args.dtype, loc=mode, kernel="generic", fwd_us=round(tf, 2), bwd_us=round(tb, 2))
            print(json.dumps(rec), flush=True)
            f.write(json.dumps(rec) + "\n")
            del sets
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
