#!/usr/bin/env python3
"""Time the backward with the tile-binned kernel (bwd_tile_mode=2) against the unit-ordered kernel (bwd_tile_mode=1) on the
encoder shapes (GPU box).  JSON lines to --out.

    python tools/exp_bwd_tile.py [--out gpurun_out/bwd_tile.jsonl] [--workloads ENC,C5ENC,C4ENC] [--modes raster,unit]
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
    ap.add_argument("--out", default="gpurun_out/bwd_tile.jsonl")
    ap.add_argument("--workloads", default="ENC,C5ENC,C4ENC")
    ap.add_argument("--modes", default="raster,unit")
    ap.add_argument("--ctas", default="2")
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    msda.load_ops()
    dev = torch.device("cuda:0")
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    bwd = lambda s: msda.ms_deform_attn_backward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"], s["grad_out"])
    with open(args.out, "a") as f:
        for name in args.workloads.split(","):
            w = WORKLOADS[name]
            for mode in args.modes.split(","):
                sb = w.algorithmic_bytes(4, True)
                n_sets = max(2, min(6, int(2 * L2 / sb) + 2))
                sets = [device_inputs(w, seed=5 + i, device=dev, loc_mode=mode) for i in range(n_sets)]
                _capi.set_tuning("bwd_tile_mode", 1)
                t_unit = time_graph(bwd, sets, n=12)
                base = bwd(sets[0])
                rec = dict(tag=args.tag, workload=name, loc=mode, unit_us=round(t_unit, 1))
                for ctas in [int(c) for c in args.ctas.split(",")]:
                    _capi.set_tuning("bwd_tile_mode", 2)
                    _capi.set_tuning("bwd_tile_ctas", ctas)
                    t_tile = time_graph(bwd, sets, n=12)
                    got = bwd(sets[0])
                    torch.cuda.synchronize()
                    errs = [float((a.double() - b.double()).abs().max() / b.double().abs().max()) for a, b in zip(got, base)]
                    rec[f"tile_us_ctas{ctas}"] = round(t_tile, 1)
                    rec[f"speedup_ctas{ctas}"] = round(t_unit / t_tile, 3)
                    rec["max_rel_to_peak"] = [float(f"{e:.2e}") for e in errs]
                _capi.set_tuning("bwd_tile_mode", 0)
                _capi.set_tuning("bwd_tile_ctas", 0)
                print(json.dumps(rec), flush=True)
                f.write(json.dumps(rec) + "\n")
                del sets
                torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
