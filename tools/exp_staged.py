#!/usr/bin/env python3
"""Experiment (GPU box): TMA-staged forward (knob staged_mode=2) vs the default forward: bit-equality + timing.

    python tools/exp_staged.py [--out gpurun_out/staged.jsonl] [--workloads ENC,C5ENC] [--warps 32,24,16] [--kb 0]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import aloception_oss_b200 as msda
from aloception_oss_b200 import _capi
from aloception_oss_b200.synthetic import WORKLOADS, Workload, device_inputs
from tools.sweep import time_graph, L2


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/staged.jsonl")
    ap.add_argument("--workloads", default="ENC,C5ENC")
    ap.add_argument("--warps", default="0,24")
    ap.add_argument("--kb", default="0")
    ap.add_argument("--variants", default="0,1")
    ap.add_argument("--dtype", default="f32")
    args = ap.parse_args()
    msda.load_ops()
    dev = torch.device("cuda:0")
    tdt = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}[args.dtype]
    fwd = lambda s: msda.ms_deform_attn_forward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"])
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "a") as f:
        def emit(rec):
            print(json.dumps(rec), flush=True)
            f.write(json.dumps(rec) + "\n")

        # correctness on awkward shapes first: out-of-range samples, ragged L*P, levels that do / do not fit
        for name, w, mode in (
            ("small_wide", Workload("small_wide", 2, ((20, 27), (10, 14), (5, 7), (3, 4)), 700), "wide"),
            ("one_level", Workload("one_level", 1, ((9, 11),), 333), "wide"),
            ("p3", Workload("p3", 2, ((16, 16), (8, 8), (4, 4)), 500, M=8, P=3, D=32), "wide"),
            ("m5", Workload("m5", 1, ((16, 16), (8, 8)), 100, M=5, P=4, D=32), "unit"),
        ):
            s = device_inputs(w, seed=3, device=dev, dtype=tdt, loc_mode=mode)
            _capi.set_tuning("staged_mode", 1)
            want = fwd(s)
            for var in (0, 1):
                for kb in (0, 1, 4):
                    _capi.set_tuning("staged_mode", 2)
                    _capi.set_tuning("staged_variant", var)
                    _capi.set_tuning("staged_kb", kb)
                    got = fwd(s)
                    torch.cuda.synchronize()
                    emit(dict(check=name, variant=var, kb=kb, bit_equal=bool(torch.equal(got, want)), max_abs=float((got - want).abs().max())))
        _capi.set_tuning("staged_kb", 0)
        for name in args.workloads.split(","):
            w = WORKLOADS[name]
            mode = "raster" if w.Lq == w.S else "unit"
            sb = w.algorithmic_bytes(4, False)
            n_sets = max(2, min(8, int(3 * L2 / sb) + 2))
            sets = [device_inputs(w, seed=5 + i, device=dev, dtype=tdt, loc_mode=mode) for i in range(n_sets)]
            _capi.set_tuning("staged_mode", 1)
            want = fwd(sets[0]).clone()
            t0 = time_graph(fwd, sets)
            emit(dict(workload=name, kernel="default", fwd_us=round(t0, 2), gsps=round(w.samples / t0 / 1e3, 2)))
            for var, warps in [(v, wp) for v in [int(x) for x in args.variants.split(",")] for wp in [int(x) for x in args.warps.split(",")]]:
                for kb in [int(x) for x in args.kb.split(",")]:
                    _capi.set_tuning("staged_mode", 2)
                    _capi.set_tuning("staged_variant", var)
                    _capi.set_tuning("staged_warps", warps)
                    _capi.set_tuning("staged_kb", kb)
                    got = fwd(sets[0])
                    torch.cuda.synchronize()
                    eq = bool(torch.equal(got, want))
                    t = time_graph(fwd, sets)
                    emit(dict(workload=name, kernel="staged", variant=var, warps=warps, kb=kb, bit_equal=eq, fwd_us=round(t, 2),
                              gsps=round(w.samples / t / 1e3, 2), speedup=round(t0 / t, 3)))
            _capi.set_tuning("staged_mode", 0)
            del sets
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
