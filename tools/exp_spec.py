#!/usr/bin/env python3
"""Experiment (GPU box): speculative regular-window forward gather (knob spec_mode: 1 = off, 2 = on): bit-equality on
awkward inputs (out-of-range samples, 1-pixel-wide levels, NaN / Inf in value) and timing on the COCO shapes."""
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
    ap.add_argument("--out", default="gpurun_out/spec.jsonl")
    ap.add_argument("--workloads", default="C2,C5DEC,C4DEC,ENC,C5ENC")
    ap.add_argument("--dtypes", default="f32,bf16")
    args = ap.parse_args()
    msda.load_ops()
    dev = torch.device("cuda:0")
    fwd = lambda s: msda.ms_deform_attn_forward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"])
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    f = open(args.out, "a")

    def emit(rec):
        print(json.dumps(rec), flush=True)
        f.write(json.dumps(rec) + "\n")

    for dname in args.dtypes.split(","):
        tdt = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}[dname]
        for name, w, mode in (
            ("small_wide", Workload("small_wide", 2, ((20, 27), (10, 14), (5, 7), (3, 4)), 700), "wide"),
            ("thin_levels", Workload("thin", 2, ((9, 1), (1, 7), (2, 2), (1, 1)), 300), "wide"),
            ("p17", Workload("p17", 1, ((16, 16), (8, 8), (2, 3)), 200, M=8, P=17, D=32), "wide"),
            ("m5", Workload("m5", 1, ((16, 16), (8, 8)), 100, M=5, P=4, D=32), "wide"),
        ):
            s = device_inputs(w, seed=3, device=dev, dtype=tdt, loc_mode=mode)
            for poison in (None, float("nan"), float("inf")):
                if poison is not None:
                    v = s["value"].clone()
                    v[:, ::29] = poison  # sparse non-finite pixels
                    s = dict(s, value=v)
                _capi.set_tuning("spec_mode", 1)
                want = fwd(s)
                _capi.set_tuning("spec_mode", 2)
                got = fwd(s)
                torch.cuda.synchronize()
                same = bool(torch.equal(torch.nan_to_num(got.float(), nan=12345.0, posinf=23456.0, neginf=-23456.0),
                                        torch.nan_to_num(want.float(), nan=12345.0, posinf=23456.0, neginf=-23456.0)))
                emit(dict(check=name, dtype=dname, poison=str(poison), bit_equal=same, nonfinite_out=int((~torch.isfinite(want.float())).sum())))
        for name in args.workloads.split(","):
            w = WORKLOADS[name]
            mode = "raster" if w.Lq == w.S else "unit"
            sb = w.algorithmic_bytes(4, False)
            n_sets = max(2, min(12, int(4 * L2 / sb) + 2))
            sets = [device_inputs(w, seed=5 + i, device=dev, dtype=tdt, loc_mode=mode) for i in range(n_sets)]
            res = {}
            for sm in (1, 2):
                _capi.set_tuning("spec_mode", sm)
                o = fwd(sets[0]).clone()
                res[sm] = (o, min(time_graph(fwd, sets) for _ in range(2)))
            emit(dict(workload=name, dtype=dname, loc=mode, bit_equal=bool(torch.equal(res[1][0], res[2][0])), flagged_us=round(res[1][1], 2),
                      spec_us=round(res[2][1], 2), speedup=round(res[1][1] / res[2][1], 3), gsps=round(w.samples / res[2][1] / 1e3, 2)))
            del sets
            torch.cuda.empty_cache()
    _capi.set_tuning("spec_mode", 0)


if __name__ == "__main__":
    main()
