#!/usr/bin/env python3
"""Experiment (GPU box): forward / backward time of the plain operator per storage type -- fp32, bf16, and the mixed entry (bf16 value /
grad_output, fp32 sampling_loc / attn_weight: what the reference module produces under torch.autocast) -- on the BASELINE shapes.

    python tools/exp_dtypes.py [out.jsonl] [workloads]
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aloception_oss_b200 as msda
from aloception_oss_b200.synthetic import WORKLOADS, device_inputs
from tools.sweep import time_graph, L2

msda.load_ops()
dev = torch.device("cuda:0")
out = open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/dtypes.jsonl", "a")
names = sys.argv[2].split(",") if len(sys.argv) > 2 else ("C2", "C5DEC", "C4DEC", "ENC", "C5ENC", "C4ENC")
for name in names:
    w = WORKLOADS[name]
    mode = "raster" if w.Lq == w.S else "unit"
    n_sets = max(2, min(10, int(3 * L2 / w.algorithmic_bytes(4, False)) + 2))
    base = [device_inputs(w, seed=5 + i, device=dev, loc_mode=mode) for i in range(n_sets)]
    for kind in ("f32", "bf16", "bf16 value + f32 loc/attn"):
        def conv(s):
            if kind == "f32":
                return s
            r = dict(s)
            for k in ("value", "grad_out"):
                r[k] = s[k].bfloat16()
            if kind == "bf16":
                for k in ("loc", "attn"):
                    r[k] = s[k].bfloat16()
            return r
        sets = [conv(s) for s in base]
        fwd = lambda s: msda.ms_deform_attn_forward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"])
        bwd = lambda s: msda.ms_deform_attn_backward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"], s["grad_out"])
        tf = min(time_graph(fwd, sets, n=24) for _ in range(2))
        tb = min(time_graph(bwd, sets, n=24) for _ in range(2))
        rec = dict(workload=name, loc=mode, storage=kind, fwd_us=round(tf, 2), bwd_us=round(tb, 2),
                   fwd_gsamples_s=round(w.samples / tf / 1e3, 2), bwd_gsamples_s=round(w.samples / tb / 1e3, 2))
        print(json.dumps(rec), flush=True)
        out.write(json.dumps(rec) + "\n"); out.flush()
        del sets
    del base
    torch.cuda.empty_cache()
