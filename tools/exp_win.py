#!/usr/bin/env python3
"""Experiment (GPU box): windowed forward (msda_fwd_win.cuh, knob fwd_win_mode = 2) against the unit-ordered forward on the
encoder shapes, with pixel-aligned local sampling locations ("raster") and with non-local ones ("unit": uniform over the image).

    python tools/exp_win.py [out.jsonl] [workloads] [loc modes]
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aloception_oss_b200 as msda
from aloception_oss_b200 import _capi
from aloception_oss_b200.synthetic import WORKLOADS, device_inputs
from tools.sweep import time_graph, L2

msda.load_ops()
dev = torch.device("cuda:0")
out = open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/win.jsonl", "a")
names = sys.argv[2].split(",") if len(sys.argv) > 2 else ("ENC", "C5ENC", "C4ENC")
modes = sys.argv[3].split(",") if len(sys.argv) > 3 else ("raster", "unit")
fwd = lambda s: msda.ms_deform_attn_forward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"])
for name in names:
    w = WORKLOADS[name]
    for mode in modes:
        sb = w.algorithmic_bytes(4, False)
        n_sets = max(2, min(12, int(4 * L2 / sb) + 2))
        sets = [device_inputs(w, seed=5 + i, device=dev, loc_mode=mode) for i in range(n_sets)]
        _capi.set_tuning("fwd_win_mode", 1)
        want = fwd(sets[0]).clone()
        t0 = min(time_graph(fwd, sets) for _ in range(3))
        for ctas in (2, 1):
            _capi.set_tuning("fwd_win_mode", 2); _capi.set_tuning("fwd_win_ctas", ctas)
            got = fwd(sets[0]); torch.cuda.synchronize()
            t = min(time_graph(fwd, sets) for _ in range(3))
            rec = dict(workload=name, loc=mode, ctas_per_sm=ctas, bit_equal=bool(torch.equal(got, want)),
                       max_abs_diff=float((got - want).abs().max()), unit_us=round(t0, 2), win_us=round(t, 2), speedup=round(t0 / t, 3))
            print(json.dumps(rec), flush=True); out.write(json.dumps(rec) + "\n"); out.flush()
        del sets; torch.cuda.empty_cache()
for k in ("fwd_win_mode", "fwd_win_ctas"):
    _capi.set_tuning(k, 0)
