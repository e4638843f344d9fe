#!/usr/bin/env python3
"""Experiment (GPU box): patch-ordered persistent forward with the speculative gather, encoder shapes."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aloception_oss_b200 as msda
from aloception_oss_b200 import _capi
from aloception_oss_b200.synthetic import WORKLOADS, device_inputs
from tools.sweep import time_graph, L2

msda.load_ops()
dev = torch.device("cuda:0")
out = open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/patch_spec.jsonl", "a")
fwd = lambda s: msda.ms_deform_attn_forward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"])
for name in (sys.argv[2].split(",") if len(sys.argv) > 2 else ("ENC", "C5ENC")):
    w = WORKLOADS[name]
    sb = w.algorithmic_bytes(4, False)
    n_sets = max(2, min(12, int(4 * L2 / sb) + 2))
    sets = [device_inputs(w, seed=5 + i, device=dev, loc_mode="raster") for i in range(n_sets)]
    _capi.set_tuning("patch_mode", 1)
    _capi.set_tuning("spec_mode", 1)
    want = fwd(sets[0]).clone()
    t_flag = min(time_graph(fwd, sets) for _ in range(2))
    _capi.set_tuning("spec_mode", 2)
    t_spec = min(time_graph(fwd, sets) for _ in range(2))
    rec = dict(workload=name, kernel="unit-ordered", flagged_us=round(t_flag, 2), spec_us=round(t_spec, 2))
    print(json.dumps(rec), flush=True); out.write(json.dumps(rec) + "\n")
    for spec in (2, 1):
        for px, py, ctas in ((8, 8, 4), (8, 8, 6), (8, 16, 2), (8, 16, 3), (16, 8, 6), (4, 8, 6), (8, 4, 12), (16, 4, 12), (8, 8, 5)):
            _capi.set_tuning("patch_mode", 2); _capi.set_tuning("spec_mode", spec)
            _capi.set_tuning("patch_px", px); _capi.set_tuning("patch_py", py); _capi.set_tuning("patch_ctas", ctas)
            got = fwd(sets[0]); torch.cuda.synchronize()
            t = min(time_graph(fwd, sets) for _ in range(2))
            rec = dict(workload=name, kernel="patch", spec=spec, px=px, py=py, ctas=ctas, warps_per_sm=py * ctas, bit_equal=bool(torch.equal(got, want)),
                       fwd_us=round(t, 2), vs_unit_spec=round(t_spec / t, 3), gsps=round(w.samples / t / 1e3, 2))
            print(json.dumps(rec), flush=True); out.write(json.dumps(rec) + "\n")
    del sets; torch.cuda.empty_cache()
for k in ("patch_mode", "spec_mode", "patch_px", "patch_py", "patch_ctas"):
    _capi.set_tuning(k, 0)
