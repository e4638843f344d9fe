#!/usr/bin/env python3
"""Experiment (GPU box): zero-fill variants (STG kernel footprint, TMA bulk-store kernel) vs backward time."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import aloception_oss_b200 as msda
from aloception_oss_b200 import _capi
from aloception_oss_b200.synthetic import WORKLOADS, device_inputs
from tools.sweep import time_graph, L2

msda.load_ops()
dev = torch.device("cuda:0")
out = open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/zero.jsonl", "a")
for name in ("C2", "C5DEC", "ENC", "C4DEC"):
    w = WORKLOADS[name]
    mode = "raster" if w.Lq == w.S else "unit"
    sb = w.algorithmic_bytes(4, False) + w.algorithmic_bytes(4, True)
    n_sets = max(2, min(24, int(6 * L2 / sb) + 2))
    sets = [device_inputs(w, seed=5 + i, device=dev, loc_mode=mode) for i in range(n_sets)]
    bwd = lambda s: msda.ms_deform_attn_backward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"], s["grad_out"])
    want = [g.clone() for g in bwd(sets[0])]
    for zm, zc, ck in ((1, 0, 0), (2, 1, 32), (2, 1, 64), (2, 2, 32), (2, 2, 16), (2, 4, 16), (2, 1, 128), (2, 1, 8), (2, 4, 8)):
        _capi.set_tuning("zero_mode", zm)
        _capi.set_tuning("zero_ctas", zc)
        _capi.set_tuning("zero_chunk_kb", ck)
        got = bwd(sets[0])
        torch.cuda.synchronize()
        # grad_value: same reds in a different order -> tiny fp differences allowed; an unzeroed row would be far off
        ok = all(torch.allclose(a, b, rtol=1e-3, atol=1e-6) for a, b in zip(got, want))
        t = min(time_graph(bwd, sets) for _ in range(2))
        rec = dict(workload=name, zero_mode=zm, zero_ctas=zc, chunk_kb=ck, ok=ok, bwd_us=round(t, 2))
        print(json.dumps(rec), flush=True)
        out.write(json.dumps(rec) + "\n")
    for k in ("zero_mode", "zero_ctas", "zero_chunk_kb"):
        _capi.set_tuning(k, 0)
    del sets
    torch.cuda.empty_cache()
