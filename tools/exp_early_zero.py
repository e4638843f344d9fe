#!/usr/bin/env python3
"""Experiment (GPU box): the backward's zero-fill started on a side stream before the forward kernel
(begin_backward_zero_fill + MSDA_BWD_PREZEROED) vs the in-order zero -> scatter pair; step = forward + backward."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import aloception_oss_b200 as msda
from aloception_oss_b200.synthetic import WORKLOADS, device_inputs
from tools.sweep import time_graph, L2

msda.load_ops()
dev = torch.device("cuda:0")
out = open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/early_zero.jsonl", "a")
for name, dt in (("C2", torch.float32), ("C5DEC", torch.float32), ("C1", torch.float32), ("ENC", torch.float32), ("C4DEC", torch.float32),
                 ("C2", torch.bfloat16)):
    w = WORKLOADS[name]
    mode = "raster" if w.Lq == w.S else "unit"
    elt = 4 if dt == torch.float32 else 2
    sb = w.algorithmic_bytes(elt, False) + w.algorithmic_bytes(elt, True)
    n_sets = max(2, min(24, int(6 * L2 / sb) + 2))
    sets = [device_inputs(w, seed=5 + i, device=dev, dtype=dt, loc_mode=mode) for i in range(n_sets)]

    def fwd(s):
        return msda.ms_deform_attn_forward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"])

    def step_inorder(s):
        return fwd(s), msda.ms_deform_attn_backward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"], s["grad_out"])

    def step_early(s):
        h = msda.begin_backward_zero_fill(s["value"])
        o = fwd(s)
        return o, msda.ms_deform_attn_backward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"], s["grad_out"], prezeroed=h)

    want = step_inorder(sets[0])
    got = step_early(sets[0])
    torch.cuda.synchronize()
    tol = dict(rtol=1e-3, atol=1e-6) if dt == torch.float32 else dict(rtol=2e-2, atol=1e-4)
    ok = torch.equal(want[0], got[0]) and all(torch.allclose(a.float(), b.float(), **tol) for a, b in zip(got[1], want[1]))
    t0 = min(time_graph(step_inorder, sets) for _ in range(2))
    t1 = min(time_graph(step_early, sets) for _ in range(2))
    rec = dict(workload=name, dtype=str(dt).split(".")[-1], ok=bool(ok), step_inorder_us=round(t0, 2), step_early_zero_us=round(t1, 2), speedup=round(t0 / t1, 3))
    print(json.dumps(rec), flush=True)
    out.write(json.dumps(rec) + "\n")
    del sets, want, got
    torch.cuda.empty_cache()
