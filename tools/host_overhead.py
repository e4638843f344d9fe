#!/usr/bin/env python3
"""Host-side cost of one operator call (GPU box): wall time per call of the Python API on a tiny problem (the kernel takes
~3 us, so the loop is host-bound), through the plain function, torch.ops, the autograd Function and the fused module path;
plus a cProfile of the plain function."""
import cProfile
import json
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import aloception_oss_b200 as msda
from aloception_oss_b200.synthetic import Workload, device_inputs

msda.load_ops()
dev = torch.device("cuda:0")
w = Workload("tiny", 1, ((16, 16), (8, 8), (4, 4), (2, 2)), 32)
s = device_inputs(w, seed=1, device=dev)


def bench(fn, n=5000):
    for _ in range(200):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    dt = time.perf_counter() - t0
    torch.cuda.synchronize()
    return round(dt / n * 1e6, 2)


res = {}
res["forward_fn_us"] = bench(lambda: msda.ms_deform_attn_forward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"]))
res["forward_torch_op_us"] = bench(lambda: torch.ops.alonet_custom.ms_deform_attn_forward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"], 64))
res["forward_autograd_fn_us"] = bench(lambda: msda.MSDeformAttnFunction.apply(s["value"], s["shapes"], s["start"], s["loc"], s["attn"], 64))
res["backward_fn_us"] = bench(lambda: msda.ms_deform_attn_backward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"], s["grad_out"]))
res["torch_empty_us"] = bench(lambda: torch.empty((1, 32, 256), device=dev))
res["torch_add_us"] = bench(lambda: torch.add(s["attn"], 1.0))
print(json.dumps(res), flush=True)
pr = cProfile.Profile()
pr.enable()
for _ in range(3000):
    msda.ms_deform_attn_forward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"])
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
