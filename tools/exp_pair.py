#!/usr/bin/env python3
"""Experiment (GPU box): paired forward (two heads of a query per warp), static order and SM-affine patch order, against the
unit-ordered forward.  Bit-equality checked on every configuration.

    python tools/exp_pair.py [out.jsonl] [workloads] [modes]
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
out = open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/pair.jsonl", "a")
names = sys.argv[2].split(",") if len(sys.argv) > 2 else ("ENC", "C5ENC", "C4ENC", "C4DEC", "C2")
dtypes = {"f32": torch.float32, "bf16": torch.bfloat16}
dts = sys.argv[3].split(",") if len(sys.argv) > 3 else ("f32",)
fwd = lambda s: msda.ms_deform_attn_forward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"])


def emit(rec):
    print(json.dumps(rec), flush=True)
    out.write(json.dumps(rec) + "\n")
    out.flush()


for name in names:
    for dtn in dts:
        w = WORKLOADS[name]
        mode = "raster" if w.Lq == w.S else "unit"
        sb = w.algorithmic_bytes(4 if dtn == "f32" else 2, False)
        n_sets = max(2, min(12, int(4 * L2 / sb) + 2))
        sets = [device_inputs(w, seed=5 + i, device=dev, loc_mode=mode, dtype=dtypes[dtn]) for i in range(n_sets)]
        _capi.set_tuning("fwd_pair_mode", 1)
        want = fwd(sets[0]).clone()
        t0 = min(time_graph(fwd, sets) for _ in range(3))
        emit(dict(workload=name, dtype=dtn, loc=mode, kernel="unit-ordered", fwd_us=round(t0, 2)))
        _capi.set_tuning("fwd_pair_mode", 2)
        got = fwd(sets[0]); torch.cuda.synchronize()
        t = min(time_graph(fwd, sets) for _ in range(3))
        emit(dict(workload=name, dtype=dtn, loc=mode, kernel="paired static", bit_equal=bool(torch.equal(got, want)),
                  fwd_us=round(t, 2), speedup=round(t0 / t, 3)))
        if mode == "raster":
            for px, py, ctas in ((3, 3, 5), (2, 2, 5), (3, 2, 5), (2, 1, 5)):
                _capi.set_tuning("fwd_pair_mode", 3)
                _capi.set_tuning("fwd_pair_px", px); _capi.set_tuning("fwd_pair_py", py); _capi.set_tuning("fwd_pair_ctas", ctas)
                got = fwd(sets[0]); torch.cuda.synchronize()
                t = min(time_graph(fwd, sets) for _ in range(3))
                emit(dict(workload=name, dtype=dtn, loc=mode, kernel="paired SM-affine", px=1 << px, py=1 << py, ctas_per_sm=ctas,
                          bit_equal=bool(torch.equal(got, want)), fwd_us=round(t, 2), speedup=round(t0 / t, 3)))
        del sets; torch.cuda.empty_cache()
for k in ("fwd_pair_mode", "fwd_pair_px", "fwd_pair_py", "fwd_pair_ctas"):
    _capi.set_tuning(k, 0)
