#!/usr/bin/env python3
"""Two backward measurements (GPU box), JSON lines:
  * "bwd_chunk_mb": large-batch backward issued whole vs as L2-sized groups of images (C4DEC N = 32 and its N = 8 / 16 cuts);
  * the deterministic mode (MSDA_BWD_DETERMINISTIC) next to the default backward at C2 / C5DEC / ENC.

    python tools/exp_chunk_det.py [--out gpurun_out/chunk_det.jsonl]
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/chunk_det.jsonl")
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--chunk-only", action="store_true")
    args = ap.parse_args()
    tdt = {"f32": torch.float32, "bf16": torch.bfloat16}[args.dtype]
    msda.load_ops()
    dev = torch.device("cuda:0")
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "a") as f:
        def emit(rec):
            print(json.dumps(rec), flush=True)
            f.write(json.dumps(rec) + "\n")

        for n in (32, 16, 8):
            w = WORKLOADS["C4DEC"].with_batch(n)
            sets = [device_inputs(w, seed=5 + i, device=dev, loc_mode="unit", dtype=tdt) for i in range(2)]
            for s in sets:
                s["grads"] = [torch.empty_like(s["value"]), torch.empty_like(s["loc"]), torch.empty_like(s["attn"])]
            bwd = lambda s: msda.ms_deform_attn_backward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"], s["grad_out"], grads=s["grads"])
            rec = dict(experiment="bwd_chunk_mb", workload=f"C4DEC N={n}", dtype=args.dtype, grad_value_mb=round(w.N * w.S * w.M * w.D * 4 / 1e6))
            _capi.set_tuning("bwd_chunk_mb", -1)
            base = [g.clone() for g in bwd(sets[0])]
            rec["whole_us"] = round(time_graph(bwd, sets, n=12), 1)
            for mb in (0, 32, 96):
                _capi.set_tuning("bwd_chunk_mb", mb)
                got = bwd(sets[0])
                torch.cuda.synchronize()
                assert torch.equal(got[1], base[1]) and torch.equal(got[2], base[2])
                assert torch.allclose(got[0].float(), base[0].float(), rtol=1e-4 if args.dtype == "f32" else 2e-2, atol=1e-6 if args.dtype == "f32" else 1e-3)
                rec[f"chunk_{mb or 64}mb_us"] = round(time_graph(bwd, sets, n=12), 1)
            _capi.set_tuning("bwd_chunk_mb", 0)
            emit(rec)
            del sets, base, got
            torch.cuda.empty_cache()

        for name in (() if args.chunk_only else ("C2", "C5DEC", "ENC")):
            w = WORKLOADS[name]
            mode = "raster" if w.Lq == w.S else "unit"
            sets = [device_inputs(w, seed=9 + i, device=dev, loc_mode=mode) for i in range(6 if w.samples < 1e6 else 3)]
            for s in sets:
                s["grads"] = [torch.empty_like(s["value"]), torch.empty_like(s["loc"]), torch.empty_like(s["attn"])]
            n = 100 if w.samples < 1e6 else 12
            t_def = time_graph(lambda s: msda.ms_deform_attn_backward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"], s["grad_out"], grads=s["grads"], deterministic=False), sets, n=n)
            t_det = time_graph(lambda s: msda.ms_deform_attn_backward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"], s["grad_out"], grads=s["grads"], deterministic=True), sets, n=n)
            emit(dict(experiment="deterministic", workload=name, loc=mode, default_us=round(t_def, 2), deterministic_us=round(t_det, 2),
                      slowdown=round(t_det / t_def, 2)))
            del sets
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
