#!/usr/bin/env python3
"""Who is right where our fp32 backward and the reference CUDA backward disagree on grad_loc? (GPU box)

Arbiter: the same inputs evaluated in float64 (our generic double kernel, itself pinned to the oracle in the tests).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import aloception_oss_b200 as msda
from aloception_oss_b200.synthetic import WORKLOADS, device_inputs
from oracle import build_ref_cuda

msda.load_ops()
ref = build_ref_cuda.load_ops()
dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "C4DEC"
w = WORKLOADS[name]
s = device_inputs(w, seed=5, device=dev, loc_mode="unit")
ours = msda.ms_deform_attn_backward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"], s["grad_out"])
theirs = ref.ms_deform_attn_backward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"], s["grad_out"], 64)
d = {k: v.double() if v.is_floating_point() else v for k, v in s.items()}
gold = msda.ms_deform_attn_backward(d["value"], d["shapes"], d["start"], d["loc"], d["attn"], d["grad_out"])
for i, nm in enumerate(("grad_value", "grad_loc", "grad_attn")):
    peak = gold[i].abs().max().item()
    eo = (ours[i].double() - gold[i]).abs()
    et = (theirs[i].double() - gold[i]).abs()
    print(f"{nm}: peak {peak:.3e}  ours max err {eo.max().item():.3e}  reference-cuda max err {et.max().item():.3e}")
    if nm == "grad_loc":
        for who, e, t in (("ours", eo, ours[i]), ("ref-cuda", et, theirs[i])):
            idx = torch.unravel_index(e.argmax(), e.shape)
            idx = tuple(int(x) for x in idx)
            b, q, m, l, p, c = idx
            H, W = [int(x) for x in s["shapes"][l].tolist()]
            lx, ly = s["loc"][b, q, m, l, p].tolist()
            print(f"  worst for {who}: idx {idx} got {t[idx].item():.6e} gold {gold[i][idx].item():.6e} "
                  f"loc=({lx!r},{ly!r}) -> x*W-0.5={lx * W - 0.5!r} y*H-0.5={ly * H - 0.5!r} (f32 x {torch.tensor(lx) * W - 0.5})")
        print("  n(|ours-gold| > 1e-3*peak) =", int((eo > 1e-3 * peak).sum()), " n(|ref-gold| > 1e-3*peak) =", int((et > 1e-3 * peak).sum()))

# elements where the two fp32 implementations disagree with each other
gl_o, gl_t, gl_g = ours[1], theirs[1], gold[1]
peak = gl_g.abs().max().item()
bad = ((gl_o - gl_t).abs() > 1e-3 * peak).nonzero()
print("ours vs ref-cuda disagreements on grad_loc:", bad.shape[0])
for row in bad[:8].tolist():
    b, q, m, l, p, c = row
    H, W = [int(x) for x in s["shapes"][l].tolist()]
    lx, ly = s["loc"][b, q, m, l, p].tolist()
    xf = (torch.tensor(lx) * W - 0.5).item()
    yf = (torch.tensor(ly) * H - 0.5).item()
    print(f"  idx {row}: ours {gl_o[tuple(row)].item():.6e} ref {gl_t[tuple(row)].item():.6e} gold {gl_g[tuple(row)].item():.6e} "
          f"H,W=({H},{W}) loc=({lx!r},{ly!r}) f32 x={xf!r} y={yf!r} f64 x={lx * W - 0.5!r} y={ly * H - 0.5!r}")
