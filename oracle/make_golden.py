"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

    python oracle/make_golden.py

TEST INFRASTRUCTURE ONLY.  Imports ``ms_deform_attn_core_pytorch`` from /root/reference through
oracle/ref_loader.py, evaluates it (and autograd through it) on the deterministic inputs of
``aloception_oss_b200.synthetic`` and stores ONLY the reference outputs plus the recipe (workload
name / dims, seed, location mode).  Inputs are re-derived from the recipe by the tests, so the fixtures
stay small.  The reference holds no stored vectors for this path; the recipe of
its op test (ops/test.py:26-41: dims, seed 3, value*0.01, normalised weights, D in {30,32,64,71}) is
what these cases follow.

For the large cases grad_value (N,S,M,D) is stored as two projections (sum over channels -> (N,S,M),
sum over pixels -> (N,M,D)) plus 4096 randomly chosen entries, which pins the scatter without a
multi-MB fixture.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from aloception_oss_b200.synthetic import WORKLOADS, Workload, torch_inputs  # noqa: E402
from oracle import ref_loader  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# (case name, workload, seed, loc_mode, reference dtype, store grad_value in full?)
CASES = []
for D in (2, 30, 32, 64, 71):  # ops/test.py:130 channel sweep (+ its forward-test D=2)
    w = WORKLOADS["optest"]
    CASES.append((f"optest_D{D}_f64", Workload(f"optest_D{D}", w.N, w.levels, w.Lq, w.M, w.P, D), 3, "unit", "f64", True))
CASES.append(("optest_wide_f64", WORKLOADS["optest"], 4, "wide", "f64", True))
CASES.append(("ragged_wide_f64", Workload("ragged", 2, ((5, 7), (1, 9), (4, 1)), 7, M=3, P=3, D=5), 5, "wide", "f64", True))
CASES.append(("small4lvl_wide_f32", Workload("small4lvl", 2, ((12, 16), (6, 8), (3, 4), (2, 2)), 37, M=8, P=4, D=32), 6, "wide", "f32", True))
CASES.append(("C1_unit_f32", WORKLOADS["C1"], 3, "unit", "f32", False))
CASES.append(("C2_unit_f32", WORKLOADS["C2"], 3, "unit", "f32", False))
CASES.append(("C2_wide_f32", WORKLOADS["C2"], 7, "wide", "f32", False))


def main():
    os.makedirs(GOLD, exist_ok=True)
    assert ref_loader.available(), "needs /root/reference"
    torch.set_num_threads(8)
    for name, w, seed, mode, rdt, full_gv in CASES:
        tdt = torch.float64 if rdt == "f64" else torch.float32
        # inputs are always DRAWN in float32 then widened, so f32 and f64 runs see the same numbers
        x = torch_inputs(w, seed, mode, dtype=tdt)
        out, gv, gl, ga = ref_loader.reference_fwd_bwd(x["value"], x["shapes"], x["loc"], x["attn"], x["grad_out"])
        rec = dict(
            name=name, N=w.N, levels=np.asarray(w.levels, np.int32), Lq=w.Lq, M=w.M, P=w.P, D=w.D,
            seed=seed, loc_mode=mode, ref_dtype=rdt,
            out=out.numpy(), grad_loc=gl.numpy(), grad_attn=ga.numpy(),
        )
        if rdt == "f32":
            # also the float64 evaluation of the same float32 inputs = the value both fp32 codes approximate
            x64 = {k: (v.double() if v.is_floating_point() else v) for k, v in x.items()}
            o64, gv64, gl64, ga64 = ref_loader.reference_fwd_bwd(
                x64["value"], x64["shapes"], x64["loc"], x64["attn"], x64["grad_out"]
            )
            rec.update(out64=o64.numpy().astype(np.float32), grad_loc64=gl64.numpy().astype(np.float32),
                       grad_attn64=ga64.numpy().astype(np.float32))
            gv_best = gv64
        else:
            gv_best = gv
        if full_gv:
            rec["grad_value"] = gv.numpy()
        else:
            rng = np.random.default_rng(1234)
            flat = gv_best.numpy().reshape(-1)
            nz = np.flatnonzero(flat)
            pick = np.sort(np.concatenate([rng.choice(nz, 3072, replace=False), rng.choice(flat.size, 1024, replace=False)]))
            rec.update(
                gv_sum_channels=gv_best.sum(-1).numpy().astype(np.float32),
                gv_sum_pixels=gv_best.sum(1).numpy().astype(np.float32),
                gv_pick_idx=pick.astype(np.int64),
                gv_pick_val=flat[pick].astype(np.float32),
            )
        path = os.path.join(GOLD, name + ".npz")
        np.savez_compressed(path, **rec)
        print(f"{name:24s} out{tuple(out.shape)} -> {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
