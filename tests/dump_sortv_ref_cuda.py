"""Run the reference's OWN sort_vertices CUDA kernel (oracle/_ref/sort_vertices_ref.so, built from the reference sources
by oracle/build_ref_sortv.py) on the committed fixture inputs ON THE GPU BOX and write its indices to
gpurun_out/sortv_ref_idx.npz; `python oracle/make_golden_sortv.py --merge gpurun_out/sortv_ref_idx.npz` then stores them in
tests/golden_sortv/ as `idx_ref_cuda`.  Test infrastructure (lives under tests/ because it executes oracle/)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import build_ref_sortv  # noqa: E402


def main():
    assert build_ref_sortv.built(), "oracle/_ref/sort_vertices_ref.so missing (python oracle/build_ref_sortv.py)"
    out = {}
    for name in ("known_answers", "random_pairs"):
        r = np.load(os.path.join(ROOT, "tests", "golden_sortv", name + ".npz"))
        v = torch.from_numpy(r["vertices_norm"]).cuda()
        m = torch.from_numpy(r["mask"]).cuda()
        nv = torch.from_numpy(r["num_valid"]).cuda()
        idx = build_ref_sortv.reference_sort_vertices(v, m, nv).cpu().numpy()
        same = (idx == r["idx_oracle"]).all()
        print(f"{name}: reference CUDA kernel vs oracle identical: {bool(same)}")
        out[name] = idx
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "sortv_ref_idx.npz"), **out)


if __name__ == "__main__":
    main()
