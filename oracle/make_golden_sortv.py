"""Generate tests/golden_sortv/*.npz from the UNMODIFIED reference rotated-IoU pipeline (build container only).

    python oracle/make_golden_sortv.py

TEST INFRASTRUCTURE ONLY.  The reference's `sort_vertices` op exists only as a CUDA kernel (no CPU implementation, no stored
vectors), so the fixtures are made from everything AROUND it that is pure Python and runs here:

  * aloscene/utils/rotated_iou/box_intersection_2d.py (box_intersection_th, box_in_box_th, build_vertices, sort_indices,
    calculate_area) and oriented_iou_loss.py (box2corners_th, cal_iou) are imported from /root/reference with the CUDA op
    `sort_v` (cuda_op/cuda_ext.py:30) replaced by a recorder around oracle/sortv_oracle.py -- the recorder captures the exact
    tensors the reference hands to the op (normalised vertices, mask, num_valid);
  * aloscene/utils/rotated_iou/utiles.py `box_intersection_area` is the reference's INDEPENDENT numpy implementation of the
    intersection area (its own vertex sort by arctan2): the area it returns is stored as the known answer for every pair;
  * the known-answer cases of the reference's tests: _test_corner_cases.py:12-34 (IoU 1, 0, 0.3333, 1),
    unittest/test_oriented_boxes_2d.py:17-88 (IoU 1, 0, 1/3, 1/7, 1/7, 0.5/5.5), _test_box_intersection_2d.py:51-56.

Stored per case: boxes, the op's inputs (vertices_norm, mask, num_valid), the un-normalised vertices (for sort_indices /
calculate_area mirrors), the numpy area / expected IoU, and `idx_oracle`.  `idx_ref_cuda` (the reference's own kernel, built
for sm_100a by oracle/build_ref_sortv.py, run on the GPU box by tests/dump_sortv_ref_cuda.py) is merged in afterwards by
`python oracle/make_golden_sortv.py --merge gpurun_out/sortv_ref_idx.npz`.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import sortv_oracle  # noqa: E402

REFERENCE_ROOT = os.environ.get("MSDA_REFERENCE_ROOT", "/root/reference")
RIOU = os.path.join(REFERENCE_ROOT, "aloscene/utils/rotated_iou")
GOLD = os.path.join(ROOT, "tests", "golden_sortv")

captured = []


def _recording_sort_v(vertices, mask, num_valid):
    captured.append((vertices.detach().clone(), mask.detach().clone(), num_valid.detach().clone()))
    return torch.from_numpy(sortv_oracle.sort_vertices(vertices.numpy(), mask.numpy(), num_valid.numpy()))


def load_reference_pipeline():
    """box_intersection_2d / oriented_iou_loss / utiles of the reference, `aloscene` stubbed to bare namespaces."""
    def ns(name, path):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
        return m

    ns("aloscene", os.path.join(REFERENCE_ROOT, "aloscene"))
    ns("aloscene.utils", os.path.join(REFERENCE_ROOT, "aloscene/utils"))
    ns("aloscene.utils.rotated_iou", RIOU)
    ns("aloscene.utils.rotated_iou.cuda_op", os.path.join(RIOU, "cuda_op"))
    ext = types.ModuleType("aloscene.utils.rotated_iou.cuda_op.cuda_ext")
    ext.sort_v = _recording_sort_v
    sys.modules[ext.__name__] = ext
    for name in ("matplotlib", "matplotlib.pyplot"):  # utiles.py:10 imports pyplot for its demo plots only
        sys.modules.setdefault(name, types.ModuleType(name))
    # min_enclosing_box.py:53 uses np.int (removed in numpy 1.24) at import time; only the GIoU loss needs the module, and
    # nothing here does, so oriented_iou_loss.py:4 gets an empty stand-in
    meb = types.ModuleType("aloscene.utils.rotated_iou.min_enclosing_box")
    meb.smallest_bounding_box = None
    sys.modules[meb.__name__] = meb

    def load(name):
        full = f"aloscene.utils.rotated_iou.{name}"
        spec = importlib.util.spec_from_file_location(full, os.path.join(RIOU, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[full] = mod
        spec.loader.exec_module(mod)
        return mod

    return load("box_intersection_2d"), load("oriented_iou_loss"), load("utiles")


def known_answer_cases():
    """(box1, box2, expected IoU, source) -- boxes are (x, y, w, h, alpha)."""
    s2, p = float(np.sqrt(2)), float(np.pi)
    b0 = [0.0, 0.0, 2.0, 2.0, 0.0]
    return [
        (b0, [0.0, 0.0, 2.0, 2.0, 0.0], 1.0, "_test_corner_cases.py:9-13 / test_oriented_boxes_2d.py:17-26"),
        (b0, [0.0, 2.0, 2.0, 2.0, 0.0], 0.0, "_test_corner_cases.py:16-20"),
        (b0, [0.0, 1.0, 2.0, 2.0, 0.0], 1.0 / 3.0, "_test_corner_cases.py:23-27"),
        ([38.0, 120.0, 1.3, 20.0, 50.0], [38.0, 120.0, 1.3, 20.0, 50.0], 1.0, "_test_corner_cases.py:30-34"),
        (b0, [2.0, 0.0, 2.0, 2.0, 0.0], 0.0, "test_oriented_boxes_2d.py:29-39"),
        (b0, [1.0, 0.0, 2.0, 2.0, 0.0], 1.0 / 3.0, "test_oriented_boxes_2d.py:42-51"),
        (b0, [1.0, 1.0, 2.0, 2.0, 0.0], 1.0 / 7.0, "test_oriented_boxes_2d.py:54-63"),
        (b0, [1.0, 1.0, 2.0, 2.0, p / 2], 1.0 / 7.0, "test_oriented_boxes_2d.py:66-75"),
        (b0, [1.0, 1.0, s2, s2, p / 4], 0.5 / 5.5, "test_oriented_boxes_2d.py:78-88"),
        ([0.0, 0.0, 2.0, 3.0, p / 6], [1.0, 1.0, 4.0, 4.0, -p / 4], None, "_test_box_intersection_2d.py:51-56"),
        ([0.0, 0.0, 2.0, 3.0, p / 6], [0.0, 0.0, 2.0, 3.0, p / 6], 1.0, "_test_box_intersection_2d.py:55 (same box)"),
    ]


def run_case(bi2d, loss, utiles, box1, box2):
    """box1, box2: (B, N, 5) float32 tensors -> dict of arrays."""
    captured.clear()
    corners1 = loss.box2corners_th(box1)
    corners2 = loss.box2corners_th(box2)
    inters, mask_inter = bi2d.box_intersection_th(corners1, corners2)
    c12, c21 = bi2d.box_in_box_th(corners1, corners2)
    vertices, mask = bi2d.build_vertices(corners1, corners2, c12, c21, inters, mask_inter)
    idx = bi2d.sort_indices(vertices, mask)  # -> the recorder
    area, _ = bi2d.calculate_area(idx, vertices)
    vn, mk, nv = captured[-1]
    B, N = box1.shape[:2]
    area_np = np.zeros((B, N), np.float64)
    for i in range(B):
        for j in range(N):
            try:
                a, _ = utiles.box_intersection_area(box1[i, j].double().numpy(), box2[i, j].double().numpy())
            except Exception:  # the numpy version cannot sort fewer than 3 vertices
                a = 0.0
            area_np[i, j] = a
    u = (box1[..., 2] * box1[..., 3] + box2[..., 2] * box2[..., 3]).double().numpy()
    return dict(
        box1=box1.numpy(), box2=box2.numpy(), vertices=vertices.numpy(), vertices_norm=vn.numpy(), mask=mk.numpy(),
        num_valid=nv.numpy(), idx_oracle=idx.numpy().astype(np.int32), area_pipeline=area.numpy(), area_numpy=area_np,
        area_sum=u,
    )


def main():
    assert os.path.isdir(RIOU), "needs /root/reference"
    os.makedirs(GOLD, exist_ok=True)
    bi2d, loss, utiles = load_reference_pipeline()

    cases = known_answer_cases()
    b1 = torch.tensor([c[0] for c in cases], dtype=torch.float32)[None]
    b2 = torch.tensor([c[1] for c in cases], dtype=torch.float32)[None]
    rec = run_case(bi2d, loss, utiles, b1, b2)
    rec["expected_iou"] = np.array([np.nan if c[2] is None else c[2] for c in cases], np.float64)
    rec["source"] = np.array([c[3] for c in cases])
    np.savez_compressed(os.path.join(GOLD, "known_answers.npz"), **rec)
    iou = rec["area_pipeline"][0] / (rec["area_sum"][0] - rec["area_pipeline"][0])
    for c, got, a_np in zip(cases, iou, rec["area_numpy"][0]):
        print(f"  expected IoU {c[2]}  pipeline+oracle {got:.6f}  numpy area {a_np:.6f}   [{c[3]}]")

    # seeded random pairs: overlapping, disjoint, contained, near-axis-aligned and tiny-angle boxes
    g = torch.Generator().manual_seed(11)
    B, N = 4, 192
    xy = torch.rand(B, N, 2, generator=g) * 4 - 2
    wh = torch.rand(B, N, 2, generator=g) * 3 + 0.2
    al = (torch.rand(B, N, 1, generator=g) - 0.5) * 2 * np.pi
    box1 = torch.cat([xy, wh, al], -1)
    d = torch.randn(B, N, 2, generator=g) * 1.2
    wh2 = torch.rand(B, N, 2, generator=g) * 3 + 0.2
    al2 = (torch.rand(B, N, 1, generator=g) - 0.5) * 2 * np.pi
    box2 = torch.cat([xy + d, wh2, al2], -1)
    box2[0, :24] = box1[0, :24]  # identical boxes (the num_valid == 8 corner case)
    box2[1, :24, 4] = box1[1, :24, 4]  # parallel edges
    box2[1, :24, :2] = box1[1, :24, :2]
    box1[2, :24, 4] = 0.0  # axis-aligned pairs: vertices with y exactly 0 after normalisation are likely
    box2[2, :24, 4] = 0.0
    box2[3, :24, 2:4] = box1[3, :24, 2:4] * 0.3  # contained
    box2[3, :24, :2] = box1[3, :24, :2]
    rec = run_case(bi2d, loss, utiles, box1, box2)
    np.savez_compressed(os.path.join(GOLD, "random_pairs.npz"), **rec)
    err = np.abs(rec["area_pipeline"] - rec["area_numpy"])
    print(f"random pairs: {B * N} polygons, num_valid histogram {np.bincount(rec['num_valid'].ravel(), minlength=9)}, "
          f"max |area(pipeline+oracle) - area(numpy)| = {err.max():.3e}")


def merge(path):
    """Merge the reference CUDA kernel's indices (dumped on the GPU box) into the fixtures."""
    dump = np.load(path)
    for name in ("known_answers", "random_pairs"):
        f = os.path.join(GOLD, name + ".npz")
        rec = dict(np.load(f))
        rec["idx_ref_cuda"] = dump[name].astype(np.int32)
        same = (rec["idx_ref_cuda"] == rec["idx_oracle"]).all()
        print(f"{name}: reference CUDA indices merged; identical to the oracle's: {bool(same)}")
        np.savez_compressed(f, **rec)


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--merge":
        merge(sys.argv[2])
    else:
        main()
