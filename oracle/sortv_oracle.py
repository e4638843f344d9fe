"""ctypes front-end of the C oracle for `sort_vertices` (oracle/sortv_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of sortv_oracle.c.  Imported by tests/, oracle/make_golden_sortv.py and the
comparison legs of tools/bench_sortv.py, never by the product package ``aloception_oss_b200``.

Restates aloscene/utils/rotated_iou/cuda_op/sort_vert_kernel.cu:16-134 on the CPU.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "sortv_oracle.c")
_SO = os.path.join(_HERE, "libsortv_oracle.so")
_lock = threading.Lock()
_lib = None


def build(force: bool = False) -> str:
    """Compile sortv_oracle.c with gcc next to its source (no contraction: the one GPU fma is written out); returns the path."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c11", "-o", _SO, _SRC, "-lm"]
        subprocess.run(cmd, check=True, capture_output=True)
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    with _lock:
        if _lib is None:
            _lib = ctypes.CDLL(build())
            _lib.sortv_oracle.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int] * 3
            _lib.sortv_oracle.restype = None
    return _lib


def sort_vertices(vertices, mask, num_valid) -> np.ndarray:
    """(b, n, m, 2) float32, (b, n, m) bool, (b, n) int32 -> (b, n, 9) int32 (sort_vert.cpp:6-29 semantics)."""
    vertices = np.ascontiguousarray(vertices, dtype=np.float32)
    mask = np.ascontiguousarray(mask).astype(np.uint8)
    num_valid = np.ascontiguousarray(num_valid, dtype=np.int32)
    b, n, m, two = vertices.shape
    assert two == 2 and mask.shape == (b, n, m) and num_valid.shape == (b, n)
    idx = np.zeros((b, n, 9), dtype=np.int32)
    if b * n:
        lib().sortv_oracle(vertices.ctypes.data, mask.ctypes.data, num_valid.ctypes.data, idx.ctypes.data, b, n, m)
    return idx


def sort_indices(vertices, mask) -> np.ndarray:
    """numpy restatement of box_intersection_2d.py:132-154 around the oracle (float32 arithmetic, sequential sums)."""
    vertices = np.asarray(vertices, dtype=np.float32)
    mask = np.asarray(mask).astype(bool)
    num_valid = mask.sum(axis=2).astype(np.int32)
    with np.errstate(invalid="ignore", divide="ignore"):
        mean = (vertices * mask[..., None].astype(np.float32)).sum(axis=2, keepdims=True, dtype=np.float32) / \
            num_valid[..., None, None].astype(np.float32)
    return sort_vertices(vertices - mean, mask, num_valid).astype(np.int64)


def calculate_area(idx_sorted, vertices):
    """Shoelace formula over the ordered polygon (box_intersection_2d.py:157-174), float64 for headroom."""
    v = np.asarray(vertices, dtype=np.float64)
    sel = np.take_along_axis(v, np.repeat(np.asarray(idx_sorted)[..., None], 2, axis=-1), axis=2)
    total = sel[:, :, :-1, 0] * sel[:, :, 1:, 1] - sel[:, :, :-1, 1] * sel[:, :, 1:, 0]
    return np.abs(total.sum(axis=2)) / 2, sel
