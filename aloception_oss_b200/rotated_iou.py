"""Host-side mirror of the reference's rotated-IoU `sort_vertices` op over libsortv_b200.so.

Same names and argument meaning as the reference:

  * ``sort_vertices_forward(vertices, mask, num_valid)``  -- the pybind entry point of the reference extension
    (aloscene/utils/rotated_iou/cuda_op/sort_vert.cpp:6-29: contiguous + CUDA + float32 / bool / int32 checks, allocates the
    (b, n, 9) int32 result);
  * ``SortVertices`` / ``sort_v``  -- the autograd wrapper (cuda_op/cuda_ext.py:13-30: forward only, output marked
    non-differentiable);
  * ``sort_indices`` / ``calculate_area``  -- the two callers either side of the op (box_intersection_2d.py:132-174), so a
    maintainer can switch ``from aloscene.utils.rotated_iou.cuda_op.cuda_ext import sort_v`` to this module and nothing else.

No fallback: the C-ABI library (include/sortv_b200.h) must be built (``__graft_entry__.build()``); CPU tensors raise like the
reference's CHECK_CUDA.  The CPU oracle (oracle/sortv_oracle.c) is test infrastructure and is never imported here.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess

import torch
from torch.autograd import Function

_PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_PKG)
LIB_PATH = os.path.join(_PKG, "libsortv_b200.so")
SOURCE = os.path.join(_PKG, "csrc", "sortv_capi.cu")
HEADER = os.path.join(ROOT, "include", "sortv_b200.h")
ABI_VERSION = 1
MAX_NUM_VERT_IDX = 9  # sort_vert_kernel.cu:6

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC", "-shared"]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(d) > t for d in (SOURCE, HEADER))


def build_library(force: bool = False) -> str:
    """Compile csrc/sortv_capi.cu for sm_100a into the in-tree shared library (nvcc cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB_PATH, SOURCE]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    return LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    """The loaded C-ABI library; raises (never falls back) if it is unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the sm_100a extension has not been built. "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). There is no CPU fallback."
        )
    L = ctypes.CDLL(LIB_PATH)
    L.sortv_version.restype = ctypes.c_int
    L.sortv_version.argtypes = []
    L.sortv_last_error_string.restype = ctypes.c_char_p
    L.sortv_last_error_string.argtypes = []
    L.sortv_kernel_launch_count.restype = ctypes.c_uint64
    L.sortv_kernel_launch_count.argtypes = []
    L.sortv_set_variant.restype = ctypes.c_int
    L.sortv_set_variant.argtypes = [ctypes.c_int]
    L.sortv_sort_vertices.restype = ctypes.c_int
    L.sortv_sort_vertices.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int] * 3 + [ctypes.c_void_p]
    if L.sortv_version() != ABI_VERSION:
        raise RuntimeError(f"libsortv_b200.so ABI {L.sortv_version()} != binding ABI {ABI_VERSION}: rebuild")
    _lib = L
    return L


def last_error() -> str:
    return lib().sortv_last_error_string().decode("utf-8", "replace")


def kernel_launch_count() -> int:
    return int(lib().sortv_kernel_launch_count())


def set_variant(variant: int) -> None:
    """Test / measurement knob (include/sortv_b200.h): 0 default (balanced TMA tile kernel), 1 register kernels, 2 generic kernel, 3 unbalanced tile kernel, 4 no sorted-order fast path."""
    if lib().sortv_set_variant(int(variant)) != 0:
        raise ValueError(last_error())


def _check(cond: bool, what: str) -> None:
    if not cond:
        raise RuntimeError(what)


def sort_vertices_forward(vertices: torch.Tensor, mask: torch.Tensor, num_valid: torch.Tensor) -> torch.Tensor:
    """(b, n, m, 2) float32, (b, n, m) bool, (b, n) int32, all contiguous CUDA tensors -> (b, n, 9) int32.

    Error behaviour of sort_vert.cpp:7-15 (utils.h CHECK_* macros raise RuntimeError through TORCH_CHECK)."""
    for name, t in (("vertices", vertices), ("mask", mask), ("num_valid", num_valid)):
        _check(t.is_contiguous(), f"{name} must be a contiguous tensor")
    for name, t in (("vertices", vertices), ("mask", mask), ("num_valid", num_valid)):
        _check(t.is_cuda, f"{name} must be a CUDA tensor")
    _check(vertices.dtype == torch.float32, "vertices must be a float tensor")
    _check(mask.dtype == torch.bool, "mask must be a bool tensor")
    _check(num_valid.dtype == torch.int32, "num_valid must be a int tensor")
    _check(vertices.dim() == 4 and vertices.size(3) == 2, "vertices must have shape (b, n, m, 2)")
    b, n, m = vertices.size(0), vertices.size(1), vertices.size(2)
    _check(tuple(mask.shape) == (b, n, m), "mask must have shape (b, n, m)")
    _check(tuple(num_valid.shape) == (b, n), "num_valid must have shape (b, n)")
    _check(mask.device == vertices.device and num_valid.device == vertices.device, "tensors must be on the same device")
    L = lib()
    with torch.cuda.device(vertices.device):
        # every element is written by the kernel (the reference zero-fills first, sort_vert.cpp:20-21)
        idx = torch.empty((b, n, MAX_NUM_VERT_IDX), dtype=torch.int32, device=vertices.device)
        stream = torch.cuda.current_stream(vertices.device).cuda_stream
        rc = L.sortv_sort_vertices(vertices.data_ptr(), mask.data_ptr(), num_valid.data_ptr(), idx.data_ptr(), b, n, m, stream)
    if rc != 0:
        raise RuntimeError(last_error())
    return idx


class SortVertices(Function):
    """cuda_op/cuda_ext.py:13-27 -- forward only; the index tensor carries no gradient."""

    @staticmethod
    def forward(ctx, vertices, mask, num_valid):
        idx = sort_vertices_forward(vertices, mask, num_valid)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, gradout):
        return ()


sort_v = SortVertices.apply


def sort_indices(vertices: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """box_intersection_2d.py:132-154: (B, N, 24, 2) vertices + (B, N, 24) mask -> (B, N, 9) int64 polygon order."""
    num_valid = torch.sum(mask.int(), dim=2).int()
    mean = torch.sum(vertices * mask.float().unsqueeze(-1), dim=2, keepdim=True) / num_valid.unsqueeze(-1).unsqueeze(-1)
    vertices_normalized = vertices - mean
    return sort_v(vertices_normalized, mask, num_valid).long()


def calculate_area(idx_sorted: torch.Tensor, vertices: torch.Tensor):
    """box_intersection_2d.py:157-174: shoelace area of the ordered polygon; returns (area (B, N), selected (B, N, 9, 2))."""
    idx_ext = idx_sorted.unsqueeze(-1).repeat([1, 1, 1, 2])
    selected = torch.gather(vertices, 2, idx_ext)
    total = selected[:, :, 0:-1, 0] * selected[:, :, 1:, 1] - selected[:, :, 0:-1, 1] * selected[:, :, 1:, 0]
    total = torch.sum(total, dim=2)
    area = torch.abs(total) / 2
    return area, selected
