"""Build the reference's OWN `sort_vertices` CUDA kernel for sm_100a into oracle/_ref/ (checker, test infrastructure).

    python oracle/build_ref_sortv.py

aloscene/utils/rotated_iou/cuda_op/sort_vert_kernel.cu is compiled WHERE IT LIES (nothing is copied into the repository) with
torch's include path (its cuda_utils.h pulls ATen in for the current-stream query) into a plain shared library; the C++ entry
point ``sort_vertices_wrapper(int, int, int, const float*, const bool*, const int*, int*)`` (sort_vert_kernel.cu:136-140) is
called through ctypes by its mangled name inside a process that has torch loaded.  It is the bit-exact reference for
include/sortv_b200.h on the GPU box (tests/test_sortv_gpu.py); it launches on the legacy default stream.

Output: oracle/_ref/sort_vertices_ref.so (git-ignored, travels to the GPU box).  Only runs where /root/reference exists.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("MSDA_REFERENCE_ROOT", "/root/reference")
SRC_DIR = os.path.join(REFERENCE_ROOT, "aloscene/utils/rotated_iou/cuda_op")
SRC = os.path.join(SRC_DIR, "sort_vert_kernel.cu")
OUT_DIR = os.path.join(HERE, "_ref")
OUT_SO = os.path.join(OUT_DIR, "sort_vertices_ref.so")
MANGLED = "_Z21sort_vertices_wrapperiiiPKfPKbPKiPi"


def reference_available() -> bool:
    return os.path.isfile(SRC)


def built() -> bool:
    return os.path.exists(OUT_SO)


def build(force: bool = False) -> str:
    if built() and not force:
        return OUT_SO
    if not reference_available():
        raise FileNotFoundError(SRC)
    from torch.utils.cpp_extension import include_paths, library_paths

    os.makedirs(OUT_DIR, exist_ok=True)
    # linked against libc10 / libc10_cuda (at::cuda::getCurrentCUDAStream, sort_vert_kernel.cu:137): torch's libraries are
    # loaded RTLD_LOCAL, so the symbol must come in through DT_NEEDED, found via rpath where torch is installed
    libs = []
    for lp in library_paths():
        libs += ["-L" + lp, "-Xlinker", "-rpath=" + lp]
    cmd = ["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-shared",
           "-D_GLIBCXX_USE_CXX11_ABI=1", "-I" + SRC_DIR] + ["-I" + p for p in include_paths()] + ["-o", OUT_SO, SRC] + \
        libs + ["-lc10", "-lc10_cuda"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stderr[-3000:])
    return OUT_SO


def reference_sort_vertices(vertices, mask, num_valid):
    """Run the reference kernel on CUDA tensors: (b, n, m, 2) f32, (b, n, m) bool, (b, n) i32 -> (b, n, 9) i32."""
    import torch  # noqa: F401  (its shared libraries resolve the at::cuda symbols of the reference object)

    lib = ctypes.CDLL(OUT_SO)
    fn = getattr(lib, MANGLED)
    fn.restype = None
    fn.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p] * 4
    b, n, m = vertices.shape[:3]
    idx = torch.zeros((b, n, 9), dtype=torch.int32, device=vertices.device)
    torch.cuda.synchronize()
    fn(b, n, m, vertices.data_ptr(), mask.data_ptr(), num_valid.data_ptr(), idx.data_ptr())
    torch.cuda.synchronize()
    return idx


if __name__ == "__main__":
    print(build(force=True))
