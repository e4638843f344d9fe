"""ctypes binding of the C ABI declared in include/msda_b200.h.

The product path has NO fallback: if ``libmsda_b200.so`` is missing or does not load, importing the
operator raises.  (The CPU oracle under oracle/ is test infrastructure and is never imported from here.)
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_PKG)
# MSDA_LIB_PATH / MSDA_NVCC_EXTRA: build or load an experimental variant next to the product library (tuning only)
LIB_PATH = os.environ.get("MSDA_LIB_PATH") or os.path.join(_PKG, "libmsda_b200.so")
CSRC = os.path.join(_PKG, "csrc")
HEADER = os.path.join(ROOT, "include", "msda_b200.h")

ABI_VERSION = 1
F32, BF16, F16, F64 = 0, 1, 2, 3
BWD_PREZEROED = 1  # MSDA_BWD_PREZEROED
BWD_DETERMINISTIC = 2  # MSDA_BWD_DETERMINISTIC
FUSED_REF_F32 = 4  # MSDA_FUSED_REF_F32
LOC_F32, ATTN_F32 = 0x100, 0x200  # MSDA_LOC_F32 / MSDA_ATTN_F32, OR-ed into the dtype of msda_forward / msda_backward

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
]


class MsdaDims(ctypes.Structure):
    """struct msda_dims (include/msda_b200.h)."""

    _fields_ = [(n, ctypes.c_int) for n in
                ("batch", "spatial_size", "num_heads", "channels", "num_levels", "num_query", "num_point")]


def sources():
    return [os.path.join(CSRC, "msda_capi.cu")]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + [os.path.join(CSRC, "msda_kernels.cuh"), os.path.join(CSRC, "msda_bwd_tile.cuh"), os.path.join(CSRC, "msda_fwd_win.cuh"), HEADER]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a into the in-tree shared library (nvcc cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    extra = os.environ.get("MSDA_NVCC_EXTRA", "").split()
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + sources()
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


SHIM_PATH = os.path.join(_PKG, "libmsda_torch_shim.so")
SHIM_SRC = os.path.join(CSRC, "msda_torch_shim.cpp")


def shim_enabled() -> bool:
    """The C++ torch-op registration (csrc/msda_torch_shim.cpp) is used when it has been built, unless MSDA_NO_SHIM=1 or an
    experimental library is selected with MSDA_LIB_PATH (the shim is linked against the product library)."""
    return (os.environ.get("MSDA_NO_SHIM", "0") != "1" and not os.environ.get("MSDA_LIB_PATH") and os.path.exists(SHIM_PATH)
            and os.path.exists(LIB_PATH))


def build_shim(force: bool = False) -> str:
    """Compile csrc/msda_torch_shim.cpp against the installed torch (g++, no nvcc) into the in-tree libmsda_torch_shim.so."""
    deps = [SHIM_SRC, HEADER, LIB_PATH]
    if not force and os.path.exists(SHIM_PATH) and all(os.path.getmtime(d) <= os.path.getmtime(SHIM_PATH) for d in deps if os.path.exists(d)):
        return SHIM_PATH
    import torch

    tl = os.path.dirname(os.path.abspath(torch.__file__))
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
           f"-I{tl}/include", f"-I{tl}/include/torch/csrc/api/include", f"-I{cuda_home}/include", SHIM_SRC, "-o", SHIM_PATH,
           f"-L{_PKG}", "-lmsda_b200", f"-L{tl}/lib", "-ltorch", "-ltorch_cpu", "-lc10", "-lc10_cuda", "-ltorch_cuda",
           "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{tl}/lib"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    return SHIM_PATH


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
    vp, i, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
    dp = ctypes.POINTER(MsdaDims)
    L.msda_version.restype = i
    L.msda_version.argtypes = []
    L.msda_last_error_string.restype = ctypes.c_char_p
    L.msda_last_error_string.argtypes = []
    L.msda_forward.restype = i
    L.msda_forward.argtypes = [vp] * 6 + [dp, i, vp]
    L.msda_forward_workspace_bytes.restype = sz
    L.msda_forward_workspace_bytes.argtypes = [dp, i]
    L.msda_forward_ws.restype = i
    L.msda_forward_ws.argtypes = [vp] * 7 + [sz, dp, i, vp]
    L.msda_forward_host.restype = i
    L.msda_forward_host.argtypes = [vp] * 6 + [dp, i, vp]
    L.msda_backward_workspace_bytes.restype = sz
    L.msda_backward_workspace_bytes.argtypes = [dp, i]
    L.msda_backward_workspace_bytes_ex.restype = sz
    L.msda_backward_workspace_bytes_ex.argtypes = [dp, i, i]
    L.msda_backward.restype = i
    L.msda_backward.argtypes = [vp] * 10 + [sz, dp, i, i, vp]
    L.msda_zero_fill.restype = i
    L.msda_zero_fill.argtypes = [vp, sz, vp]
    L.msda_im2col_inference.restype = i
    L.msda_im2col_inference.argtypes = [vp] * 6 + [i] * 7 + [vp, i]
    L.msda_fused_supported.restype = i
    L.msda_fused_supported.argtypes = [dp, i, i]
    L.msda_fused_forward.restype = i
    L.msda_fused_forward.argtypes = [vp, vp, vp, vp, i, vp, vp, vp, dp, i, vp]
    L.msda_fused_forward_ex.restype = i
    L.msda_fused_forward_ex.argtypes = [vp, vp, vp, vp, i, vp, vp, vp, dp, i, i, vp]
    L.msda_fused_backward.restype = i
    L.msda_fused_backward.argtypes = [vp] * 5 + [i] + [vp] * 7 + [sz, dp, i, i, vp]
    L.msda_set_tuning.restype = i
    L.msda_set_tuning.argtypes = [ctypes.c_char_p, i]
    L.msda_get_tuning.restype = i
    L.msda_get_tuning.argtypes = [ctypes.c_char_p, ctypes.POINTER(i)]
    L.msda_kernel_launch_count.restype = ctypes.c_uint64
    L.msda_kernel_launch_count.argtypes = []
    if L.msda_version() != ABI_VERSION:
        raise RuntimeError(f"libmsda_b200.so ABI {L.msda_version()} != binding ABI {ABI_VERSION}: rebuild")
    _lib = L
    return L


def last_error() -> str:
    return lib().msda_last_error_string().decode("utf-8", "replace")


def set_tuning(name: str, value: int) -> None:
    if lib().msda_set_tuning(name.encode(), int(value)) != 0:
        raise ValueError(last_error())


def get_tuning(name: str) -> int:
    v = ctypes.c_int(0)
    if lib().msda_get_tuning(name.encode(), ctypes.byref(v)) != 0:
        raise ValueError(last_error())
    return v.value


def kernel_launch_count() -> int:
    return int(lib().msda_kernel_launch_count())
