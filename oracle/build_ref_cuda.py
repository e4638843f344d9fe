"""Build the reference's OWN CUDA extension for sm_100a into oracle/_ref/ (comparison arm, test infrastructure).

    python oracle/build_ref_cuda.py

The reference ships generic SIMT kernels (alonet/deformable_detr/ops/src/cuda/ms_deform_im2col_cuda.cuh) and no
Blackwell path; recompiled for sm_100a they are "the reference on the same box" -- the kernels this repository is
measured against in tests/compare_ref_cuda.py.  Nothing of the reference is copied into the repository: the sources
are compiled where they lie, through a scratch copy under /tmp that carries the two edits without which they do
not build / cannot be loaded next to our operator:

  * ms_deform_attn_cuda.cu:64,134  ``AT_DISPATCH_FLOATING_TYPES(value.type(), ...)`` -> ``value.scalar_type()``
    (torch >= 2.x no longer converts DeprecatedTypeProperties to ScalarType);
  * vision.cpp:21  ``TORCH_LIBRARY(alonet_custom, m)`` -> ``TORCH_LIBRARY(alonet_ref, m)`` (a second definition of
    the ``alonet_custom`` namespace in one process aborts; our operator owns that name).

Output: oracle/_ref/alonet_ref_msda.so (git-ignored, travels to the GPU box).  Only runs where /root/reference exists.
"""
from __future__ import annotations

import os
import re
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("MSDA_REFERENCE_ROOT", "/root/reference")
SRC = os.path.join(REFERENCE_ROOT, "alonet/deformable_detr/ops/src")
OUT_DIR = os.path.join(HERE, "_ref")
OUT_SO = os.path.join(OUT_DIR, "alonet_ref_msda.so")
SCRATCH = "/tmp/msda_ref_src"


def reference_available() -> bool:
    return os.path.isdir(SRC)


def built() -> bool:
    return os.path.exists(OUT_SO)


def build(force: bool = False) -> str:
    if built() and not force:
        return OUT_SO
    if not reference_available():
        raise FileNotFoundError(SRC)
    if os.path.isdir(SCRATCH):
        shutil.rmtree(SCRATCH)
    shutil.copytree(SRC, SCRATCH)
    cu = os.path.join(SCRATCH, "cuda", "ms_deform_attn_cuda.cu")
    txt = open(cu).read()
    txt, n = re.subn(r"AT_DISPATCH_FLOATING_TYPES\(value\.type\(\)", "AT_DISPATCH_FLOATING_TYPES(value.scalar_type()", txt)
    assert n == 2, n
    open(cu, "w").write(txt)
    vis = os.path.join(SCRATCH, "vision.cpp")
    txt = open(vis).read()
    txt, n = re.subn(r"TORCH_LIBRARY\(alonet_custom,", "TORCH_LIBRARY(alonet_ref,", txt)
    assert n == 1, n
    open(vis, "w").write(txt)

    os.makedirs(OUT_DIR, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    from torch.utils.cpp_extension import load

    sources = [vis, os.path.join(SCRATCH, "cpu", "ms_deform_attn_cpu.cpp"), cu]
    load(
        name="alonet_ref_msda",
        sources=sources,
        extra_include_paths=[SCRATCH],
        extra_cflags=["-O2", "-DWITH_CUDA"],
        extra_cuda_cflags=["-O3", "-DWITH_CUDA", "-DCUDA_HAS_FP16=1", "-D__CUDA_NO_HALF_OPERATORS__",
                           "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__",
                           "-gencode", "arch=compute_100a,code=sm_100a"],
        build_directory=OUT_DIR,
        with_cuda=True,
        is_python_module=False,
        verbose=False,
    )
    assert os.path.exists(OUT_SO), os.listdir(OUT_DIR)
    return OUT_SO


def load_ops():
    """torch.ops.alonet_ref.ms_deform_attn_forward / _backward (GPU box: loads the prebuilt .so)."""
    import torch

    if not hasattr(torch.ops, "alonet_ref") or not hasattr(torch.ops.alonet_ref, "ms_deform_attn_forward"):
        if not built():
            raise FileNotFoundError(f"{OUT_SO} not built (run oracle/build_ref_cuda.py where /root/reference exists)")
        torch.ops.load_library(OUT_SO)
    return torch.ops.alonet_ref


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
