"""Recipe: make the reference's OWN CPU implementation of the path travel to the GPU box.

``bench.py --impl reference`` and the ``cpu_baseline`` leg time ``ms_deform_attn_core_pytorch`` + autograd
(alonet/deformable_detr/ops/functions/ms_deform_attn_func.py:85-190) on the box's host cores.  /root/reference does not exist
there, so -- exactly like the compiled ``oracle/_ref/*.so`` of the reference's CUDA kernels -- this script places an
UNMODIFIED copy of that one source file under the git-ignored ``oracle/_ref/`` (listed in .gitignore, not in .gpurunignore):
nothing of the reference enters the repository's history, and the file is rebuilt from where it lies by
``__graft_entry__.build()`` whenever /root/reference is present.

TEST / MEASUREMENT INFRASTRUCTURE ONLY: loaded by bench.py's reference arm and by tests that compare the port with it; never
imported by the product package.
"""
from __future__ import annotations

import importlib.util
import os
import shutil
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("MSDA_REFERENCE_ROOT", "/root/reference")
_SRC = os.path.join(REFERENCE_ROOT, "alonet/deformable_detr/ops/functions/ms_deform_attn_func.py")
_DST_DIR = os.path.join(_HERE, "_ref", "alonet_ref_py")
_DST = os.path.join(_DST_DIR, "ms_deform_attn_func.py")


def reference_available() -> bool:
    return os.path.isfile(_SRC)


def build() -> str:
    """Copy the reference file (byte for byte) into oracle/_ref/alonet_ref_py/."""
    os.makedirs(_DST_DIR, exist_ok=True)
    shutil.copyfile(_SRC, _DST)
    return _DST


def bundled() -> bool:
    return os.path.isfile(_DST)


def load():
    """The bundled reference module (``ms_deform_attn_core_pytorch`` lives in it).  Its only ``alonet`` import is
    ``ALONET_ROOT`` (ms_deform_attn_func.py:19); a stub module satisfies it."""
    name = "_bundled_reference_ms_deform_attn_func"
    if name in sys.modules:
        return sys.modules[name]
    if not bundled():
        raise FileNotFoundError(_DST)
    injected = False
    if "alonet" not in sys.modules:
        stub = types.ModuleType("alonet")
        stub.ALONET_ROOT = _DST_DIR
        stub.__path__ = []
        sys.modules["alonet"] = stub
        injected = True
    try:
        spec = importlib.util.spec_from_file_location(name, _DST)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if injected:
            del sys.modules["alonet"]
    sys.modules[name] = mod
    return mod


def fwd_bwd(value, shapes, loc, attn, grad_out):
    """One forward + autograd backward of the reference's CPU path; returns (out, grad_value, grad_loc, grad_attn)."""
    import torch

    ref = load()
    v = value.detach().requires_grad_(True)
    l = loc.detach().requires_grad_(True)
    a = attn.detach().requires_grad_(True)
    out = ref.ms_deform_attn_core_pytorch(v, shapes, l, a)
    gv, gl, ga = torch.autograd.grad(out, (v, l, a), grad_out.reshape_as(out))
    return out.detach(), gv, gl, ga


if __name__ == "__main__":
    print(build() if reference_available() else "reference tree not present")
