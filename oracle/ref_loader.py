"""Load the UNMODIFIED reference implementation from /root/reference (this container only).

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, so nothing in the
``-m gpu`` tests, ``smoke()`` or ``bench.py`` may call this at run time; it is used by
oracle/make_golden.py (to generate tests/golden/) and by tests/test_oracle_vs_reference.py
(skipped when the tree is absent).

Recipe: SURVEY.md Appendix A -- the reference file's only ``alonet`` import is ``ALONET_ROOT``
(alonet/deformable_detr/ops/functions/ms_deform_attn_func.py:19), so a stub module satisfies it and
the heavy ``alonet/__init__.py`` (matplotlib, pytorch_lightning, ...) is never executed.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MSDA_REFERENCE_ROOT", "/root/reference")
_FUNC_FILE = os.path.join(REFERENCE_ROOT, "alonet/deformable_detr/ops/functions/ms_deform_attn_func.py")


def available() -> bool:
    return os.path.isfile(_FUNC_FILE)


def load_reference_functions():
    """Returns the reference module object holding ``ms_deform_attn_core_pytorch``."""
    if not available():
        raise FileNotFoundError(f"reference tree not present at {REFERENCE_ROOT}")
    name = "_reference_ms_deform_attn_func"
    if name in sys.modules:
        return sys.modules[name]
    injected = False
    if "alonet" not in sys.modules:
        stub = types.ModuleType("alonet")
        stub.ALONET_ROOT = os.path.join(REFERENCE_ROOT, "alonet")
        stub.__path__ = []
        sys.modules["alonet"] = stub
        injected = True
    try:
        spec = importlib.util.spec_from_file_location(name, _FUNC_FILE)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if injected:
            del sys.modules["alonet"]
    sys.modules[name] = mod
    return mod


def reference_forward(value, spatial_shapes, sampling_locations, attention_weights):
    return load_reference_functions().ms_deform_attn_core_pytorch(
        value, spatial_shapes, sampling_locations, attention_weights
    )


def reference_fwd_bwd(value, spatial_shapes, sampling_locations, attention_weights, grad_output):
    """Forward with the reference function, gradients by autograd through it."""
    import torch

    v = value.detach().clone().requires_grad_(True)
    loc = sampling_locations.detach().clone().requires_grad_(True)
    a = attention_weights.detach().clone().requires_grad_(True)
    out = reference_forward(v, spatial_shapes, loc, a)
    gv, gl, ga = torch.autograd.grad(out, (v, loc, a), grad_output.reshape_as(out))
    return out.detach(), gv, gl, ga
