"""Run the UNMODIFIED reference code on the B200 operator.

The reference resolves its native op in two places only:

* ``load_MultiScaleDeformableAttention()`` -- called from every ``MSDeformAttn.__init__``
  (alonet/deformable_detr/ops/modules/ms_deform_attn.py:68); it ``torch.ops.load_library``s the reference ``.so`` and,
  on failure, shells out to ``make.sh`` (ops/functions/ms_deform_attn_func.py:22-46);
* ``torch.ops.alonet_custom.ms_deform_attn_forward/backward`` -- looked up by name at call time
  (ms_deform_attn_func.py:55,72).

``install()`` therefore (1) registers our implementations under the same torch-op names and (2) replaces the loader
by ours in the reference modules that bind it.  Everything above (``MSDeformAttnFunction``, ``MSDeformAttn``,
``DeformableTransformer``, ``DeformableDETR``, the panoptic head, the exporters) is left untouched and simply
finds the ops registered.

If ``alonet`` cannot be imported as a whole (it pulls matplotlib / pytorch_lightning / ... at import time),
``import_reference_ops(alonet_root)`` imports just ``alonet.deformable_detr.ops`` from a source tree through
namespace stubs, with the loader already replaced.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

from .functions import load_MultiScaleDeformableAttention, load_ops


def install(patch_imported: bool = True) -> None:
    """Register the torch ops and point every already-imported reference module at our loader."""
    load_ops()
    if not patch_imported:
        return
    for name in ("alonet.deformable_detr.ops.functions.ms_deform_attn_func", "alonet.deformable_detr.ops.functions",
                 "alonet.deformable_detr.ops.modules.ms_deform_attn"):
        mod = sys.modules.get(name)
        if mod is not None:
            if hasattr(mod, "load_MultiScaleDeformableAttention"):
                mod.load_MultiScaleDeformableAttention = load_MultiScaleDeformableAttention
            if hasattr(mod, "load_ops"):
                mod.load_ops = load_ops


def _namespace(name: str, path: str) -> types.ModuleType:
    mod = sys.modules.get(name)
    if mod is None:
        mod = types.ModuleType(name)
        mod.__path__ = [path]
        sys.modules[name] = mod
    return mod


def import_reference_ops(alonet_root: str):
    """Import the reference's ``alonet.deformable_detr.ops`` (functions + modules) from ``alonet_root`` (the directory
    that contains ``deformable_detr/``) WITHOUT executing ``alonet/__init__.py``, with the loader replaced by ours.

    Returns ``(functions_module, modules_module)``; ``modules_module.MSDeformAttn`` is the reference class, running on
    the B200 kernels."""
    alonet_root = os.path.abspath(alonet_root)
    if not os.path.isdir(os.path.join(alonet_root, "deformable_detr", "ops")):
        raise FileNotFoundError(f"{alonet_root} does not look like an alonet source tree")
    load_ops()
    top = _namespace("alonet", alonet_root)
    if not hasattr(top, "ALONET_ROOT"):
        top.ALONET_ROOT = alonet_root
    _namespace("alonet.deformable_detr", os.path.join(alonet_root, "deformable_detr"))
    _namespace("alonet.deformable_detr.ops", os.path.join(alonet_root, "deformable_detr", "ops"))
    func = importlib.import_module("alonet.deformable_detr.ops.functions.ms_deform_attn_func")
    func.load_MultiScaleDeformableAttention = load_MultiScaleDeformableAttention
    func.load_ops = load_ops
    functions = importlib.import_module("alonet.deformable_detr.ops.functions")
    functions.load_MultiScaleDeformableAttention = load_MultiScaleDeformableAttention
    functions.load_ops = load_ops
    modules = importlib.import_module("alonet.deformable_detr.ops.modules")
    ms = sys.modules.get("alonet.deformable_detr.ops.modules.ms_deform_attn")
    if ms is not None:
        ms.load_MultiScaleDeformableAttention = load_MultiScaleDeformableAttention
    return functions, modules
