"""Registration of ``torch.ops.alonet_custom.ms_deform_attn_{forward,backward}``.

Replaces ``TORCH_LIBRARY(alonet_custom, m)`` of the reference (alonet/deformable_detr/ops/src/vision.cpp:21-24)
with the same two schemas, so the UNMODIFIED reference ``MSDeformAttnFunction`` (and anything that looks the
ops up by name, e.g. the exporter's graph surgery, deformable_detr/trt_exporter.py:41) resolves to the sm_100a
kernels.  Adds what the reference lacks: a Meta (fake-tensor) kernel, and a CPU kernel that raises the
reference's "Not implemented on the CPU" error (ops/src/ms_deform_attn.h:38,60) instead of a dispatcher miss.
"""
from __future__ import annotations

import threading

import torch

_lock = threading.Lock()
_lib = None
_shim = False  # the ops are registered by libmsda_torch_shim.so (C++: csrc/msda_torch_shim.cpp) instead of from Python


def using_shim() -> bool:
    return _shim

FWD_SCHEMA = ("ms_deform_attn_forward(Tensor value, Tensor spatial_shapes, Tensor level_start_index, "
              "Tensor sampling_loc, Tensor attn_weight, int im2col_step) -> Tensor")
BWD_SCHEMA = ("ms_deform_attn_backward(Tensor value, Tensor spatial_shapes, Tensor level_start_index, "
              "Tensor sampling_loc, Tensor attn_weight, Tensor grad_output, int im2col_step) -> Tensor[]")


def _fwd_cuda(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step):
    from .functions import ms_deform_attn_forward

    return ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step)


def _bwd_cuda(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output, im2col_step):
    from .functions import ms_deform_attn_backward

    return ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                                   im2col_step)


def _fwd_cpu(value, *args):
    raise RuntimeError("Not implemented on the CPU")


def _bwd_cpu(value, *args):
    raise RuntimeError("Not implemented on the CPU")


def _fwd_meta(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step):
    N, _, M, D = value.shape
    return value.new_empty((N, sampling_loc.shape[1], M * D))


def _bwd_meta(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output, im2col_step):
    return [torch.empty_like(value), torch.empty_like(sampling_loc), torch.empty_like(attn_weight)]


def registered() -> bool:
    return hasattr(torch.ops, "alonet_custom") and hasattr(torch.ops.alonet_custom, "ms_deform_attn_forward")


def register() -> None:
    """Idempotent.  Raises if another library (e.g. the reference's own .so) already owns the namespace."""
    global _lib, _shim
    with _lock:
        if _lib is not None:
            return
        from . import _capi

        if _capi.shim_enabled() and not registered():
            # C++ registration: same schemas / checks / messages, no interpreter or ctypes on the call path
            try:
                torch.ops.load_library(_capi.SHIM_PATH)
                _lib, _shim = "shim", True
                return
            except Exception:  # stale build against another torch: fall through to the Python registration
                if registered():
                    raise
        try:
            lib = torch.library.Library("alonet_custom", "DEF")
        except RuntimeError as e:  # pragma: no cover - needs the reference .so in-process
            raise RuntimeError(
                "torch namespace 'alonet_custom' is already defined (the reference MultiScaleDeformableAttention.so "
                "loaded in this process?); the B200 operator cannot co-exist with it") from e
        lib.define(FWD_SCHEMA)
        lib.define(BWD_SCHEMA)
        lib.impl("ms_deform_attn_forward", _fwd_cuda, "CUDA")
        lib.impl("ms_deform_attn_backward", _bwd_cuda, "CUDA")
        lib.impl("ms_deform_attn_forward", _fwd_cpu, "CPU")
        lib.impl("ms_deform_attn_backward", _bwd_cpu, "CPU")
        lib.impl("ms_deform_attn_forward", _fwd_meta, "Meta")
        lib.impl("ms_deform_attn_backward", _bwd_meta, "Meta")
        _lib = lib
