"""Host side of the operator: mirror of ``alonet/deformable_detr/ops/functions``.

Same names, argument meaning and error behaviour as the reference
(alonet/deformable_detr/ops/functions/ms_deform_attn_func.py and ops/functions/__init__.py:9-14):

* ``MSDeformAttnFunction``                      <- ms_deform_attn_func.py:49-82
* ``load_MultiScaleDeformableAttention``        <- ms_deform_attn_func.py:29-46 (here: register the torch ops, never shells out)
* ``load_ops``                                  <- ms_deform_attn_func.py:22-26
* ``ms_deform_attn_core_pytorch``               <- ms_deform_attn_func.py:85-107 (tracing / export path only)
* ``ms_deform_attn_forward`` / ``_backward``    <- the C++ entry points behind torch.ops.alonet_custom.*
  (ops/src/ms_deform_attn.h:20-62, ops/src/cuda/ms_deform_attn_cuda.cu:20-153)

The compute is done by the C-ABI library (include/msda_b200.h) through ctypes; torch supplies device
memory and the current stream.  There is no CPU implementation: CPU tensors raise, as in the reference
("Not implemented on the CPU", ops/src/ms_deform_attn.h:38,60).
"""
from __future__ import annotations

import ctypes
import os

import torch
import torch.nn.functional as F
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _capi
from . import torch_ops as _torch_ops

# MSDA_EARLY_ZERO_FILL=1: MSDeformAttnFunction starts the backward's zero-fill on a side stream during the forward pass.
# Off by default: with the backward directly behind the forward it is SLOWER than the in-stream zero-fill -> scatter pair
# under programmatic dependent launch (C2 step 19.3 -> 21.4 us, profiles/r1_sweep_rejected_early_zero_fill.jsonl: the
# cross-stream fork / join costs more than the 4.9 us fill it hides); it can only pay when other work separates the two.
EARLY_ZERO_FILL = os.environ.get("MSDA_EARLY_ZERO_FILL", "0") == "1"
# MSDA_DETERMINISTIC=1 (or torch.use_deterministic_algorithms(True)): the backward accumulates grad_value in 64-bit fixed
# point (C-ABI flag MSDA_BWD_DETERMINISTIC) -- bit-reproducible, unlike the reference's atomicAdd scatter
# (ms_deform_im2col_cuda.cuh:125-152) and the default red-based path here.
DETERMINISTIC = os.environ.get("MSDA_DETERMINISTIC", "0") == "1"


def deterministic_requested() -> bool:
    return DETERMINISTIC or torch.are_deterministic_algorithms_enabled()

_DTYPES = {torch.float32: _capi.F32, torch.bfloat16: _capi.BF16, torch.float16: _capi.F16, torch.float64: _capi.F64}


def _check_inputs(named):
    # same order and wording as the AT_ASSERTM blocks of ms_deform_attn_cuda.cu:28-38 / :93-105
    for name, t in named:
        if not t.is_contiguous():
            raise RuntimeError(f"{name} tensor has to be contiguous")
    for name, t in named:
        if not t.is_cuda:
            raise RuntimeError(f"{name} must be a CUDA tensor")


def _dims_of(value, spatial_shapes, sampling_loc, attn_weight):
    if value.dim() != 4:
        raise RuntimeError(f"value must be (N, S, M, D), got {tuple(value.shape)}")
    if sampling_loc.dim() != 6 or sampling_loc.shape[-1] != 2:
        raise RuntimeError(f"sampling_loc must be (N, Lq, M, L, P, 2), got {tuple(sampling_loc.shape)}")
    N, S, M, D = value.shape
    L = spatial_shapes.shape[0]
    Lq, P = sampling_loc.shape[1], sampling_loc.shape[4]
    if spatial_shapes.dim() != 2 or spatial_shapes.shape[1] != 2:
        raise RuntimeError(f"spatial_shapes must be (L, 2), got {tuple(spatial_shapes.shape)}")
    if tuple(sampling_loc.shape) != (N, Lq, M, L, P, 2):
        raise RuntimeError(
            f"sampling_loc {tuple(sampling_loc.shape)} inconsistent with value {tuple(value.shape)} and {L} levels")
    if tuple(attn_weight.shape) != (N, Lq, M, L, P):
        raise RuntimeError(f"attn_weight must be {(N, Lq, M, L, P)}, got {tuple(attn_weight.shape)}")
    return _capi.MsdaDims(N, S, M, D, L, Lq, P)


def _meta_i32(t, name):
    if t.dtype == torch.int32:
        return t
    if t.dtype == torch.int64:  # upstream Deformable-DETR passes int64; this fork int32
        return t.to(torch.int32)
    raise RuntimeError(f"{name} must be an int32 (or int64) tensor, got {t.dtype}")


def _common_checks(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step, extra=()):
    _check_inputs([("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
                   ("sampling_loc", sampling_loc), ("attn_weight", attn_weight), *extra])
    if value.dtype not in _DTYPES:
        raise RuntimeError(f"ms_deform_attn: unsupported dtype {value.dtype} (float32, float64, bfloat16, float16)")
    half = value.dtype in (torch.bfloat16, torch.float16)
    for name, t in (("sampling_loc", sampling_loc), ("attn_weight", attn_weight), *extra):
        # mixed precision: sampling_loc / attn_weight may stay float32 next to 16-bit value (MSDA_LOC_F32 / MSDA_ATTN_F32)
        if t.dtype != value.dtype and not (half and t.dtype == torch.float32 and name != "grad_output"):
            raise RuntimeError(f"{name} has dtype {t.dtype}, expected {value.dtype} (same as value)")
        if t.device != value.device:
            raise RuntimeError(f"{name} is on {t.device}, value on {value.device}")
    dims = _dims_of(value, spatial_shapes, sampling_loc, attn_weight)
    if level_start_index.numel() != dims.num_levels:
        raise RuntimeError("level_start_index must have one entry per level")
    batch = dims.batch
    step = min(batch, int(im2col_step))
    # ms_deform_attn_cuda.cu:50-52 -- kept for API fidelity; the kernels themselves do not batch by im2col_step
    if batch > 0 and (step <= 0 or batch % step != 0):
        raise RuntimeError(f"batch({batch}) must divide im2col_step({step})")
    return dims


def _io_dtype(value, sampling_loc, attn_weight):
    """The C ABI's dtype word: value's type, plus MSDA_LOC_F32 / MSDA_ATTN_F32 for fp32 locations / weights next to 16-bit value."""
    dt = _DTYPES[value.dtype]
    if sampling_loc.dtype != value.dtype:
        dt |= _capi.LOC_F32
    if attn_weight.dtype != value.dtype:
        dt |= _capi.ATTN_F32
    return dt


def _ptr(t):
    # plain int (ctypes converts it for a c_void_p parameter): building c_void_p objects cost ~1 us per argument
    return t.data_ptr() if t.numel() else None


# raw-stream / current-device queries without the Python wrappers of torch.cuda (5 us -> 0.5 us per call); the public API is
# the fallback if a torch build does not expose them
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_get_device = getattr(torch._C, "_cuda_getDevice", None)


class _on_device:
    """``with _on_device(t) as stream_handle``: makes t's GPU current only if it is not already (the common case costs
    two cheap queries instead of a device-guard round trip) and yields the raw current ``cudaStream_t``."""

    __slots__ = ("idx", "prev")

    def __init__(self, t):
        self.idx = t.device.index
        self.prev = None

    def __enter__(self):
        cur = _get_device() if _get_device is not None else torch.cuda.current_device()
        idx = cur if self.idx is None else self.idx
        if cur != idx:
            self.prev = cur
            torch.cuda.set_device(idx)
        if _raw_stream is not None:
            return _raw_stream(idx)
        return torch.cuda.current_stream().cuda_stream

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)
        return False


def _check_out(t, shape, like, name):
    if tuple(t.shape) != tuple(shape) or t.dtype != like.dtype or t.device != like.device or not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous {like.dtype} tensor of shape {tuple(shape)} on {like.device}")
    return t


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step=64, out=None):
    """CUDA forward: (N,S,M,D), (L,2) i32, (L,) i32, (N,Lq,M,L,P,2), (N,Lq,M,L,P) -> (N, Lq, M*D).

    ``out`` (optional) is a caller-provided result buffer, e.g. a view into a packed staging arena."""
    if (out is None and _torch_ops.using_shim() and value.is_cuda and sampling_loc.is_cuda and attn_weight.is_cuda
            and spatial_shapes.is_cuda and level_start_index.is_cuda):
        # C++ entry: same checks and messages, no ctypes (csrc/msda_torch_shim.cpp)
        return torch.ops.alonet_custom.ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                                                              im2col_step)
    dims = _common_checks(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step)
    shapes = _meta_i32(spatial_shapes, "spatial_shapes")
    start = _meta_i32(level_start_index, "level_start_index")
    oshape = (dims.batch, dims.num_query, dims.num_heads * dims.channels)
    out = torch.empty(oshape, dtype=value.dtype, device=value.device) if out is None else _check_out(out, oshape, value, "out")
    L = _capi.lib()
    dt = _io_dtype(value, sampling_loc, attn_weight)
    # scheduling words for encoder-sized calls (msda_forward_ws): 0 bytes -- and no allocation -- for everything else
    ws_bytes = L.msda_forward_workspace_bytes(ctypes.byref(dims), dt)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=value.device) if ws_bytes else None
    with _on_device(value) as stream:
        rc = L.msda_forward_ws(_ptr(value), _ptr(shapes), _ptr(start), _ptr(sampling_loc), _ptr(attn_weight), _ptr(out),
                               _ptr(ws) if ws is not None else None, ws_bytes, ctypes.byref(dims), dt, stream)
    if rc != 0:
        raise RuntimeError("msda_forward failed: " + _capi.last_error())
    return out


class BackwardZeroFill:
    """Handle of an early zero-fill of backward's accumulation buffer (``begin_backward_zero_fill``)."""

    __slots__ = ("buffer", "event", "shape", "dtype")

    def __init__(self, buffer, event, shape, dtype):
        self.buffer, self.event, self.shape, self.dtype = buffer, event, shape, dtype


_side_streams = {}


def _side_stream(device):
    key = device.index if device.index is not None else torch.cuda.current_device()
    s = _side_streams.get(key)
    if s is None:
        s = _side_streams[key] = torch.cuda.Stream(device=device)
    return s


def begin_backward_zero_fill(value):
    """Enqueue the zero-fill that ``ms_deform_attn_backward`` starts with -- grad_value (fp32 / fp64) or its fp32 accumulation
    image (bf16 / fp16) -- on a side stream NOW, so that it overlaps whatever runs between this call and the backward pass
    (at least the forward kernel; in a training step the rest of the network).  Pass the handle as ``prezeroed=`` to
    ``ms_deform_attn_backward``, which waits for the fill and skips its own (C ABI: msda_zero_fill + MSDA_BWD_PREZEROED).
    The reference allocates and zero-fills inside the backward op (ms_deform_attn_cuda.cu:121).  Works under CUDA-graph
    capture (the side stream forks from and joins the capturing stream) provided the backward call is captured too."""
    if not value.is_cuda:
        raise RuntimeError("value must be a CUDA tensor")
    dev = value.device
    acc_dtype = value.dtype if value.dtype in (torch.float32, torch.float64) else torch.float32
    side = _side_stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        buf = torch.empty(value.shape, dtype=acc_dtype, device=dev)
        if buf.numel():
            rc = _capi.lib().msda_zero_fill(buf.data_ptr(), buf.numel() * buf.element_size(), side.cuda_stream)
            if rc != 0:
                raise RuntimeError("msda_zero_fill failed: " + _capi.last_error())
        ev = torch.cuda.Event()
        ev.record(side)
    return BackwardZeroFill(buf, ev, tuple(value.shape), value.dtype)


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                            im2col_step=64, grads=None, prezeroed=None, deterministic=None):
    """CUDA backward -> [grad_value, grad_sampling_loc, grad_attn_weight] (ms_deform_attn_cuda.cu:83-153).

    ``deterministic`` (None = follow MSDA_DETERMINISTIC / torch.use_deterministic_algorithms): bit-reproducible grad_value
    through 64-bit fixed-point accumulation (include/msda_b200.h, MSDA_BWD_DETERMINISTIC); float32 / bfloat16 / float16 with
    D in {16, 32, 64, 128}; float64 and other channel counts raise.

    ``grads`` (optional): three caller-provided result buffers shaped like value / sampling_loc / attn_weight.
    ``prezeroed`` (optional): handle from ``begin_backward_zero_fill(value)``; its buffer becomes grad_value (fp32 / fp64) or
    the accumulation workspace (16-bit types) and the library's own zero-fill is skipped."""
    if (grads is None and prezeroed is None and deterministic is None and _torch_ops.using_shim() and value.is_cuda
            and sampling_loc.is_cuda and attn_weight.is_cuda and spatial_shapes.is_cuda and level_start_index.is_cuda
            and grad_output.is_cuda):
        # C++ entry (csrc/msda_torch_shim.cpp): same checks, same C call, no ctypes; it honours an implicit deterministic request
        return torch.ops.alonet_custom.ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                                                               grad_output, im2col_step)
    dims = _common_checks(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step,
                          extra=(("grad_output", grad_output),))
    if grad_output.numel() != dims.batch * dims.num_query * dims.num_heads * dims.channels:
        raise RuntimeError(f"grad_output {tuple(grad_output.shape)} does not match the forward output")
    shapes = _meta_i32(spatial_shapes, "spatial_shapes")
    start = _meta_i32(level_start_index, "level_start_index")
    L = _capi.lib()
    dt = _io_dtype(value, sampling_loc, attn_weight)
    det = deterministic_requested() if deterministic is None else bool(deterministic)
    if det and deterministic is None and (value.dtype == torch.float64 or dims.channels not in (16, 32, 64, 128)):
        det = False  # (an implicit request must not break float64 gradcheck / odd channel counts: they keep the reds)
    flags, pre_gv, ws = (_capi.BWD_DETERMINISTIC if det else 0), None, None
    if det and prezeroed is not None:
        if deterministic is None:
            det, flags = False, 0  # an implicit request gives way to the caller's explicit early zero-fill
        else:
            raise RuntimeError("deterministic=True cannot be combined with prezeroed=")
    ws_bytes = L.msda_backward_workspace_bytes_ex(ctypes.byref(dims), dt, flags)
    if prezeroed is not None:
        if prezeroed.shape != tuple(value.shape) or prezeroed.dtype != value.dtype or prezeroed.buffer.device != value.device:
            raise RuntimeError("prezeroed handle was made for a different value tensor")
        cur = torch.cuda.current_stream(value.device)
        cur.wait_event(prezeroed.event)
        prezeroed.buffer.record_stream(cur)  # allocated on the side stream, used (and possibly freed) on this one
        flags |= _capi.BWD_PREZEROED
        if ws_bytes:
            ws = prezeroed.buffer
        else:
            pre_gv = prezeroed.buffer
    if grads is None:
        grad_value = pre_gv if pre_gv is not None else torch.empty_like(value)
        grad_loc, grad_attn = torch.empty_like(sampling_loc), torch.empty_like(attn_weight)
    else:
        if pre_gv is not None:
            raise RuntimeError("grads= and prezeroed= both name a grad_value buffer")
        grad_value = _check_out(grads[0], value.shape, value, "grads[0]")
        grad_loc = _check_out(grads[1], sampling_loc.shape, sampling_loc, "grads[1]")
        grad_attn = _check_out(grads[2], attn_weight.shape, attn_weight, "grads[2]")
    if ws is None and ws_bytes:
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=value.device)
    with _on_device(value) as stream:
        rc = L.msda_backward(_ptr(grad_output), _ptr(value), _ptr(shapes), _ptr(start), _ptr(sampling_loc),
                             _ptr(attn_weight), _ptr(grad_value), _ptr(grad_loc), _ptr(grad_attn),
                             _ptr(ws) if ws is not None else None, ws_bytes, ctypes.byref(dims), dt, flags,
                             stream)
    if rc != 0:
        raise RuntimeError("msda_backward failed: " + _capi.last_error())
    return [grad_value, grad_loc, grad_attn]


# ------------------------------------------------------------------------------------------------------------
# fused operator: softmax + sampling-location arithmetic of MSDeformAttn.forward inside the kernels
# (reference ops/modules/ms_deform_attn.py:121-137 + the op)
# ------------------------------------------------------------------------------------------------------------
def _fused_dims(value, spatial_shapes, reference_points, sampling_offsets, attn_logits):
    if sampling_offsets.dim() != 6 or sampling_offsets.shape[-1] != 2:
        raise RuntimeError(f"sampling_offsets must be (N, Lq, M, L, P, 2), got {tuple(sampling_offsets.shape)}")
    N, S, M, D = value.shape
    _, Lq, M2, L, P, _ = sampling_offsets.shape
    if M2 != M or spatial_shapes.shape[0] != L:
        raise RuntimeError("sampling_offsets inconsistent with value / spatial_shapes")
    if attn_logits.numel() != N * Lq * M * L * P:
        raise RuntimeError(f"attn_logits must hold N*Lq*M*L*P = {N * Lq * M * L * P} elements")
    if reference_points.dim() != 4 or tuple(reference_points.shape[:3]) != (N, Lq, L) or reference_points.shape[-1] not in (2, 4):
        raise ValueError(
            "Last dim of reference_points must be 2 or 4, but get {} instead.".format(reference_points.shape[-1]))
    return _capi.MsdaDims(N, S, M, D, L, Lq, P), int(reference_points.shape[-1])


def fused_supported(value, spatial_shapes, reference_points, sampling_offsets, attn_logits) -> bool:
    """True when the fused kernels serve this problem (else compose softmax / location arithmetic + the plain op)."""
    if not value.is_cuda or value.dtype not in (torch.float32, torch.bfloat16, torch.float16):
        return False
    dims, rd = _fused_dims(value, spatial_shapes, reference_points, sampling_offsets, attn_logits)
    return bool(_capi.lib().msda_fused_supported(ctypes.byref(dims), _DTYPES[value.dtype], rd))


def _ref_flag(value, reference_points):
    """fp32 reference points next to 16-bit value / offsets / logits (torch.autocast): MSDA_FUSED_REF_F32."""
    if reference_points.dtype == value.dtype:
        return 0
    if reference_points.dtype == torch.float32 and value.dtype in (torch.bfloat16, torch.float16):
        return _capi.FUSED_REF_F32
    raise RuntimeError(f"reference_points must have value's dtype ({value.dtype}) or be float32 next to 16-bit tensors, "
                       f"got {reference_points.dtype}")


def ms_deform_attn_fused_forward(value, spatial_shapes, level_start_index, reference_points, sampling_offsets,
                                 attn_logits):
    """(N,S,M,D), (L,2), (L,), (N,Lq,L,2|4), raw offsets (N,Lq,M,L,P,2), raw logits (N,Lq,M,L*P) -> (N, Lq, M*D).

    ``reference_points`` may stay float32 when the other tensors are bfloat16 / float16 (include/msda_b200.h,
    MSDA_FUSED_REF_F32): the sampling locations are then computed from the exact points."""
    named = [("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
             ("reference_points", reference_points), ("sampling_offsets", sampling_offsets), ("attn_logits", attn_logits)]
    _check_inputs(named)
    for name, t in named[4:]:
        if t.dtype != value.dtype or t.device != value.device:
            raise RuntimeError(f"{name} must have value's dtype and device")
    if reference_points.device != value.device:
        raise RuntimeError("reference_points must be on value's device")
    flags = _ref_flag(value, reference_points)
    dims, rd = _fused_dims(value, spatial_shapes, reference_points, sampling_offsets, attn_logits)
    shapes = _meta_i32(spatial_shapes, "spatial_shapes")
    start = _meta_i32(level_start_index, "level_start_index")
    out = torch.empty((dims.batch, dims.num_query, dims.num_heads * dims.channels), dtype=value.dtype, device=value.device)
    with _on_device(value) as stream:
        rc = _capi.lib().msda_fused_forward_ex(_ptr(value), _ptr(shapes), _ptr(start), _ptr(reference_points), rd,
                                               _ptr(sampling_offsets), _ptr(attn_logits), _ptr(out), ctypes.byref(dims),
                                               _DTYPES[value.dtype], flags, stream)
    if rc != 0:
        raise RuntimeError("msda_fused_forward failed: " + _capi.last_error())
    return out


def ms_deform_attn_fused_backward(value, spatial_shapes, level_start_index, reference_points, sampling_offsets,
                                  attn_logits, grad_output, need_grad_ref=True):
    """-> (grad_value, grad_reference_points | None, grad_sampling_offsets, grad_attn_logits)."""
    named = [("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
             ("reference_points", reference_points), ("sampling_offsets", sampling_offsets), ("attn_logits", attn_logits),
             ("grad_output", grad_output)]
    _check_inputs(named)
    flags = _ref_flag(value, reference_points)
    dims, rd = _fused_dims(value, spatial_shapes, reference_points, sampling_offsets, attn_logits)
    shapes = _meta_i32(spatial_shapes, "spatial_shapes")
    start = _meta_i32(level_start_index, "level_start_index")
    grad_value = torch.empty_like(value)
    grad_off = torch.empty_like(sampling_offsets)
    grad_logits = torch.empty_like(attn_logits)
    grad_ref = torch.zeros(reference_points.shape, dtype=torch.float32, device=value.device) if need_grad_ref else None
    L = _capi.lib()
    dt = _DTYPES[value.dtype]
    ws_bytes = L.msda_backward_workspace_bytes(ctypes.byref(dims), dt)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=value.device) if ws_bytes else None
    with _on_device(value) as stream:
        rc = L.msda_fused_backward(_ptr(grad_output), _ptr(value), _ptr(shapes), _ptr(start), _ptr(reference_points), rd,
                                   _ptr(sampling_offsets), _ptr(attn_logits), _ptr(grad_value), _ptr(grad_off),
                                   _ptr(grad_logits), _ptr(grad_ref) if grad_ref is not None else None,
                                   _ptr(ws) if ws is not None else None, ws_bytes, ctypes.byref(dims), dt, flags,
                                   stream)
    if rc != 0:
        raise RuntimeError("msda_fused_backward failed: " + _capi.last_error())
    if grad_ref is not None and grad_ref.dtype != reference_points.dtype:
        grad_ref = grad_ref.to(reference_points.dtype)
    return grad_value, grad_ref, grad_off, grad_logits


class MSDeformAttnFusedFunction(Function):
    """``apply(value, spatial_shapes, level_start_index, reference_points, sampling_offsets, attn_logits)``
    == softmax + location arithmetic + ``MSDeformAttnFunction`` of the reference module, in one kernel per pass."""

    @staticmethod
    def forward(ctx, value, spatial_shapes, level_start_index, reference_points, sampling_offsets, attn_logits):
        out = ms_deform_attn_fused_forward(value, spatial_shapes, level_start_index, reference_points, sampling_offsets,
                                           attn_logits)
        ctx.save_for_backward(value, spatial_shapes, level_start_index, reference_points, sampling_offsets, attn_logits)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, start, ref, off, logits = ctx.saved_tensors
        gv, gref, goff, glog = ms_deform_attn_fused_backward(value, shapes, start, ref, off, logits,
                                                             grad_output.contiguous(), need_grad_ref=ctx.needs_input_grad[3])
        return gv, None, None, gref, goff, glog


def load_ops():
    """Reference: torch.ops.load_library(<build dir>/MultiScaleDeformableAttention.so).  Here: load the C-ABI
    library and make sure ``torch.ops.alonet_custom.ms_deform_attn_{forward,backward}`` exist."""
    from . import torch_ops

    _capi.lib()
    torch_ops.register()


def load_MultiScaleDeformableAttention():
    """Must be called once before using MSDeformAttnFunction (same contract as the reference); idempotent and
    cheap.  Unlike the reference it never shells out to a build script: a missing library raises."""
    load_ops()


class MSDeformAttnFunction(Function):
    """``apply(value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
    im2col_step)`` -> (N, Lq, M*D); gradients for arguments 0, 3 and 4 only."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
                im2col_step):
        ctx.im2col_step = im2col_step
        # the backward's zero-fill starts now, on a side stream, when a backward can follow (knob: EARLY_ZERO_FILL); not
        # under CUDA-graph capture, where the side stream could stay unjoined if the backward is not captured
        ctx.prezeroed = None
        if (EARLY_ZERO_FILL and not deterministic_requested() and value.is_cuda and any(ctx.needs_input_grad) and value.numel()
                and not torch.cuda.is_current_stream_capturing()):
            ctx.prezeroed = begin_backward_zero_fill(value)
        output = torch.ops.alonet_custom.ms_deform_attn_forward(
            value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
            ctx.im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                              attention_weights)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights = ctx.saved_tensors
        if ctx.prezeroed is not None:
            pre, ctx.prezeroed = ctx.prezeroed, None
            grad_value, grad_sampling_loc, grad_attn_weight = ms_deform_attn_backward(
                value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
                grad_output.contiguous(), ctx.im2col_step, prezeroed=pre)
            return grad_value, None, None, grad_sampling_loc, grad_attn_weight, None
        grad_value, grad_sampling_loc, grad_attn_weight = torch.ops.alonet_custom.ms_deform_attn_backward(
            value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
            grad_output.contiguous(), ctx.im2col_step)
        return grad_value, None, None, grad_sampling_loc, grad_attn_weight, None


def _bilinear_sample_gather(img, grid):
    """Bilinear sampling with zero padding, ``align_corners=False``, from gather + arithmetic only.

    img (B, C, H, W); grid (B, Hg, Wg, 2) in [-1, 1], (x, y) order -> (B, C, Hg, Wg).  Same function as
    ``F.grid_sample(img, grid, "bilinear", "zeros", False)``; the reference spells it out by hand for the same reason
    (ms_deform_attn_func.py:99-102, 110-190): a ``GridSample`` node needs ONNX opset >= 16 and is not understood by its
    TensorRT tool chain, and padding an N-d tensor is replaced by concatenation ("TRT < 8.5.1", :153-160)."""
    B, C, H, W = img.shape
    _, Hg, Wg, _ = grid.shape
    x = ((grid[..., 0] + 1) * W - 1) / 2  # pixel coordinates, pixel centres at integers
    y = ((grid[..., 1] + 1) * H - 1) / 2
    x = x.reshape(B, -1)
    y = y.reshape(B, -1)
    x_lo = torch.floor(x)
    y_lo = torch.floor(y)
    fx = x - x_lo
    fy = y - y_lo
    # one ring of zeros around the image (concatenations, not F.pad): index 0 and H+1 / W+1 are the zero border
    zc = img.new_zeros((B, C, H, 1))
    ring = torch.cat((zc, img, zc), dim=3)
    zr = img.new_zeros((B, C, 1, W + 2))
    ring = torch.cat((zr, ring, zr), dim=2).reshape(B, C, (H + 2) * (W + 2))
    xi0 = (x_lo.long() + 1).clamp(0, W + 1)
    xi1 = (x_lo.long() + 2).clamp(0, W + 1)
    yi0 = (y_lo.long() + 1).clamp(0, H + 1)
    yi1 = (y_lo.long() + 2).clamp(0, H + 1)

    def tap(yi, xi):
        return torch.gather(ring, 2, (yi * (W + 2) + xi).unsqueeze(1).expand(-1, C, -1))

    w00 = ((1 - fx) * (1 - fy)).unsqueeze(1)
    w01 = (fx * (1 - fy)).unsqueeze(1)
    w10 = ((1 - fx) * fy).unsqueeze(1)
    w11 = (fx * fy).unsqueeze(1)
    out = tap(yi0, xi0) * w00 + tap(yi0, xi1) * w01 + tap(yi1, xi0) * w10 + tap(yi1, xi1) * w11
    return out.reshape(B, C, Hg, Wg)


def ms_deform_attn_core_pytorch(value, value_spatial_shapes, sampling_locations, attention_weights, use_grid_sample=False):
    """Pure-PyTorch formulation for tracing / ONNX / TensorRT export ONLY (the ``is_tracing`` branch of
    ``MSDeformAttn.forward``, reference ops/modules/ms_deform_attn.py:138-144).  It is not a fallback of the CUDA
    operator: nothing in this package routes to it unless the caller asks for the traceable graph.

    Like the reference (ms_deform_attn_func.py:85-107) the exported graph consists of split / gather / arithmetic nodes --
    no ``GridSample``; ``use_grid_sample=True`` swaps in ``F.grid_sample`` (the same function, for exporters that have it)."""
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = sampling_locations.shape
    sizes = [(int(h), int(w)) for h, w in value_spatial_shapes]
    levels = value.split([h * w for h, w in sizes], dim=1)
    grids = 2 * sampling_locations - 1
    per_level = []
    for lvl, (h, w) in enumerate(sizes):
        img = levels[lvl].flatten(2).transpose(1, 2).reshape(N * M, D, h, w)
        grid = grids[:, :, :, lvl].transpose(1, 2).flatten(0, 1)
        if use_grid_sample:
            per_level.append(F.grid_sample(img, grid, mode="bilinear", padding_mode="zeros", align_corners=False))
        else:
            per_level.append(_bilinear_sample_gather(img, grid))
    weights = attention_weights.transpose(1, 2).reshape(N * M, 1, Lq, L * P)
    out = (torch.stack(per_level, dim=-2).flatten(-2) * weights).sum(-1).view(N, M * D, Lq)
    return out.transpose(1, 2).contiguous()
