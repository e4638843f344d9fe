"""``MSDeformAttn`` -- mirror of alonet/deformable_detr/ops/modules/ms_deform_attn.py:34-155.

Same constructor, sub-module names (=> identical ``state_dict`` keys: ``sampling_offsets.*``,
``attention_weights.*``, ``value_proj.*``, ``output_proj.*``), ``im2col_step`` attribute, ``_reset_parameters``
initialisation and ``forward`` signature (including the ``is_tracing`` kwarg that selects the traceable
pure-PyTorch graph for export).  The sampling itself runs in the sm_100a operator.
"""
from __future__ import annotations

import math
import warnings
import weakref

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn.init import constant_, xavier_uniform_

from .functions import (MSDeformAttnFunction, MSDeformAttnFusedFunction, deterministic_requested, fused_supported,
                        load_MultiScaleDeformableAttention, ms_deform_attn_core_pytorch)


def _graph_is_being_recorded() -> bool:
    """True while torch.jit.trace / torch.onnx.export / torch.compile records the module: the sampling must then appear in
    the graph as the ``alonet_custom::ms_deform_attn_forward`` node the reference's exporter looks for
    (alonet/torch2trt/trt_exporter.py:41), not as an opaque ctypes call."""
    if torch.jit.is_tracing() or torch.onnx.is_in_onnx_export():
        return True
    comp = getattr(torch, "compiler", None)
    return bool(comp is not None and hasattr(comp, "is_compiling") and comp.is_compiling())


# shapes tensors whose ``sum(H*W) == Len_in`` check has passed: id(tensor) -> (weak reference, version counter, Len_in).
# Keyed on the LIVE tensor object (not on its address: the caching allocator hands a freed block to the next, different,
# shapes tensor) and shared by all modules, so a forward pass of 12 layers on one shapes tensor validates -- and syncs -- once.
_VALIDATED_LEVELS = {}


def _is_power_of_2(n):
    if (not isinstance(n, int)) or (n < 0):
        raise ValueError("invalid input for _is_power_of_2: {} (type: {})".format(n, type(n)))
    return (n & (n - 1) == 0) and n != 0


class MSDeformAttn(nn.Module):
    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4, fused=True):
        """Same arguments as the reference.  ``fused`` (extension, default on): softmax and the sampling-location
        arithmetic run inside the sampling kernels instead of ~5 elementwise kernels (same maths, same roundings of
        the locations).  ``fused=False`` reproduces the reference's op sequence."""
        super().__init__()
        self.fused = fused
        if d_model % n_heads != 0:
            raise ValueError("d_model must be divisible by n_heads, but got {} and {}".format(d_model, n_heads))
        if not _is_power_of_2(d_model // n_heads):
            warnings.warn("d_model // n_heads is not a power of 2: the operator falls back from its 128-bit vector "
                          "kernels to the shape-generic ones.")
        self.im2col_step = 64
        self.d_model = d_model
        self.n_levels = n_levels
        self.n_heads = n_heads
        self.n_points = n_points

        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)
        self._reset_parameters()
        load_MultiScaleDeformableAttention()

    def _check_level_sizes(self, input_spatial_shapes, Len_in):
        """The reference asserts ``sum_l H_l*W_l == Len_in`` on every call (ms_deform_attn.py:113), which costs a
        device->host sync per layer when the shapes live on the GPU.  Same check here, but remembered per LIVE shapes tensor
        (object identity + version counter; a new tensor is always checked, wherever the allocator put it) and skipped
        while a CUDA graph is being captured (a sync is illegal there)."""
        key = id(input_spatial_shapes)
        hit = _VALIDATED_LEVELS.get(key)
        if hit is not None and hit[0]() is input_spatial_shapes and hit[1] == input_spatial_shapes._version and hit[2] == int(Len_in):
            return
        if input_spatial_shapes.is_cuda and torch.cuda.is_current_stream_capturing():
            return
        assert (input_spatial_shapes[:, 0] * input_spatial_shapes[:, 1]).sum() == Len_in
        if len(_VALIDATED_LEVELS) > 256:
            _VALIDATED_LEVELS.clear()
        try:
            ref = weakref.ref(input_spatial_shapes, lambda _r, k=key: _VALIDATED_LEVELS.pop(k, None))
        except TypeError:
            return
        _VALIDATED_LEVELS[key] = (ref, input_spatial_shapes._version, int(Len_in))

    def _reset_parameters(self):
        # compass-pattern offset bias, zero attention logits, xavier projections (ms_deform_attn.py:70-88)
        constant_(self.sampling_offsets.weight.data, 0.0)
        angle = torch.arange(self.n_heads, dtype=torch.float32) * (2.0 * math.pi / self.n_heads)
        direction = torch.stack([angle.cos(), angle.sin()], -1)
        direction = direction / direction.abs().max(-1, keepdim=True)[0]
        bias = direction.view(self.n_heads, 1, 1, 2).repeat(1, self.n_levels, self.n_points, 1)
        bias = bias * torch.arange(1, self.n_points + 1, dtype=torch.float32).view(1, 1, self.n_points, 1)
        with torch.no_grad():
            self.sampling_offsets.bias = nn.Parameter(bias.reshape(-1))
        constant_(self.attention_weights.weight.data, 0.0)
        constant_(self.attention_weights.bias.data, 0.0)
        xavier_uniform_(self.value_proj.weight.data)
        constant_(self.value_proj.bias.data, 0.0)
        xavier_uniform_(self.output_proj.weight.data)
        constant_(self.output_proj.bias.data, 0.0)

    def forward(self, query, reference_points, input_flatten, input_spatial_shapes, input_level_start_index,
                input_padding_mask=None, **kwargs):
        """query (N, Lq, C); reference_points (N, Lq, L, 2|4) in [0,1]; input_flatten (N, S, C);
        input_spatial_shapes (L, 2); input_level_start_index (L,); input_padding_mask (N, S) True=padding
        -> (N, Lq, C)."""
        N, Len_q, _ = query.shape
        N, Len_in, _ = input_flatten.shape
        self._check_level_sizes(input_spatial_shapes, Len_in)

        value = self.value_proj(input_flatten)
        if input_padding_mask is not None:
            value = value.masked_fill(input_padding_mask[..., None], float(0))
        value = value.view(N, Len_in, self.n_heads, self.d_model // self.n_heads)
        offsets = self.sampling_offsets(query).view(N, Len_q, self.n_heads, self.n_levels, self.n_points, 2)
        weights = self.attention_weights(query).view(N, Len_q, self.n_heads, self.n_levels * self.n_points)
        if reference_points.shape[-1] not in (2, 4):
            raise ValueError(
                "Last dim of reference_points must be 2 or 4, but get {} instead.".format(reference_points.shape[-1]))
        if reference_points.dtype != value.dtype and "is_tracing" not in kwargs:
            # mixed precision (torch.autocast runs the Linear layers in bf16/fp16 while the reference points stay fp32): the
            # operator takes ONE storage dtype for value / offsets / logits -- value's -- and does its arithmetic in fp32
            offsets, weights = offsets.to(value.dtype), weights.to(value.dtype)
        # (while a graph is being recorded the unfused op sequence runs: it goes through torch.ops.alonet_custom.*)
        # (... and when a deterministic backward is requested -- torch.use_deterministic_algorithms / MSDA_DETERMINISTIC --, because
        # the fused backward accumulates grad_value and the reference-point gradient with fp32 atomics; the plain operator has
        # the fixed-point mode, MSDA_BWD_DETERMINISTIC)
        if (self.fused and "is_tracing" not in kwargs and not _graph_is_being_recorded()
                and not (torch.is_grad_enabled() and deterministic_requested())
                and (reference_points.dtype == value.dtype or reference_points.dtype == torch.float32)
                and fused_supported(value, input_spatial_shapes, reference_points, offsets, weights)):
            # fp32 reference points stay fp32 next to 16-bit tensors (MSDA_FUSED_REF_F32): a bf16 reference point would be
            # off by up to 1/256 of the image, i.e. 0.4-0.8 px on a 100-200 px level
            output = MSDeformAttnFusedFunction.apply(value.contiguous(), input_spatial_shapes, input_level_start_index,
                                                     reference_points.contiguous(), offsets.contiguous(),
                                                     weights.contiguous())
            return self.output_proj(output)
        if reference_points.dtype != value.dtype and "is_tracing" not in kwargs:
            # unfused path: the location arithmetic below runs in fp32 (type promotion with the fp32 points) and the operator
            # takes the fp32 locations as they are next to 16-bit value (MSDA_LOC_F32)
            offsets = offsets.to(reference_points.dtype)
        weights = F.softmax(weights, -1).view(N, Len_q, self.n_heads, self.n_levels, self.n_points)
        if reference_points.shape[-1] == 2:
            normalizer = torch.stack([input_spatial_shapes[..., 1], input_spatial_shapes[..., 0]], -1)
            locations = reference_points[:, :, None, :, None, :] + offsets / normalizer[None, None, None, :, None, :]
        elif reference_points.shape[-1] == 4:
            locations = (reference_points[:, :, None, :, None, :2]
                         + offsets / self.n_points * reference_points[:, :, None, :, None, 2:] * 0.5)
        else:
            raise ValueError(
                "Last dim of reference_points must be 2 or 4, but get {} instead.".format(reference_points.shape[-1]))
        if "is_tracing" in kwargs:
            output = ms_deform_attn_core_pytorch(value, input_spatial_shapes, locations, weights)
        else:
            half = value.dtype in (torch.bfloat16, torch.float16)
            if not (half and locations.dtype == torch.float32):
                locations = locations.to(value.dtype)
            if not (half and weights.dtype == torch.float32):
                weights = weights.to(value.dtype)
            output = MSDeformAttnFunction.apply(value.contiguous(), input_spatial_shapes, input_level_start_index,
                                                locations.contiguous(), weights.contiguous(), self.im2col_step)
        return self.output_proj(output)
