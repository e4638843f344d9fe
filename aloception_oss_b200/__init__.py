"""B200-native multi-scale deformable attention, drop-in for aloception's ``alonet.deformable_detr.ops``.

Public surface (same names as alonet/deformable_detr/ops/functions/__init__.py:9-14 and
ops/modules/__init__.py:9):

    MSDeformAttnFunction, ms_deform_attn_core_pytorch, load_MultiScaleDeformableAttention, load_ops, MSDeformAttn
"""
from .functions import (  # noqa: F401
    MSDeformAttnFunction,
    MSDeformAttnFusedFunction,
    begin_backward_zero_fill,
    fused_supported,
    ms_deform_attn_fused_backward,
    ms_deform_attn_fused_forward,
    load_MultiScaleDeformableAttention,
    load_ops,
    ms_deform_attn_backward,
    ms_deform_attn_core_pytorch,
    ms_deform_attn_forward,
)
from .modules import MSDeformAttn  # noqa: F401
from . import transformer  # noqa: F401  (encoder / decoder layer loop around the operator)

__all__ = [
    "MSDeformAttnFunction", "ms_deform_attn_core_pytorch", "load_MultiScaleDeformableAttention", "load_ops",
    "MSDeformAttn", "ms_deform_attn_forward", "ms_deform_attn_backward", "MSDeformAttnFusedFunction",
    "ms_deform_attn_fused_forward", "ms_deform_attn_fused_backward", "fused_supported", "begin_backward_zero_fill",
]
__version__ = "0.1.0"
