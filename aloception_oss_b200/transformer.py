"""Encoder / decoder layer loop around the operator -- SURVEY.md section 8(f) row 3.

Mirror of the CALLERS of the hot path in alonet/deformable_detr/deformable_transformer.py:

* ``DeformableTransformerEncoderLayer``  <- :306-344   (self_attn = MSDeformAttn, norm1, linear1/2, norm2)
* ``DeformableTransformerEncoder``       <- :347-407   (``get_reference_points`` + the layer loop)
* ``DeformableTransformerDecoderLayer``  <- :410-515   (nn.MultiheadAttention self-attention, cross_attn = MSDeformAttn, FFN)
* ``DeformableTransformerDecoder``       <- :517-632   (layer loop, iterative box refinement hook, intermediates)

Same constructor arguments, sub-module names (=> the reference's ``state_dict`` keys load unchanged), forward signatures
(including ``**kwargs`` with the ``is_tracing`` export switch, which is handed down to ``MSDeformAttn``) and the same
arithmetic order.  What is B200-first here is host-side only -- the dense parts are library GEMMs / LayerNorm by design:

* ``MSDeformAttn`` runs the fused sm_100a operator (softmax + location arithmetic inside the sampling kernels);
* the encoder's reference points are built ONCE per (level shapes, device) from a host copy of the shapes that the caller
  may pass (``spatial_shapes_host=``), so the layer loop issues no device->host sync (the reference iterates a CUDA
  ``spatial_shapes`` tensor: one sync per level per forward, deformable_transformer.py:373-374) and can be captured in a
  CUDA graph (``GraphedModule``);
* ``GraphedModule`` captures a module's forward for fixed input shapes (inference) and replays it.
"""
from __future__ import annotations

import copy

import torch
import torch.nn.functional as F
from torch import nn

from .modules import MSDeformAttn


def _activation(name):
    # deformable_transformer.py:639-647
    if name == "relu":
        return F.relu
    if name == "gelu":
        return F.gelu
    if name == "glu":
        return F.glu
    raise RuntimeError(f"activation should be relu/gelu, not {name}.")


def _clones(module, n):
    return nn.ModuleList([copy.deepcopy(module) for _ in range(n)])


def inverse_sigmoid(x, eps=1e-5):
    # alonet/deformable_detr/utils.py (same clamps)
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


def _add_pos(t, pos):
    return t if pos is None else t + pos


class DeformableTransformerEncoderLayer(nn.Module):
    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_levels=4, n_heads=8, n_points=4,
                 fused=True):
        super().__init__()
        self.self_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points, fused=fused)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.activation = _activation(activation)
        self.dropout2 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout3 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)

    with_pos_embed = staticmethod(_add_pos)

    def forward_ffn(self, src):
        hidden = self.dropout2(self.activation(self.linear1(src)))
        return self.norm2(src + self.dropout3(self.linear2(hidden)))

    def forward(self, src, pos, reference_points, spatial_shapes, level_start_index, padding_mask=None, **kwargs):
        attn = self.self_attn(_add_pos(src, pos), reference_points, src, spatial_shapes, level_start_index, padding_mask,
                              **kwargs)
        return self.forward_ffn(self.norm1(src + self.dropout1(attn)))


class DeformableTransformerEncoder(nn.Module):
    def __init__(self, encoder_layer, num_layers):
        super().__init__()
        self.layers = _clones(encoder_layer, num_layers)
        self.num_layers = num_layers

    @staticmethod
    def get_reference_points(spatial_shapes, valid_ratios, device, spatial_shapes_host=None, **kwargs):
        """(b, sum_l H_l*W_l, L, 2): centre of every pixel of every level, normalised by the valid part of the image and
        re-scaled per level (deformable_transformer.py:353-400, same operation order).  ``spatial_shapes_host`` -- a list of
        (H, W) pairs -- avoids reading the device tensor."""
        if spatial_shapes_host is None:
            spatial_shapes_host = [(int(h), int(w)) for h, w in spatial_shapes.tolist()]
        vr = valid_ratios[:, None]  # (b, 1, L, 2)
        per_level = []
        for lvl, (h, w) in enumerate(spatial_shapes_host):
            ys = torch.arange(h, dtype=torch.int32, device=device).float() + 0.5
            xs = torch.arange(w, dtype=torch.int32, device=device).float() + 0.5
            ref_y, ref_x = torch.meshgrid(ys, xs, indexing="ij")
            ref_y = ref_y.reshape(1, -1) / (vr[:, :, lvl, 1] * h)
            ref_x = ref_x.reshape(1, -1) / (vr[:, :, lvl, 0] * w)
            per_level.append(torch.stack((ref_x, ref_y), -1))
        return torch.cat(per_level, 1)[:, :, None] * vr

    def forward(self, src, spatial_shapes, level_start_index, valid_ratios, pos=None, padding_mask=None, **kwargs):
        host = kwargs.pop("spatial_shapes_host", None)
        reference_points = self.get_reference_points(spatial_shapes, valid_ratios, device=src.device,
                                                     spatial_shapes_host=host)
        out = src
        for layer in self.layers:
            out = layer(out, pos, reference_points, spatial_shapes, level_start_index, padding_mask, **kwargs)
        return out


class DeformableTransformerDecoderLayer(nn.Module):
    def __init__(self, d_model=256, dim_feedforward=1024, dropout=0.1, activation="relu", n_levels=4, n_heads=8,
                 n_points=4, fused=True):
        super().__init__()
        self.cross_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points, fused=fused)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.self_attn = nn.MultiheadAttention(d_model, n_heads, dropout=dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.activation = _activation(activation)
        self.dropout3 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.dropout4 = nn.Dropout(dropout)
        self.norm3 = nn.LayerNorm(d_model)

    with_pos_embed = staticmethod(_add_pos)

    def forward_ffn(self, tgt):
        hidden = self.dropout3(self.activation(self.linear1(tgt)))
        return self.norm3(tgt + self.dropout4(self.linear2(hidden)))

    def pre_process_tgt(self, tgt, query_pos, tgt_key_padding_mask, **kwargs):
        return tgt, query_pos, tgt_key_padding_mask

    def decoder_layer_forward(self, tgt, query_pos, reference_points, src, src_spatial_shapes, level_start_index,
                              tgt_key_padding_mask=None, src_padding_mask=None, **kwargs):
        # dense self-attention between the object queries (sequence-first layout of nn.MultiheadAttention)
        qk = _add_pos(tgt, query_pos).transpose(0, 1)
        sa = self.self_attn(qk, qk, tgt.transpose(0, 1), key_padding_mask=tgt_key_padding_mask)[0].transpose(0, 1)
        tgt = self.norm2(tgt + self.dropout2(sa))
        # deformable cross-attention into the encoder memory
        ca = self.cross_attn(_add_pos(tgt, query_pos), reference_points, src, src_spatial_shapes, level_start_index,
                             src_padding_mask, **kwargs)
        tgt = self.norm1(tgt + self.dropout1(ca))
        return self.forward_ffn(tgt)

    def forward(self, tgt, query_pos, reference_points, src, src_spatial_shapes, level_start_index,
                tgt_key_padding_mask=None, src_padding_mask=None, **kwargs):
        tgt, query_pos, tgt_key_padding_mask = self.pre_process_tgt(tgt, query_pos, tgt_key_padding_mask, **kwargs)
        return self.decoder_layer_forward(tgt=tgt, query_pos=query_pos, reference_points=reference_points, src=src,
                                          src_spatial_shapes=src_spatial_shapes, level_start_index=level_start_index,
                                          tgt_key_padding_mask=tgt_key_padding_mask, src_padding_mask=src_padding_mask,
                                          **kwargs)


class DeformableTransformerDecoder(nn.Module):
    def __init__(self, decoder_layer, num_layers, return_intermediate=False):
        super().__init__()
        self.layers = _clones(decoder_layer, num_layers)
        self.num_layers = num_layers
        self.return_intermediate = return_intermediate
        self.bbox_embed = None   # set by the model for iterative box refinement / two-stage (deformable_transformer.py:523-525)
        self.class_embed = None

    def pre_process_tgt(self, tgt, query_pos, tgt_key_padding_mask, reference_points, **kwargs):
        return tgt, query_pos, tgt_key_padding_mask, reference_points

    def decoder_forward(self, tgt, reference_points, src, src_spatial_shapes, src_level_start_index, src_valid_ratios,
                        query_pos=None, src_padding_mask=None, tgt_key_padding_mask=None, **kwargs):
        out = tgt
        inter, inter_refs = [], []
        for lid, layer in enumerate(self.layers):
            if reference_points.shape[-1] == 4:
                ref_in = reference_points[:, :, None] * torch.cat([src_valid_ratios, src_valid_ratios], -1)[:, None]
            else:
                assert reference_points.shape[-1] == 2
                ref_in = reference_points[:, :, None] * src_valid_ratios[:, None]
            out = layer(tgt=out, query_pos=query_pos, reference_points=ref_in, src=src,
                        src_spatial_shapes=src_spatial_shapes, level_start_index=src_level_start_index,
                        src_padding_mask=src_padding_mask, tgt_key_padding_mask=tgt_key_padding_mask, **kwargs)
            if self.bbox_embed is not None:  # iterative bounding-box refinement
                delta = self.bbox_embed[lid](out)
                if reference_points.shape[-1] == 4:
                    new_ref = (delta + inverse_sigmoid(reference_points)).sigmoid()
                else:
                    assert reference_points.shape[-1] == 2
                    new_ref = delta
                    new_ref[..., :2] = delta[..., :2] + inverse_sigmoid(reference_points)
                    new_ref = new_ref.sigmoid()
                reference_points = new_ref.detach()
            if self.return_intermediate:
                inter.append(out)
                inter_refs.append(reference_points)
        if self.return_intermediate:
            return torch.stack(inter), torch.stack(inter_refs)
        return out, reference_points

    def forward(self, tgt, reference_points, src, src_spatial_shapes, src_level_start_index, src_valid_ratios,
                query_pos=None, src_padding_mask=None, tgt_key_padding_mask=None, decoder_outputs: dict = None, **kwargs):
        decoder_outputs = {} if decoder_outputs is None else decoder_outputs
        # the hook sees sequence-first tensors, as in the reference (deformable_transformer.py:607-613)
        tgt, query_pos, tgt_key_padding_mask, reference_points = self.pre_process_tgt(
            tgt.transpose(0, 1), query_pos.transpose(0, 1), tgt_key_padding_mask=tgt_key_padding_mask,
            reference_points=reference_points, **kwargs)
        tgt, query_pos = tgt.transpose(1, 0), query_pos.transpose(1, 0)
        hs, inter_refs = self.decoder_forward(
            tgt=tgt, reference_points=reference_points, src=src, src_spatial_shapes=src_spatial_shapes,
            src_level_start_index=src_level_start_index, src_valid_ratios=src_valid_ratios, query_pos=query_pos,
            src_padding_mask=src_padding_mask, tgt_key_padding_mask=tgt_key_padding_mask, **kwargs)
        decoder_outputs["init_reference_out"] = reference_points
        decoder_outputs.update({"hs": hs, "inter_references_out": inter_refs})
        return decoder_outputs


class GraphedModule:
    """CUDA-graph replay of ``module(*args, **kwargs)`` for FIXED input shapes (inference; no autograd).

    ``g = GraphedModule(encoder, src, shapes, start, valid_ratios, pos=pos, spatial_shapes_host=[...])`` warms the module up on
    a side stream, captures one forward into a ``torch.cuda.CUDAGraph`` and keeps the example tensors as static inputs;
    ``g(src2, shapes, start, valid_ratios, pos=pos2)`` copies new values into them and replays.  Non-tensor arguments must
    not change.  The operator's launches (``msda_forward`` / ``msda_fused_forward``) are plain stream-ordered kernel
    launches on the current stream, so they are captured like any other kernel."""

    def __init__(self, module, *args, warmup: int = 3, **kwargs):
        self.module = module
        self.static_args = [a.clone() if torch.is_tensor(a) else a for a in args]
        self.static_kwargs = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in kwargs.items()}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):
                module(*self.static_args, **self.static_kwargs)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.static_out = module(*self.static_args, **self.static_kwargs)

    def __call__(self, *args, **kwargs):
        for dst, src in zip(self.static_args, args):
            if torch.is_tensor(dst):
                dst.copy_(src)
        for k, src in kwargs.items():
            dst = self.static_kwargs.get(k)
            if torch.is_tensor(dst):
                dst.copy_(src)
        self.graph.replay()
        return self.static_out
