"""Autograd bridge with the reference's surface (ops/functions/ms_deform_attn_func.py:32-49).

`MSDeformAttnFunction.apply(value, value_spatial_shapes, value_level_start_index,
sampling_locations, attention_weights, im2col_step)` — same argument order and meaning, same
saved tensors, backward returns `(grad_value, None, None, grad_sampling_loc, grad_attn_weight,
None)` and is once-differentiable.

Differences from the reference file, both deliberate:
  * the native module is the sm_100a build loaded from this tree (bm2f_b200.load_extension);
  * the pure-torch `ms_deform_attn_core_pytorch` is NOT here: it is a test oracle and lives in
    oracle/ (the product path has no CPU route).  When this tree is dropped into the reference,
    the reference keeps its own functions/ file and only the native module is replaced.
"""
from __future__ import annotations

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ... import load_extension

MSDA = load_extension()


class MSDeformAttnFunction(Function):
    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations,
                attention_weights, im2col_step):
        ctx.im2col_step = im2col_step
        output = MSDA.ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index,
                                             sampling_locations, attention_weights, ctx.im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                              attention_weights)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, start, loc, attn = ctx.saved_tensors
        grad_value, grad_loc, grad_attn = MSDA.ms_deform_attn_backward(
            value, shapes, start, loc, attn, grad_output.contiguous(), ctx.im2col_step)
        return grad_value, None, None, grad_loc, grad_attn, None


class MSDeformAttnFusedFunction(Function):
    """Fused variant used by bm2f_b200.ops.modules.MSDeformAttn (not part of the reference surface):
    `apply(value, value_spatial_shapes, value_level_start_index, reference_points, sampling_offsets,
    attention_logits)` where sampling_offsets / attention_logits are the raw outputs of the two Linear
    layers.  The softmax over the L*P points and `ref + offset / (W_l, H_l)`
    (ops/modules/ms_deform_attn.py:101-109) run inside the sampling kernels; backward returns the
    gradients of the raw Linear outputs.  reference_points (2-d form) receive no gradient."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, reference_points, sampling_offsets,
                attention_logits, value_padding_mask=None):
        output = MSDA.ms_deform_attn_fused_forward(value, value_spatial_shapes, value_level_start_index,
                                                   reference_points, sampling_offsets, attention_logits)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, reference_points,
                              sampling_offsets, attention_logits)
        ctx.value_padding_mask = value_padding_mask     # (N, S) bool or None: rows of `value` that were zero-filled
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, start, ref, off, logits = ctx.saved_tensors
        grad_value, grad_off, grad_logits = MSDA.ms_deform_attn_fused_backward(
            value, shapes, start, ref, off, logits, grad_output.contiguous())
        if ctx.value_padding_mask is not None and grad_value.dtype == torch.float32:
            # backward of value.masked_fill(mask[..., None], 0): only the masked rows of our own tensor are written
            MSDA.zero_masked_rows_(grad_value.view(-1, grad_value.shape[-2] * grad_value.shape[-1]),
                                   ctx.value_padding_mask)
        elif ctx.value_padding_mask is not None:
            grad_value = grad_value.masked_fill(ctx.value_padding_mask[..., None, None], 0)
        return grad_value, None, None, None, grad_off, grad_logits, None


class MSDeformAttnFusedPackedFunction(Function):
    """`MSDeformAttnFusedFunction` on the output of ONE offsets||logits projection: `offsets_logits` (N, Lq, M*L*P*3) holds
    per row the M*L*P*2 raw sampling offsets, then the M*L*P raw attention logits (ops/modules/ms_deform_attn.py:101-102
    with the two weight matrices stacked).  The kernels read the two column blocks through tensor maps whose row stride is
    the packed width and write their gradients into one tensor of the same layout, so the projection's input gradient is
    one GEMM too.  Bit-identical to the unpacked function."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, reference_points, offsets_logits, n_points,
                value_padding_mask=None):
        output = MSDA.ms_deform_attn_fused_forward_packed(value, value_spatial_shapes, value_level_start_index,
                                                          reference_points, offsets_logits, n_points)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, reference_points, offsets_logits)
        ctx.n_points = n_points
        ctx.value_padding_mask = value_padding_mask
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, start, ref, oa = ctx.saved_tensors
        grad_value, grad_oa = MSDA.ms_deform_attn_fused_backward_packed(value, shapes, start, ref, oa, ctx.n_points,
                                                                        grad_output.contiguous())
        if ctx.value_padding_mask is not None and grad_value.dtype == torch.float32:
            MSDA.zero_masked_rows_(grad_value.view(-1, grad_value.shape[-2] * grad_value.shape[-1]),
                                   ctx.value_padding_mask)
        elif ctx.value_padding_mask is not None:
            grad_value = grad_value.masked_fill(ctx.value_padding_mask[..., None, None], 0)
        return grad_value, None, None, None, grad_oa, None, None
