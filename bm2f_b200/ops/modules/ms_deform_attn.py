"""`MSDeformAttn` with the reference's module surface (ops/modules/ms_deform_attn.py:34-125).

Same constructor signature and defaults (d_model=256, n_levels=4, n_heads=8, n_points=4), same
sub-module names and therefore the same state-dict keys
(`sampling_offsets.{weight,bias}`, `attention_weights.{weight,bias}`, `value_proj.{weight,bias}`,
`output_proj.{weight,bias}`), same parameter initialisation (ms_deform_attn.py:64-80: zero offset
weights, compass-direction offset bias scaled by point index, zero attention weights, xavier
projections), same `forward(query, reference_points, input_flatten, input_spatial_shapes,
input_level_start_index, input_padding_mask=None)` and `im2col_step = 128`.

Deliberate difference: the reference wraps the native call in a bare `try/except` and silently
falls back to a pure-torch CPU path on ANY error (ms_deform_attn.py:116-121).  Here errors
propagate and non-CUDA inputs raise: there is no CPU fallback.
"""
from __future__ import annotations

import math
import warnings

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn.init import constant_, xavier_uniform_

from ..functions import MSDeformAttnFunction, MSDeformAttnFusedFunction, MSDeformAttnFusedPackedFunction
from ..functions.ms_deform_attn_func import MSDA
from ..functions import linear_func


def _is_power_of_2(n):
    if (not isinstance(n, int)) or (n < 0):
        raise ValueError("invalid input for _is_power_of_2: {} (type: {})".format(n, type(n)))
    return (n & (n - 1) == 0) and n != 0


class MSDeformAttn(nn.Module):
    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        if d_model % n_heads != 0:
            raise ValueError("d_model must be divisible by n_heads, but got {} and {}".format(d_model, n_heads))
        head_dim = d_model // n_heads
        if not _is_power_of_2(head_dim):
            warnings.warn("MSDeformAttn: head dimension %d is not a power of 2; only head_dim == 32 runs on the "
                          "sm_100a fast path, other sizes use the generic kernel." % head_dim)

        self.im2col_step = 128
        # softmax + sampling-location arithmetic inside the sampling kernels (fp32 / bf16 values, 2-d
        # reference points, head_dim 32, 8 heads, 4 points, <= 4 levels); set False for the unfused op
        self.fuse_prologue = True
        # forward GEMMs of the four projections on tcgen05 (tf32x3, fp32-grade); False = torch / cuBLAS
        self.tcgen05_linear = True
        # encoder self-attention: sampling_offsets || attention_weights as ONE 256 -> 288 projection feeding the packed fused
        # op; False = two projections (192 and 96 wide), same results
        self.packed_projections = True

        self.d_model = d_model
        self.n_levels = n_levels
        self.n_heads = n_heads
        self.n_points = n_points

        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)

        self._reset_parameters()

    def _reset_parameters(self):
        constant_(self.sampling_offsets.weight.data, 0.)
        # head m looks along angle 2*pi*m/M (normalised to the unit square); point p sits p+1 pixels out
        angle = torch.arange(self.n_heads, dtype=torch.float32) * (2.0 * math.pi / self.n_heads)
        direction = torch.stack([angle.cos(), angle.sin()], -1)
        direction = direction / direction.abs().max(-1, keepdim=True)[0]
        reach = torch.arange(1, self.n_points + 1, dtype=torch.float32)
        bias = direction[:, None, None, :] * reach[None, None, :, None]
        bias = bias.expand(self.n_heads, self.n_levels, self.n_points, 2)
        with torch.no_grad():
            self.sampling_offsets.bias = nn.Parameter(bias.reshape(-1).clone())
        constant_(self.attention_weights.weight.data, 0.)
        constant_(self.attention_weights.bias.data, 0.)
        xavier_uniform_(self.value_proj.weight.data)
        constant_(self.value_proj.bias.data, 0.)
        xavier_uniform_(self.output_proj.weight.data)
        constant_(self.output_proj.bias.data, 0.)

    def _proj(self, layer, x, row_mask=None):
        if self.tcgen05_linear and linear_func.supported(layer, x):
            return linear_func.linear_tf32x3(x, layer.weight, layer.bias, row_mask=row_mask)
        y = layer(x)
        return y if row_mask is None else y.masked_fill(row_mask[..., None], float(0))

    def self_attention_supported(self, src, pos, reference_points):
        """True when `forward_self_attention` can take the call (fused prologue + tcgen05 projections, 2-d reference
        points, float32 CUDA tensors)."""
        return (self.fuse_prologue and self.tcgen05_linear and src.is_cuda and src.dtype == torch.float32
                and pos is not None and pos.dtype == torch.float32 and reference_points.shape[-1] == 2
                and reference_points.dtype == torch.float32
                and not (torch.is_grad_enabled() and reference_points.requires_grad)
                and all(linear_func.supported(l, src) for l in (self.value_proj, self.sampling_offsets, self.attention_weights))
                and bool(MSDA.ms_deform_attn_fused_supported(self.n_heads, self.d_model // self.n_heads, self.n_levels,
                                                             self.n_points, False)))

    def forward_self_attention(self, src, pos, reference_points, input_spatial_shapes, input_level_start_index,
                               input_padding_mask=None):
        """Encoder self-attention: `forward(src + pos, reference_points, src, ...)` (msdeformattn.py:123) with the three
        input projections as one autograd node (`SelfAttnProjectionsFunction`): the gradients of `src + pos` and `src`
        are accumulated in GEMM epilogues instead of element-wise passes.  Same result as `forward`."""
        N, Len, _ = src.shape
        analytic = getattr(reference_points, "pixel_centres", False)
        if self.packed_projections and self.n_heads * self.n_levels * self.n_points * 3 == 288:
            # ONE offsets||logits projection (256 -> 288) feeding the packed fused op: `query` is read once, its gradient
            # is one GEMM (ops/functions/linear_func.py:SelfAttnProjectionsPackedFunction)
            value, oa = linear_func.SelfAttnProjectionsPackedFunction.apply(
                src, pos, self.value_proj.weight, self.value_proj.bias, self.sampling_offsets.weight,
                self.sampling_offsets.bias, self.attention_weights.weight, self.attention_weights.bias,
                linear_func.matmul_split(), input_padding_mask)
            output = MSDeformAttnFusedPackedFunction.apply(
                value.view(N, Len, self.n_heads, self.d_model // self.n_heads), input_spatial_shapes,
                input_level_start_index, None if analytic else reference_points.contiguous(), oa, self.n_points,
                input_padding_mask)
            return self._proj(self.output_proj, output)
        value, offsets, logits = linear_func.SelfAttnProjectionsFunction.apply(
            src, pos, self.value_proj.weight, self.value_proj.bias, self.sampling_offsets.weight,
            self.sampling_offsets.bias, self.attention_weights.weight, self.attention_weights.bias,
            linear_func.matmul_split(), input_padding_mask)
        value = value.view(N, Len, self.n_heads, self.d_model // self.n_heads)
        offsets = offsets.view(N, Len, self.n_heads, self.n_levels, self.n_points, 2)
        logits = logits.view(N, Len, self.n_heads, self.n_levels * self.n_points)
        output = MSDeformAttnFusedFunction.apply(
            value, input_spatial_shapes, input_level_start_index, None if analytic else reference_points.contiguous(),
            offsets, logits, input_padding_mask)
        return self._proj(self.output_proj, output)

    def forward(self, query, reference_points, input_flatten, input_spatial_shapes, input_level_start_index,
                input_padding_mask=None):
        """
        query                    (N, Lq, C)
        reference_points         (N, Lq, n_levels, 2) in [0,1] (x, y), or (N, Lq, n_levels, 4) boxes (x, y, w, h)
        input_flatten            (N, sum_l H_l*W_l, C)
        input_spatial_shapes     (n_levels, 2) int64 [(H_0, W_0), ...]
        input_level_start_index  (n_levels,) int64
        input_padding_mask       (N, sum_l H_l*W_l) bool, True = padding
        returns                  (N, Lq, C)
        """
        N, Len_q, _ = query.shape
        N, Len_in, _ = input_flatten.shape
        if not query.is_cuda:
            raise RuntimeError("MSDeformAttn: Not implemented on the CPU (this build has no CPU fallback)")
        # the reference's check (ops/modules/ms_deform_attn.py:96).  Callers that built the table from a Python list
        # (bm2f_b200.encoder) attach it as `hw_list`, so the check costs no device->host sync; a bare tensor is checked
        # exactly like the reference does (one sync).  The kernels additionally treat a level that does not fit
        # inside `input_flatten` as empty, so a wrong table can never read or write out of bounds.
        hw_list = getattr(input_spatial_shapes, "hw_list", None)
        if hw_list is not None:
            assert sum(int(h) * int(w) for h, w in hw_list) == Len_in
        else:
            assert (input_spatial_shapes[:, 0] * input_spatial_shapes[:, 1]).sum() == Len_in

        # the fused kernels return no gradient for reference_points (Mask2Former builds them from the level shapes):
        # learnable / refined reference points take the unfused path, which propagates it through sampling_locations
        ref_needs_grad = torch.is_grad_enabled() and reference_points.requires_grad
        use_fused = (self.fuse_prologue and reference_points.shape[-1] == 2 and reference_points.dtype == torch.float32
                     and query.dtype == torch.float32 and not ref_needs_grad
                     and MSDA.ms_deform_attn_fused_supported(self.n_heads, self.d_model // self.n_heads, self.n_levels,
                                                             self.n_points, False))
        if use_fused:
            # masked rows are zero-filled by a row-sparse in-place kernel inside the projection (forward) and inside
            # the fused attention function (backward); both full-tensor masked_fill passes of the reference disappear
            value = self._proj(self.value_proj, input_flatten, row_mask=input_padding_mask)
        else:
            value = self._proj(self.value_proj, input_flatten)
            if input_padding_mask is not None:
                value = value.masked_fill(input_padding_mask[..., None], float(0))
        value = value.view(N, Len_in, self.n_heads, self.d_model // self.n_heads)
        sampling_offsets = self._proj(self.sampling_offsets, query).view(
            N, Len_q, self.n_heads, self.n_levels, self.n_points, 2)
        attention_weights = self._proj(self.attention_weights, query).view(
            N, Len_q, self.n_heads, self.n_levels * self.n_points)
        if use_fused and value.dtype == torch.float32 and sampling_offsets.dtype == torch.float32:
            # The encoder tags reference points it built itself with valid ratios 1 (`pixel_centres`): the kernel then
            # derives them from the query index (bit-identical) instead of loading them per (query, head).
            analytic = getattr(reference_points, "pixel_centres", False) and Len_q == Len_in
            output = MSDeformAttnFusedFunction.apply(
                value, input_spatial_shapes, input_level_start_index,
                None if analytic else reference_points.contiguous(), sampling_offsets, attention_weights,
                input_padding_mask)
            return self._proj(self.output_proj, output)
        attention_weights = F.softmax(attention_weights, -1).view(
            N, Len_q, self.n_heads, self.n_levels, self.n_points)
        if reference_points.shape[-1] == 2:
            offset_normalizer = torch.stack([input_spatial_shapes[..., 1], input_spatial_shapes[..., 0]], -1)
            sampling_locations = reference_points[:, :, None, :, None, :] \
                + sampling_offsets / offset_normalizer[None, None, None, :, None, :]
        elif reference_points.shape[-1] == 4:
            sampling_locations = reference_points[:, :, None, :, None, :2] \
                + sampling_offsets / self.n_points * reference_points[:, :, None, :, None, 2:] * 0.5
        else:
            raise ValueError(
                "Last dim of reference_points must be 2 or 4, but get {} instead.".format(reference_points.shape[-1]))
        output = MSDeformAttnFunction.apply(
            value, input_spatial_shapes, input_level_start_index, sampling_locations.contiguous(),
            attention_weights, self.im2col_step)
        return self._proj(self.output_proj, output)
