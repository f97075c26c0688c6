"""Projection layers of MSDeformAttn on the 5th-generation tensor cores.

`linear_tf32x3(x, weight, bias)` = `F.linear` for the module's four nn.Linear layers
(ops/modules/ms_deform_attn.py:59-62) with the forward GEMM done by the tcgen05 kernel in
csrc/linear_tf32x3.cuh (three-term TF32 split, fp32 accumulation in TMEM: fp32-grade results).
grad_x = g W runs on the same kernel (reduction over out_features); grad_W = g^T x and grad_b = sum g
(reductions over the ~10^5 rows) run on a transposing split-K tcgen05 kernel.  `USE_TCGEN05_DW = False`
sends the weight gradient back to cuBLAS through torch."""
from __future__ import annotations

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ... import load_extension

MSDA = load_extension()
USE_TCGEN05_DW = True


def matmul_split() -> int:
    """TF32 terms of the tcgen05 GEMMs that replace nn.Linear: 3 (fp32-grade, tf32x3) unless the user has allowed TF32
    matmuls (`torch.backends.cuda.matmul.allow_tf32` / `set_float32_matmul_precision("high")`), in which case nn.Linear in
    the reference would run single-pass TF32 too and one term is used.  Mask2Former leaves the flag at its default
    (off), so the default is 3."""
    return 1 if torch.backends.cuda.matmul.allow_tf32 else 3


def conv_split() -> int:
    """Same for the 1x1 `input_proj` convolutions: torch's convolutions follow `torch.backends.cudnn.allow_tf32`, which
    defaults to True, so by default the reference runs them in single-pass TF32 and so do we."""
    return 1 if torch.backends.cudnn.allow_tf32 else 3


class LinearTF32x3Function(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, split, row_mask=None):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        ctx.split = split
        y = MSDA.linear_tf32x3(x, weight, bias, split)
        if row_mask is not None:
            # y.masked_fill(row_mask[..., None], 0) on the tensor this function owns: only masked rows are written.
            # CONTRACT: the consumer must deliver a zero gradient for the masked rows (MSDeformAttnFusedFunction does:
            # it zeroes those rows of grad_value), so backward needs no masking pass.
            MSDA.zero_masked_rows_(y, row_mask)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        g2 = g.reshape(-1, g.shape[-1])
        gx = MSDA.linear_tf32x3_backward_input(g2.contiguous(), weight, ctx.split).view_as(x) \
            if ctx.needs_input_grad[0] else None
        gw = gb = None
        want_b = ctx.has_bias and ctx.needs_input_grad[2]
        if ctx.needs_input_grad[1] and USE_TCGEN05_DW:
            gw, gb_ = MSDA.linear_tf32x3_backward_weight(g2.contiguous(), x.reshape(-1, x.shape[-1]), ctx.split, want_b)
            gb = gb_ if want_b else None
        else:
            if ctx.needs_input_grad[1]:
                gw = g2.t() @ x.reshape(-1, x.shape[-1])
            if want_b:
                gb = g2.sum(0)
        return gx, gw, gb, None, None


def linear_tf32x3(x, weight, bias=None, split=None, row_mask=None):
    """Drop-in for F.linear(x, weight, bias) on CUDA float32 tensors (in_features a multiple of 256, output width a
    multiple of 256 or 288 / 192 / 96).  split=None: follow torch's matmul precision flag (`matmul_split`)."""
    return LinearTF32x3Function.apply(x, weight, bias, matmul_split() if split is None else split, row_mask)


def supported(layer: torch.nn.Linear, x: torch.Tensor) -> bool:
    return (x.is_cuda and x.dtype == torch.float32 and layer.weight.dtype == torch.float32
            and MSDA.linear_tf32x3_supported(layer.in_features, layer.out_features))


class SelfAttnProjectionsFunction(Function):
    """The three input projections of MSDeformAttn in encoder self-attention, where query = src + pos and
    input_flatten = src (reference: msdeformattn.py:123 calling ops/modules/ms_deform_attn.py:98-104):

        value   = value_proj(src)   [masked rows zeroed]
        offsets = sampling_offsets(src + pos)
        logits  = attention_weights(src + pos)

    as ONE autograd node, so the gradient branches that meet at `src + pos` and at `src` are summed in GEMM epilogues
    (`linear_tf32x3_backward_input_accumulate`) instead of by two element-wise passes of the autograd engine:
        g_q   = g_off W_o  (+)= g_logits W_a        g_src = g_value W_v + g_q        g_pos = g_q (summed over the batch
    when pos is the shared (1, S, C) table)."""

    @staticmethod
    def forward(ctx, src, pos, wv, bv, wo, bo, wa, ba, split, row_mask):
        q = src + pos
        value = MSDA.linear_tf32x3(src, wv, bv, split)
        if row_mask is not None:
            MSDA.zero_masked_rows_(value, row_mask)          # consumer zeroes the masked rows of grad_value
        offsets = MSDA.linear_tf32x3(q, wo, bo, split)
        logits = MSDA.linear_tf32x3(q, wa, ba, split)
        ctx.save_for_backward(src, q, wv, wo, wa)
        ctx.split = split
        ctx.pos_shape = pos.shape
        return value, offsets, logits

    @staticmethod
    @once_differentiable
    def backward(ctx, g_value, g_off, g_logits):
        src, q, wv, wo, wa = ctx.saved_tensors
        sp = ctx.split
        rows = src.numel() // src.shape[-1]
        gv2 = g_value.reshape(rows, -1).contiguous()
        go2 = g_off.reshape(rows, -1).contiguous()
        gl2 = g_logits.reshape(rows, -1).contiguous()
        g_src = g_pos = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            g_q = MSDA.linear_tf32x3_backward_input(go2, wo, sp)
            MSDA.linear_tf32x3_backward_input_accumulate(gl2, wa, g_q, True, sp)            # g_q += g_logits W_a
            if ctx.needs_input_grad[0]:
                g_src = MSDA.linear_tf32x3_backward_input_accumulate(gv2, wv, g_q, False, sp).view_as(src)
            if ctx.needs_input_grad[1]:
                g_pos = g_q.view_as(src)
                if tuple(ctx.pos_shape) != tuple(src.shape):
                    g_pos = g_pos.sum_to_size(ctx.pos_shape)
        x2, q2 = src.reshape(rows, -1), q.reshape(rows, -1)
        gwv, gbv = MSDA.linear_tf32x3_backward_weight(gv2, x2, sp, True)
        gwo, gbo = MSDA.linear_tf32x3_backward_weight(go2, q2, sp, True)
        gwa, gba = MSDA.linear_tf32x3_backward_weight(gl2, q2, sp, True)
        return g_src, g_pos, gwv, gbv, gwo, gbo, gwa, gba, None, None


class SelfAttnProjectionsPackedFunction(Function):
    """`SelfAttnProjectionsFunction` with `sampling_offsets` and `attention_weights` as ONE 256 -> 288 projection
    (reference: ops/modules/ms_deform_attn.py:101-102 read `query` twice): `query = src + pos` is read once, the result
    (N, S, 288) = [offsets | logits] goes to the packed fused attention op as it is, and in backward the gradient of
    `query` is one GEMM over K = 288 and the two weight gradients one grad_W launch.  The module keeps its two nn.Linear
    layers (state-dict keys); their weights are stacked here per call (295 KB)."""

    @staticmethod
    def forward(ctx, src, pos, wv, bv, wo, bo, wa, ba, split, row_mask):
        q = src + pos
        value = MSDA.linear_tf32x3(src, wv, bv, split)
        if row_mask is not None:
            MSDA.zero_masked_rows_(value, row_mask)          # consumer zeroes the masked rows of grad_value
        woa = torch.cat((wo, wa), 0)
        oa = MSDA.linear_tf32x3(q, woa, torch.cat((bo, ba), 0), split)
        ctx.save_for_backward(src, q, wv, woa)
        ctx.split = split
        ctx.pos_shape = pos.shape
        ctx.n_off = wo.shape[0]
        return value, oa

    @staticmethod
    @once_differentiable
    def backward(ctx, g_value, g_oa):
        src, q, wv, woa = ctx.saved_tensors
        sp = ctx.split
        rows = src.numel() // src.shape[-1]
        gv2 = g_value.reshape(rows, -1).contiguous()
        goa2 = g_oa.reshape(rows, -1).contiguous()
        g_src = g_pos = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            g_q = MSDA.linear_tf32x3_backward_input(goa2, woa, sp)
            if ctx.needs_input_grad[0]:
                g_src = MSDA.linear_tf32x3_backward_input_accumulate(gv2, wv, g_q, False, sp).view_as(src)
            if ctx.needs_input_grad[1]:
                g_pos = g_q.view_as(src)
                if tuple(ctx.pos_shape) != tuple(src.shape):
                    g_pos = g_pos.sum_to_size(ctx.pos_shape)
        x2, q2 = src.reshape(rows, -1), q.reshape(rows, -1)
        gwv, gbv = MSDA.linear_tf32x3_backward_weight(gv2, x2, sp, True)
        gwoa, gboa = MSDA.linear_tf32x3_backward_weight(goa2, q2, sp, True)
        n = ctx.n_off
        return g_src, g_pos, gwv, gbv, gwoa[:n], gboa[:n], gwoa[n:], gboa[n:], None, None
