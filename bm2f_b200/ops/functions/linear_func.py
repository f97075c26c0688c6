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


def linear_tf32x3(x, weight, bias=None, split=3, row_mask=None):
    """Drop-in for F.linear(x, weight, bias) on CUDA float32 tensors (in_features a multiple of 256, output width a
    multiple of 256 or 288 / 192 / 96)."""
    return LinearTF32x3Function.apply(x, weight, bias, split, row_mask)


def supported(layer: torch.nn.Linear, x: torch.Tensor) -> bool:
    return (x.is_cuda and x.dtype == torch.float32 and layer.weight.dtype == torch.float32
            and MSDA.linear_tf32x3_supported(layer.in_features, layer.out_features))
