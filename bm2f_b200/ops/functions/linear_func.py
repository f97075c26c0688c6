"""Projection layers of MSDeformAttn on the 5th-generation tensor cores.

`linear_tf32x3(x, weight, bias)` = `F.linear` for the module's four nn.Linear layers
(ops/modules/ms_deform_attn.py:59-62) with the forward GEMM done by the tcgen05 kernel in
csrc/linear_tf32x3.cuh (three-term TF32 split, fp32 accumulation in TMEM: fp32-grade results).
grad_x = g W runs on the same kernel (reduction over out_features); grad_W = g^T x and grad_b = sum g
reduce over the ~10^5 rows and are left to cuBLAS through torch in this round."""
from __future__ import annotations

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ... import load_extension

MSDA = load_extension()


class LinearTF32x3Function(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, split):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        ctx.split = split
        return MSDA.linear_tf32x3(x, weight, bias, split)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        g2 = g.reshape(-1, g.shape[-1])
        gx = MSDA.linear_tf32x3_backward_input(g2.contiguous(), weight, ctx.split).view_as(x) \
            if ctx.needs_input_grad[0] else None
        gw = g2.t() @ x.reshape(-1, x.shape[-1]) if ctx.needs_input_grad[1] else None
        gb = g2.sum(0) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return gx, gw, gb, None


def linear_tf32x3(x, weight, bias=None, split=3):
    """Drop-in for F.linear(x, weight, bias) on CUDA float32 tensors with in_features = 256."""
    return LinearTF32x3Function.apply(x, weight, bias, split)


def supported(layer: torch.nn.Linear, x: torch.Tensor) -> bool:
    return (x.is_cuda and x.dtype == torch.float32 and layer.weight.dtype == torch.float32
            and MSDA.linear_tf32x3_supported(layer.in_features, layer.out_features))
