"""Autograd functions of the encoder layer's non-attention half (reference: msdeformattn.py:92-131):
FFN first layer with the ReLU fused into the tcgen05 GEMM epilogue, and residual-add + LayerNorm in one
pass.  Forward and backward both run on the hand-written sm_100a kernels (csrc/linear_tf32x3.cuh,
csrc/ln_kernels.cuh); no CPU path."""
from __future__ import annotations

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ... import load_extension

MSDA = load_extension()


class FFNFunction(Function):
    """y = linear2(relu(linear1(x))) (reference: msdeformattn.py:116-117, dropout 0).  Forward: two tcgen05 GEMMs, the
    ReLU in the first one's epilogue.  Backward: grad_h = (g W2) masked by h > 0 in that GEMM's epilogue, then
    grad_x = grad_h W1 and the two weight / bias gradients — five tcgen05 GEMMs, no element-wise pass."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, split):
        h = MSDA.linear_relu_tf32x3(x, w1, b1, split)
        y = MSDA.linear_tf32x3(h, w2, b2, split)
        ctx.save_for_backward(x, h, w1, w2)
        ctx.split = split
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, h, w1, w2 = ctx.saved_tensors
        gx, gw1, gb1, gw2, gb2 = MSDA.ffn_tf32x3_backward(g.contiguous(), x, h, w1, w2, ctx.split)
        return gx, gw1, gb1, gw2, gb2, None


class AddLayerNormFunction(Function):
    """y = LayerNorm(x + residual) * gamma + beta (256 channels); both inputs receive the same gradient."""

    @staticmethod
    def forward(ctx, x, residual, gamma, beta, eps):
        y, z, mean, rstd = MSDA.add_layernorm_forward(x, residual, gamma, beta, eps)
        ctx.save_for_backward(z, mean, rstd, gamma)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        z, mean, rstd, gamma = ctx.saved_tensors
        dz, dgamma, dbeta = MSDA.add_layernorm_backward(g.contiguous(), z, mean, rstd, gamma)
        return dz, dz, dgamma, dbeta, None


def ffn(x, linear1: torch.nn.Linear, linear2: torch.nn.Linear, split=None):
    from .linear_func import matmul_split
    return FFNFunction.apply(x, linear1.weight, linear1.bias, linear2.weight, linear2.bias,
                             matmul_split() if split is None else split)


def add_layernorm(x, residual, norm: torch.nn.LayerNorm):
    return AddLayerNormFunction.apply(x, residual, norm.weight, norm.bias, norm.eps)
