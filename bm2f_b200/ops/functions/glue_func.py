"""Encoder-input glue of the pixel decoder as one autograd function (reference: msdeformattn.py:214-227 `input_proj`,
316-322 `forward_features`, 66-80 flatten / transpose / cat in `MSDeformAttnTransformerEncoderOnly.forward`).

    src_flatten[:, start_l : start_l + H_l W_l, :] = GroupNorm32(conv1x1_l(x_l)) as token rows,   l = 0 .. L-1

Per level: NCHW feature -> token rows (one tiled transpose; none if the backbone tensor is channels_last), the 1x1 conv
as a tcgen05 tf32x3 GEMM over tokens (csrc/linear_tf32x3.cuh), GroupNorm on token rows written straight into the
level's slice of the (N, S, 256) encoder input (csrc/glue_kernels.cuh).  No NCHW intermediate, no flatten/transpose
copy, no torch.cat.  Backward: token GroupNorm backward, tcgen05 grad_W / grad_x GEMMs, transpose back to NCHW.
Layer widths the tcgen05 kernels are not instantiated for (in_channels not a multiple of 256, e.g. Swin's 192 / 384)
use cuBLAS through torch for that GEMM only.  CUDA float32 only; there is no CPU path."""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ... import load_extension

MSDA = load_extension()


def _tokens_view(x):
    """(N, C, H, W) -> (N, H*W, C) rows; returns (tokens, was_channels_last)."""
    n, c, h, w = x.shape
    nhwc = x.permute(0, 2, 3, 1)
    if nhwc.is_contiguous():
        return nhwc.reshape(n, h * w, c), True
    return MSDA.transpose_batched(x.reshape(n, c, h * w)), False


class InputProjFlattenFunction(Function):
    @staticmethod
    def forward(ctx, n_levels, eps, split, *args):
        xs, params = args[:n_levels], args[n_levels:]
        n = xs[0].shape[0]
        shapes = [(x.shape[2], x.shape[3]) for x in xs]
        total = sum(h * w for h, w in shapes)
        out = torch.empty((n, total, 256), device=xs[0].device, dtype=torch.float32)
        saved, meta, start = [], [], 0
        for lvl, x in enumerate(xs):
            conv_w, conv_b, gn_w, gn_b = params[4 * lvl: 4 * lvl + 4]
            c_in = x.shape[1]
            w2d = conv_w.reshape(conv_w.shape[0], c_in)
            xt, nhwc = _tokens_view(x)
            tc = bool(MSDA.linear_tf32x3_supported(c_in, w2d.shape[0]))
            y = MSDA.linear_tf32x3(xt, w2d, conv_b, split) if tc else F.linear(xt, w2d, conv_b)
            mean, rstd = MSDA.groupnorm_tokens_forward(y, gn_w, gn_b, eps, out, start)
            saved += [xt, y, mean, rstd, w2d, gn_w]
            meta.append((start, shapes[lvl], c_in, nhwc, tc, conv_w.shape))
            start += shapes[lvl][0] * shapes[lvl][1]
        ctx.save_for_backward(*saved)
        ctx.meta, ctx.n_levels, ctx.split = meta, n_levels, split
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        g = g.contiguous()
        L = ctx.n_levels
        gxs, gparams = [None] * L, []
        for lvl in range(L):
            xt, y, mean, rstd, w2d, gn_w = ctx.saved_tensors[6 * lvl: 6 * lvl + 6]
            start, (h, w), c_in, nhwc, tc, w_shape = ctx.meta[lvl]
            dy, dgamma, dbeta = MSDA.groupnorm_tokens_backward(g, start, y, mean, rstd, gn_w)
            want_w = ctx.needs_input_grad[3 + L + 4 * lvl]
            want_b = ctx.needs_input_grad[3 + L + 4 * lvl + 1]
            gw = gb = None
            if want_w or want_b:
                if tc:
                    gw, gb = MSDA.linear_tf32x3_backward_weight(dy, xt, ctx.split, True)
                else:
                    dy2 = dy.reshape(-1, dy.shape[-1])
                    gw, gb = dy2.t() @ xt.reshape(-1, c_in), dy2.sum(0)
                gw = gw.view(w_shape)
            if ctx.needs_input_grad[3 + lvl]:
                dxt = MSDA.linear_tf32x3_backward_input(dy, w2d, ctx.split) if tc else dy @ w2d
                n = dxt.shape[0]
                if nhwc:
                    gxs[lvl] = dxt.view(n, h, w, c_in).permute(0, 3, 1, 2)
                else:
                    gxs[lvl] = MSDA.transpose_batched(dxt).view(n, c_in, h, w)
            gparams += [gw if want_w else None, gb if want_b else None, dgamma, dbeta]
        return (None, None, None, *gxs, *gparams)


def input_proj_flatten(xs, input_proj, split=None):
    """xs: list of (N, C_l, H_l, W_l) CUDA float32 features, lowest resolution first (msdeformattn.py:319-321);
    input_proj: ModuleList of Sequential(Conv2d(C_l, 256, 1), GroupNorm(32, 256)).  Returns (N, S, 256)."""
    params = []
    for seq in input_proj:
        conv, gn = seq[0], seq[1]
        params += [conv.weight, conv.bias, gn.weight, gn.bias]
    eps = input_proj[0][1].eps
    if split is None:
        from .linear_func import conv_split
        split = conv_split()          # follows torch.backends.cudnn.allow_tf32 like the reference's nn.Conv2d
    return InputProjFlattenFunction.apply(len(xs), eps, split, *xs, *params)


def supported(xs, input_proj) -> bool:
    if len(xs) != len(input_proj):
        return False
    for x, seq in zip(xs, input_proj):
        conv, gn = seq[0], seq[1]
        if not (x.is_cuda and x.dtype == torch.float32 and conv.out_channels == 256 and conv.kernel_size == (1, 1)
                and conv.bias is not None and gn.num_groups == 32 and gn.affine and gn.eps == input_proj[0][1].eps):
            return False
    return True
