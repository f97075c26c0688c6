"""FPN tail of the pixel decoder as autograd functions on token rows (reference: msdeformattn.py:341-358):

    cur_fpn = lateral_conv(x)              Conv2d(C_in, 256, 1, bias=False) + GroupNorm(32, 256)
    y = cur_fpn + F.interpolate(out[-1], size=cur_fpn.shape[-2:], mode="bilinear", align_corners=False)
    y = output_conv(y)                     Conv2d(256, 256, 3, padding=1, bias=False) + GroupNorm(32, 256) + ReLU
    ...
    mask_features(out[-1])                 Conv2d(256, mask_dim, 1)

`fpn_level` is the first three lines as ONE node: the 1x1 conv is a tcgen05 GEMM over token rows, GroupNorm statistics
are one read of its output, normalisation + upsample + add write the zero-haloed image the 3x3 convolution reads, the
3x3 convolution is an implicit GEMM on tcgen05 whose filter taps are row shifts of that image (csrc/fpn_kernels.cuh,
linear_tf32x3_persistent_kernel<..., CONV>), GroupNorm + ReLU is one pass.  No NCHW intermediate, no cuDNN / ATen
convolution, interpolation or normalisation kernel.  `mask_features_tokens` is the last line (GEMM + one transpose to
the NCHW tensor the transformer decoder consumes).  Convolution precision follows `torch.backends.cudnn.allow_tf32`
like the nn.Conv2d layers replaced (`linear_func.conv_split`).  CUDA float32 only; there is no CPU path."""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ... import load_extension
from .glue_func import _tokens_view

MSDA = load_extension()


class FpnLevelFunction(Function):
    """x (N, C_in, H, W), prev (N, h * w, 256) token rows of the coarser level (batch-strided view allowed)
    -> (N, H, W, 256) token rows of relu(GN(conv3x3(GN(conv1x1(x)) + upsample(prev))))."""

    @staticmethod
    def forward(ctx, x, prev, prev_h, prev_w, lat_w, lat_gamma, lat_beta, out_w, out_gamma, out_beta, eps, split):
        n, c_in, h, w = x.shape
        w2d = lat_w.reshape(lat_w.shape[0], c_in)
        xt, nhwc = _tokens_view(x)
        tc = bool(MSDA.linear_tf32x3_supported(c_in, 256))
        lraw = (MSDA.linear_tf32x3(xt, w2d, None, split) if tc else F.linear(xt, w2d)).view(n, h, w, 256)
        mean1, rstd1 = MSDA.groupnorm_tokens_stats(lraw, eps)
        y_halo = MSDA.fpn_merge_forward(lraw, mean1, rstd1, lat_gamma, lat_beta, prev, prev_h, prev_w)
        craw = MSDA.conv3x3_tokens_forward(y_halo, out_w, split)
        mean2, rstd2 = MSDA.groupnorm_tokens_stats(craw, eps)
        out = MSDA.groupnorm_relu_tokens_apply(craw, mean2, rstd2, out_gamma, out_beta)
        ctx.save_for_backward(xt, lraw, mean1, rstd1, lat_gamma, y_halo, craw, mean2, rstd2, out_gamma, out_beta, w2d, out_w)
        ctx.meta = (n, c_in, h, w, prev_h, prev_w, nhwc, tc, lat_w.shape, split)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        xt, lraw, mean1, rstd1, g1, y_halo, craw, mean2, rstd2, g2, b2, w2d, out_w = ctx.saved_tensors
        n, c_in, h, w, ph, pw, nhwc, tc, lat_shape, split = ctx.meta
        need = ctx.needs_input_grad
        gc_halo, dg2, db2 = MSDA.groupnorm_relu_tokens_backward(g.contiguous(), craw, mean2, rstd2, g2, b2)
        gw_out = MSDA.conv3x3_tokens_backward_weight(gc_halo, y_halo, split) if need[7] else None
        gy = MSDA.conv3x3_tokens_backward_input(gc_halo, out_w, split)                  # (N, H, W, 256) dense
        g_prev = MSDA.fpn_upsample_backward(gy, ph, pw) if need[1] else None
        gl, dg1, db1 = MSDA.groupnorm_tokens_backward(gy.view(n, h * w, 256), 0, lraw.view(n, h * w, 256), mean1, rstd1, g1)
        gw_lat = gx = None
        if need[4]:
            if tc:
                gw_lat = MSDA.linear_tf32x3_backward_weight(gl, xt, split, False)[0]
            else:
                gw_lat = gl.reshape(-1, 256).t() @ xt.reshape(-1, c_in)
            gw_lat = gw_lat.view(lat_shape)
        if need[0]:
            gxt = MSDA.linear_tf32x3_backward_input(gl, w2d, split) if tc else gl @ w2d
            gx = gxt.view(n, h, w, c_in).permute(0, 3, 1, 2) if nhwc else MSDA.transpose_batched(gxt).view(n, c_in, h, w)
        return gx, g_prev, None, None, gw_lat, dg1, db1, gw_out, dg2, db2, None, None


class MaskFeaturesFunction(Function):
    """tokens (N, H, W, 256) -> NCHW (N, mask_dim, H, W) = Conv2d(256, mask_dim, 1)(tokens as an image).  Widths the tcgen05
    kernels are not instantiated for (mask_dim not 256 / 192 / 96) use cuBLAS through torch for the GEMM only."""

    @staticmethod
    def forward(ctx, tokens, weight, bias, split):
        n, h, w, c = tokens.shape
        w2d = weight.reshape(weight.shape[0], c)
        tc = bool(MSDA.linear_tf32x3_supported(c, w2d.shape[0]))
        t3 = tokens.view(n, h * w, c)
        y = MSDA.linear_tf32x3(t3, w2d, bias, split) if tc else F.linear(t3, w2d, bias)      # (N, HW, mask_dim)
        ctx.save_for_backward(tokens, w2d)
        ctx.meta = (weight.shape, bias is not None, split, tc)
        return MSDA.transpose_batched(y).view(n, w2d.shape[0], h, w)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        tokens, w2d = ctx.saved_tensors
        w_shape, has_bias, split, tc = ctx.meta
        n, h, w, c = tokens.shape
        gt, _ = _tokens_view(g)                                                         # (N, HW, mask_dim)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = (MSDA.linear_tf32x3_backward_input(gt, w2d, split) if tc else gt @ w2d).view(n, h, w, c)
        if ctx.needs_input_grad[1] or (has_bias and ctx.needs_input_grad[2]):
            if tc:
                gw, gb = MSDA.linear_tf32x3_backward_weight(gt, tokens.view(n, h * w, c), split, has_bias)
            else:
                g2 = gt.reshape(-1, gt.shape[-1])
                gw, gb = g2.t() @ tokens.view(-1, c), (g2.sum(0) if has_bias else None)
            gw = gw.view(w_shape)
        return gx, gw, (gb if has_bias else None), None


def _split(split):
    if split is None:
        from .linear_func import conv_split
        return conv_split()
    return split


def fpn_level(x, prev_tokens, prev_h, prev_w, lateral_conv, output_conv, split=None):
    """One top-down FPN step.  lateral_conv / output_conv: detectron2-style Conv2d wrappers with `.norm` = GroupNorm(32, 256)
    (msdeformattn.py:261-283).  prev_tokens: (N, prev_h * prev_w, 256) rows (a view into the encoder output is fine).
    Returns (N, H, W, 256) token rows."""
    return FpnLevelFunction.apply(x, prev_tokens, prev_h, prev_w, lateral_conv.weight, lateral_conv.norm.weight,
                                  lateral_conv.norm.bias, output_conv.weight, output_conv.norm.weight, output_conv.norm.bias,
                                  lateral_conv.norm.eps, _split(split))


def mask_features_tokens(tokens, conv, split=None):
    """`mask_features` (msdeformattn.py:243-250, 356) on token rows: returns the NCHW tensor the reference returns."""
    return MaskFeaturesFunction.apply(tokens, conv.weight, conv.bias, _split(split))


def supported(features, lateral_convs, output_convs, mask_conv) -> bool:
    """The fused tail covers the reference's standard configuration: 256 channels, GroupNorm(32) without conv biases,
    ReLU after the output conv, float32 CUDA features."""
    if mask_conv.in_channels != 256 or mask_conv.kernel_size != (1, 1) or getattr(mask_conv, "norm", None) is not None \
            or getattr(mask_conv, "activation", None) is not None:
        return False
    for x, lat, outc in zip(features, lateral_convs, output_convs):
        for conv, k in ((lat, 1), (outc, 3)):
            gn = getattr(conv, "norm", None)
            if not (isinstance(gn, torch.nn.GroupNorm) and gn.num_groups == 32 and gn.num_channels == 256 and gn.affine
                    and conv.bias is None and conv.out_channels == 256 and conv.kernel_size == (k, k)
                    and conv.stride == (1, 1) and conv.groups == 1 and conv.dilation == (1, 1)):
                return False
        if outc.in_channels != 256 or outc.padding != (1, 1) or getattr(outc, "activation", None) is not F.relu:
            return False
        if getattr(lat, "activation", None) is not None or lat.norm.eps != outc.norm.eps:
            return False
        if not (x.is_cuda and x.dtype == torch.float32):
            return False
    return True
