"""`MSDeformAttnPixelDecoder` — the caller either side of the encoder (reference:
mask2former/modeling/pixel_decoder/msdeformattn.py:165-358 and transformer_decoder/position_encoding.py:12-52).

Same constructor arguments, sub-module names and state-dict keys as the reference (`input_proj.L.0/1.*`,
`transformer.*`, `mask_features.*`, `adapter_K.*`, `layer_K.*` with `.norm.*`), without the Detectron2 dependency:
`ShapeSpec`, the `Conv2d` wrapper (conv -> norm -> activation) and `get_norm("GN")` are restated here from
detectron2.layers (v0.6, the release Mask2Former's INSTALL.md builds against).

What runs where on CUDA float32 (`fused = True`, the default):
  input_proj (1x1 conv + GroupNorm(32)) + flatten + cat   one autograd function on the tcgen05 GEMM and the token
                                                          GroupNorm kernels (ops/functions/glue_func.py)
  PositionEmbeddingSine + level_embed                     sine table computed once per level shape by
                                                          `sine_pos_embed_kernel`, kept as (1, S, 256): the reference
                                                          builds N identical NCHW copies per forward
  encoder                                                 bm2f_b200.encoder (fused layers), token-major entry
  split back to NCHW                                      views, as in the reference (msdeformattn.py:327-339)
  FPN tail (lateral 1x1, 3x3 output conv, bilinear upsample, mask_features)   library convolutions through torch:
                                                          the step after the path (§8f rank 4), no kernel of ours
`fused = False` runs the reference's op sequence in torch around the same attention module (used by the tests as
the comparison arm on the GPU)."""
from __future__ import annotations

import collections
import math
from collections import namedtuple
from typing import Callable, Dict, List, Optional, Union

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import load_extension
from .encoder import MSDeformAttnTransformerEncoderOnly
from .ops.functions import fpn_func, glue_func

MSDA = load_extension()

ShapeSpec = namedtuple("ShapeSpec", ["channels", "height", "width", "stride"], defaults=(None, None, None, None))


def get_norm(norm, out_channels):
    """detectron2.layers.get_norm restricted to what the configs of this decoder use (SEM_SEG_HEAD.NORM: "GN")."""
    if norm is None or norm == "":
        return None
    if callable(norm):
        return norm(out_channels)
    if norm == "GN":
        return nn.GroupNorm(32, out_channels)
    raise NotImplementedError(f'norm "{norm}": this decoder is configured with "GN" (or "") in every reference config')


class Conv2d(nn.Conv2d):
    """detectron2.layers.Conv2d: convolution followed by optional `norm` and `activation` sub-modules."""

    def __init__(self, *args, norm=None, activation=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.norm = norm
        self.activation = activation

    def forward(self, x):
        x = F.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
        if self.norm is not None:
            x = self.norm(x)
        if self.activation is not None:
            x = self.activation(x)
        return x


def c2_xavier_fill(module):
    """fvcore.nn.weight_init.c2_xavier_fill: kaiming_uniform(a=1) weights, zero bias."""
    nn.init.kaiming_uniform_(module.weight, a=1)
    if module.bias is not None:
        nn.init.constant_(module.bias, 0)


class PositionEmbeddingSine(nn.Module):
    """Sine position embedding for an all-False mask (position_encoding.py:12-52).  On CUDA the table is batch
    independent: `tokens(H, W)` returns it token-major, (H*W, 2*num_pos_feats), cached per shape; `forward` returns the
    reference's (N, 2F, H, W) as an expanded view of it.  Mask2Former never passes a mask here (msdeformattn.py:322)."""

    def __init__(self, num_pos_feats=64, temperature=10000, normalize=False, scale=None):
        super().__init__()
        self.num_pos_feats = num_pos_feats
        self.temperature = temperature
        self.normalize = normalize
        if scale is not None and normalize is False:
            raise ValueError("normalize should be True if scale is passed")
        self.scale = 2 * math.pi if scale is None else scale
        self._cache = collections.OrderedDict()     # (H, W, device) -> table; small LRU, see tokens()

    # Inference sees many image sizes: keep the few most recent level shapes (a training run with fixed crops hits 3),
    # not one table per shape ever seen.  The kernel that fills a table is a single cheap pass.
    CACHE_ENTRIES = 12

    def tokens(self, height, width, like):
        key = (int(height), int(width), like.device)
        tab = self._cache.get(key)
        if tab is None:
            tab = MSDA.sine_position_embedding(like, int(height), int(width), self.num_pos_feats, float(self.temperature),
                                               float(self.scale), bool(self.normalize))
            self._cache[key] = tab
            while len(self._cache) > self.CACHE_ENTRIES:
                self._cache.popitem(last=False)
        else:
            self._cache.move_to_end(key)
        return tab

    def __deepcopy__(self, memo):
        # cached device tables are derived state: a copy (or a .to() of it) starts empty
        new = type(self)(self.num_pos_feats, self.temperature, self.normalize, self.scale if self.normalize else None)
        new.scale = self.scale
        memo[id(self)] = new
        return new

    def forward(self, x, mask=None):
        if mask is not None:
            raise NotImplementedError("PositionEmbeddingSine: padding masks are not used by this pixel decoder")
        n, _, h, w = x.shape
        tab = self.tokens(h, w, x)
        return tab.t().reshape(1, 2 * self.num_pos_feats, h, w).expand(n, -1, -1, -1)

    def __repr__(self, _repr_indent=4):
        head = "Positional encoding " + self.__class__.__name__
        body = [f"num_pos_feats: {self.num_pos_feats}", f"temperature: {self.temperature}",
                f"normalize: {self.normalize}", f"scale: {self.scale}"]
        return "\n".join([head] + [" " * _repr_indent + line for line in body])


class MSDeformAttnPixelDecoder(nn.Module):
    def __init__(self, input_shape: Dict[str, ShapeSpec], *, transformer_dropout: float, transformer_nheads: int,
                 transformer_dim_feedforward: int, transformer_enc_layers: int, conv_dim: int, mask_dim: int,
                 norm: Optional[Union[str, Callable]] = None, transformer_in_features: List[str], common_stride: int):
        super().__init__()
        transformer_input_shape = {k: v for k, v in input_shape.items() if k in transformer_in_features}
        input_shape = sorted(input_shape.items(), key=lambda x: x[1].stride)
        self.in_features = [k for k, v in input_shape]                  # "res2" .. "res5"
        self.feature_strides = [v.stride for k, v in input_shape]
        self.feature_channels = [v.channels for k, v in input_shape]
        transformer_input_shape = sorted(transformer_input_shape.items(), key=lambda x: x[1].stride)
        self.transformer_in_features = [k for k, v in transformer_input_shape]
        transformer_in_channels = [v.channels for k, v in transformer_input_shape]
        self.transformer_feature_strides = [v.stride for k, v in transformer_input_shape]
        self.transformer_num_feature_levels = len(self.transformer_in_features)

        # lowest resolution first (res5 -> res3), msdeformattn.py:211-227
        chans = transformer_in_channels[::-1] if self.transformer_num_feature_levels > 1 else [transformer_in_channels[-1]]
        self.input_proj = nn.ModuleList([
            nn.Sequential(nn.Conv2d(c, conv_dim, kernel_size=1), nn.GroupNorm(32, conv_dim)) for c in chans])
        for proj in self.input_proj:
            nn.init.xavier_uniform_(proj[0].weight, gain=1)
            nn.init.constant_(proj[0].bias, 0)

        self.transformer = MSDeformAttnTransformerEncoderOnly(
            d_model=conv_dim, dropout=transformer_dropout, nhead=transformer_nheads,
            dim_feedforward=transformer_dim_feedforward, num_encoder_layers=transformer_enc_layers,
            num_feature_levels=self.transformer_num_feature_levels)
        self.pe_layer = PositionEmbeddingSine(conv_dim // 2, normalize=True)

        self.mask_dim = mask_dim
        self.mask_features = Conv2d(conv_dim, mask_dim, kernel_size=1, stride=1, padding=0)
        c2_xavier_fill(self.mask_features)
        self.maskformer_num_feature_levels = 3      # always use 3 scales
        self.common_stride = common_stride

        stride = min(self.transformer_feature_strides)
        self.num_fpn_levels = int(np.log2(stride) - np.log2(self.common_stride))
        lateral_convs, output_convs = [], []
        use_bias = norm == ""
        for idx, in_channels in enumerate(self.feature_channels[:self.num_fpn_levels]):
            lateral_conv = Conv2d(in_channels, conv_dim, kernel_size=1, bias=use_bias, norm=get_norm(norm, conv_dim))
            output_conv = Conv2d(conv_dim, conv_dim, kernel_size=3, stride=1, padding=1, bias=use_bias,
                                 norm=get_norm(norm, conv_dim), activation=F.relu)
            c2_xavier_fill(lateral_conv)
            c2_xavier_fill(output_conv)
            self.add_module("adapter_{}".format(idx + 1), lateral_conv)
            self.add_module("layer_{}".format(idx + 1), output_conv)
            lateral_convs.append(lateral_conv)
            output_convs.append(output_conv)
        # top-down order (low to high resolution)
        self.lateral_convs = lateral_convs[::-1]
        self.output_convs = output_convs[::-1]
        self.fused = True

    @classmethod
    def from_config(cls, cfg, input_shape: Dict[str, ShapeSpec]):
        """Constructor arguments from the reference's config keys (msdeformattn.py:294-312)."""
        head, mf = cfg.MODEL.SEM_SEG_HEAD, cfg.MODEL.MASK_FORMER
        return {
            "input_shape": {k: v for k, v in input_shape.items() if k in head.IN_FEATURES},
            "conv_dim": head.CONVS_DIM,
            "mask_dim": head.MASK_DIM,
            "norm": head.NORM,
            "transformer_dropout": mf.DROPOUT,
            "transformer_nheads": mf.NHEADS,
            "transformer_dim_feedforward": 1024,        # the reference hard-codes 1024 for this encoder
            "transformer_enc_layers": head.TRANSFORMER_ENC_LAYERS,
            "transformer_in_features": head.DEFORMABLE_TRANSFORMER_ENCODER_IN_FEATURES,
            "common_stride": head.COMMON_STRIDE,
        }

    @classmethod
    def build(cls, cfg, input_shape):
        return cls(**cls.from_config(cfg, input_shape))

    # ------------------------------------------------------------------------------------------------------
    def _encode_fused(self, xs):
        shapes_list = [(x.shape[2], x.shape[3]) for x in xs]
        src_flatten = glue_func.input_proj_flatten(xs, self.input_proj)
        level_embed = self.transformer.level_embed
        lvl_pos = torch.cat([self.pe_layer.tokens(h, w, xs[0]) + level_embed[lvl].view(1, -1)
                             for lvl, (h, w) in enumerate(shapes_list)], 0)[None]
        return self.transformer.forward_tokens(src_flatten, lvl_pos, shapes_list)

    def _encode_reference_sequence(self, xs):
        srcs = [self.input_proj[idx](x) for idx, x in enumerate(xs)]
        pos = [self.pe_layer(x) for x in xs]
        return self.transformer(srcs, pos)

    def forward_features(self, features):
        with torch.autocast(device_type="cuda", enabled=False):
            # lowest resolution first (msdeformattn.py:318-322); the op has no half-precision path in this decoder
            xs = [features[f].float() for f in self.transformer_in_features[::-1]]
            if not xs[0].is_cuda:
                raise RuntimeError("MSDeformAttnPixelDecoder: CUDA tensors only (no CPU path)")
            if self.fused and glue_func.supported(xs, self.input_proj):
                y, spatial_shapes, level_start_index = self._encode_fused(xs)
                shapes_list = [(x.shape[2], x.shape[3]) for x in xs]
            else:
                y, spatial_shapes, level_start_index = self._encode_reference_sequence(xs)
                shapes_list = [(x.shape[2], x.shape[3]) for x in xs]
            bs = y.shape[0]
            sizes = [h * w for h, w in shapes_list]
            out = [z.transpose(1, 2).view(bs, -1, h, w) for z, (h, w) in zip(torch.split(y, sizes, dim=1), shapes_list)]

            # extra FPN levels, top-down (msdeformattn.py:341-351)
            fpn_xs = [features[f].float() for f in self.in_features[:self.num_fpn_levels][::-1]]
            if self.fused and fpn_xs and fpn_func.supported(fpn_xs, self.lateral_convs, self.output_convs, self.mask_features):
                # token rows throughout (ops/functions/fpn_func.py): the coarser level is read where it lies in the
                # encoder output, every FPN level stays (N, H, W, 256) until mask_features writes NCHW
                prev, (ph, pw) = y[:, sum(sizes[:-1]):], shapes_list[-1]
                for idx, x in enumerate(fpn_xs):
                    tok = fpn_func.fpn_level(x, prev, ph, pw, self.lateral_convs[idx], self.output_convs[idx])
                    out.append(tok.permute(0, 3, 1, 2))          # NCHW-shaped view (channels_last memory)
                    prev, (ph, pw) = tok.view(bs, -1, tok.shape[-1]), tok.shape[1:3]
                multi_scale_features = out[:self.maskformer_num_feature_levels]
                return fpn_func.mask_features_tokens(tok, self.mask_features), out[0], multi_scale_features
            for idx, x in enumerate(fpn_xs):
                cur_fpn = self.lateral_convs[idx](x)
                y = cur_fpn + F.interpolate(out[-1], size=cur_fpn.shape[-2:], mode="bilinear", align_corners=False)
                out.append(self.output_convs[idx](y))

            multi_scale_features = out[:self.maskformer_num_feature_levels]
            return self.mask_features(out[-1]), out[0], multi_scale_features
