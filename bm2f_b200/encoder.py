"""Deformable-attention transformer encoder of the pixel decoder — the caller of the hot path
(reference: mask2former/modeling/pixel_decoder/msdeformattn.py:22-161).

Same class names, constructor arguments, sub-module names (hence state-dict keys: `self_attn.*`, `norm1`,
`linear1`, `linear2`, `norm2`, `level_embed`, `encoder.layers.N.*`) and forward signatures as the reference, so
checkpoints load.  Detectron2 is not needed for this part; the decoder around it (input projections, position
embedding, FPN tail) is mirrored in `bm2f_b200.pixel_decoder`, which enters through `forward_tokens`.

On CUDA float32 with dropout 0 (every config that selects this decoder sets DROPOUT 0.0) the layer runs:
    q = src + pos                         torch add
    src2 = MSDeformAttn(q, ref, src)      fused-prologue sampling kernels + tcgen05 projections (`fuse_projections`:
                                          the three input projections as one autograd node, off by default)
    src = LayerNorm(src + src2)           one fused kernel (csrc/ln_kernels.cuh)
    src2 = linear2(relu(linear1(src)))    two tcgen05 GEMMs (ReLU in the epilogue); backward = five tcgen05 GEMMs,
                                          the ReLU mask applied in the epilogue of the grad_h GEMM
    src = LayerNorm(src + src2)           one fused kernel
Other settings (dropout > 0 in training, other activations) use the reference's op sequence in torch around
the same attention module.
"""
from __future__ import annotations

import copy

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn.init import normal_

from .ops.functions import encoder_func, linear_func
from .ops.modules import MSDeformAttn


def _get_activation_fn(activation):
    if activation == "relu":
        return F.relu
    if activation == "gelu":
        return F.gelu
    if activation == "glu":
        return F.glu
    raise RuntimeError(f"activation should be relu/gelu, not {activation}.")


def _get_clones(module, N):
    return nn.ModuleList([copy.deepcopy(module) for _ in range(N)])


class MSDeformAttnTransformerEncoderLayer(nn.Module):
    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        self.self_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.activation = _get_activation_fn(activation)
        self._activation_name = activation
        self.dropout2 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout3 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)
        self.fused = True          # False: always the reference op sequence in torch
        # value / offset / logit projections of self-attention as ONE autograd node (MSDeformAttn.forward_self_attention):
        # with the packed 256 -> 288 offsets||logits projection the node is the fastest variant (alternating A/B at cfg 2 x
        # 16, fwd+bwd: 87.9 ms per encoder pass vs 88.4 with separate projection nodes and 89.3 with one node and two
        # projections, profiles/r02_encoder_bench.txt).  With TF32 matmuls allowed the separate nodes measured faster in
        # round 1 (84.1 vs 85.4 ms), so the node is used for fp32-grade (default) precision only.
        self.fuse_projections = True

    @staticmethod
    def with_pos_embed(tensor, pos):
        return tensor if pos is None else tensor + pos

    def _fast_ok(self, src):
        no_dropout = (not self.training) or (self.dropout1.p == 0 and self.dropout2.p == 0 and self.dropout3.p == 0)
        return (self.fused and no_dropout and self._activation_name == "relu" and src.is_cuda
                and src.dtype == torch.float32 and src.shape[-1] == 256
                and linear_func.supported(self.linear1, src) and self.linear2.in_features % 256 == 0)

    def forward_ffn(self, src):
        if self._fast_ok(src):
            src2 = encoder_func.ffn(src, self.linear1, self.linear2)
            return encoder_func.add_layernorm(src, src2, self.norm2)
        src2 = self.linear2(self.dropout2(self.activation(self.linear1(src))))
        src = src + self.dropout3(src2)
        return self.norm2(src)

    def forward(self, src, pos, reference_points, spatial_shapes, level_start_index, padding_mask=None):
        if (self.fuse_projections and not torch.backends.cuda.matmul.allow_tf32 and self._fast_ok(src)
                and self.self_attn.self_attention_supported(src, pos, reference_points)):
            # query = src + pos, input = src: the three input projections run as one autograd node
            src2 = self.self_attn.forward_self_attention(src, pos, reference_points, spatial_shapes, level_start_index,
                                                         padding_mask)
        else:
            src2 = self.self_attn(self.with_pos_embed(src, pos), reference_points, src, spatial_shapes,
                                  level_start_index, padding_mask)
        if self._fast_ok(src):
            src = encoder_func.add_layernorm(src, src2, self.norm1)
        else:
            src = self.norm1(src + self.dropout1(src2))
        return self.forward_ffn(src)


class MSDeformAttnTransformerEncoder(nn.Module):
    def __init__(self, encoder_layer, num_layers):
        super().__init__()
        self.layers = _get_clones(encoder_layer, num_layers)
        self.num_layers = num_layers

    @staticmethod
    def get_reference_points(spatial_shapes, valid_ratios, device):
        """Pixel-centre reference points per level, (N, S, L, 2) in (x, y).  `spatial_shapes` is iterated on the
        host like in the reference (msdeformattn.py:141-153); pass a list of (H, W) to avoid a device sync."""
        pts = []
        shapes = spatial_shapes.tolist() if torch.is_tensor(spatial_shapes) else list(spatial_shapes)
        for lvl, (H_, W_) in enumerate(shapes):
            ref_y, ref_x = torch.meshgrid(torch.linspace(0.5, H_ - 0.5, H_, dtype=torch.float32, device=device),
                                          torch.linspace(0.5, W_ - 0.5, W_, dtype=torch.float32, device=device),
                                          indexing="ij")
            ref_y = ref_y.reshape(-1)[None] / (valid_ratios[:, None, lvl, 1] * H_)
            ref_x = ref_x.reshape(-1)[None] / (valid_ratios[:, None, lvl, 0] * W_)
            pts.append(torch.stack((ref_x, ref_y), -1))
        reference_points = torch.cat(pts, 1)
        return reference_points[:, :, None] * valid_ratios[:, None]

    def forward(self, src, spatial_shapes, level_start_index, valid_ratios, pos=None, padding_mask=None,
                spatial_shapes_list=None, unit_valid_ratios=False):
        """`unit_valid_ratios=True` (set by `MSDeformAttnTransformerEncoderOnly`, whose masks are all-False by
        construction, msdeformattn.py:62) marks the reference points as plain pixel centres so the fused sampling
        kernels can compute them from the query index instead of loading them."""
        output = src
        reference_points = self.get_reference_points(
            spatial_shapes_list if spatial_shapes_list is not None else spatial_shapes, valid_ratios, src.device)
        if unit_valid_ratios:
            reference_points.pixel_centres = True
        if spatial_shapes_list is not None and torch.is_tensor(spatial_shapes):
            # lets MSDeformAttn.forward check the table against Len_in on the host (no device sync per layer)
            spatial_shapes.hw_list = [(int(h), int(w)) for h, w in spatial_shapes_list]
        for layer in self.layers:
            output = layer(output, pos, reference_points, spatial_shapes, level_start_index, padding_mask)
        return output


class MSDeformAttnTransformerEncoderOnly(nn.Module):
    def __init__(self, d_model=256, nhead=8, num_encoder_layers=6, dim_feedforward=1024, dropout=0.1,
                 activation="relu", num_feature_levels=4, enc_n_points=4):
        super().__init__()
        self.d_model = d_model
        self.nhead = nhead
        encoder_layer = MSDeformAttnTransformerEncoderLayer(d_model, dim_feedforward, dropout, activation,
                                                            num_feature_levels, nhead, enc_n_points)
        self.encoder = MSDeformAttnTransformerEncoder(encoder_layer, num_encoder_layers)
        self.level_embed = nn.Parameter(torch.Tensor(num_feature_levels, d_model))
        self._reset_parameters()

    def _reset_parameters(self):
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        for m in self.modules():
            if isinstance(m, MSDeformAttn):
                m._reset_parameters()
        normal_(self.level_embed)

    @staticmethod
    def get_valid_ratio(mask):
        _, H, W = mask.shape
        valid_H = torch.sum(~mask[:, :, 0], 1)
        valid_W = torch.sum(~mask[:, 0, :], 1)
        return torch.stack([valid_W.float() / W, valid_H.float() / H], -1)

    def forward(self, srcs, pos_embeds):
        """srcs / pos_embeds: lists of (N, C, H_l, W_l), smallest level first (res5 -> res3).
        Returns (memory (N, S, C), spatial_shapes (L, 2) int64, level_start_index (L) int64)."""
        masks = [torch.zeros((x.size(0), x.size(2), x.size(3)), device=x.device, dtype=torch.bool) for x in srcs]
        src_flatten, mask_flatten, lvl_pos_embed_flatten, shapes_list = [], [], [], []
        for lvl, (src, mask, pos_embed) in enumerate(zip(srcs, masks, pos_embeds)):
            bs, c, h, w = src.shape
            shapes_list.append((h, w))
            src_flatten.append(src.flatten(2).transpose(1, 2))
            mask_flatten.append(mask.flatten(1))
            lvl_pos_embed_flatten.append(pos_embed.flatten(2).transpose(1, 2) + self.level_embed[lvl].view(1, 1, -1))
        src_flatten = torch.cat(src_flatten, 1)
        mask_flatten = torch.cat(mask_flatten, 1)
        lvl_pos_embed_flatten = torch.cat(lvl_pos_embed_flatten, 1)
        spatial_shapes = torch.as_tensor(shapes_list, dtype=torch.long, device=src_flatten.device)
        level_start_index = torch.cat((spatial_shapes.new_zeros((1,)), spatial_shapes.prod(1).cumsum(0)[:-1]))
        valid_ratios = torch.stack([self.get_valid_ratio(m) for m in masks], 1)
        memory = self.encoder(src_flatten, spatial_shapes, level_start_index, valid_ratios, lvl_pos_embed_flatten,
                              mask_flatten, spatial_shapes_list=shapes_list, unit_valid_ratios=True)
        return memory, spatial_shapes, level_start_index

    def forward_tokens(self, src_flatten, lvl_pos_embed_flatten, shapes_list):
        """Token-major entry used by `bm2f_b200.pixel_decoder`: the caller already holds the concatenated encoder input
        (N, S, C) and the level position embedding (1 or N, S, C), so the flatten / transpose / cat of `forward`
        (msdeformattn.py:66-82) is skipped.  Masks are all-False in this decoder (msdeformattn.py:62), hence valid
        ratios are 1 and no padding mask is passed (masked_fill with an all-False mask is the identity)."""
        n, device = src_flatten.shape[0], src_flatten.device
        spatial_shapes = torch.as_tensor(shapes_list, dtype=torch.long, device=device)
        level_start_index = torch.cat((spatial_shapes.new_zeros((1,)), spatial_shapes.prod(1).cumsum(0)[:-1]))
        valid_ratios = torch.ones((n, len(shapes_list), 2), dtype=torch.float32, device=device)
        memory = self.encoder(src_flatten, spatial_shapes, level_start_index, valid_ratios, lvl_pos_embed_flatten,
                              None, spatial_shapes_list=list(shapes_list), unit_valid_ratios=True)
        return memory, spatial_shapes, level_start_index
