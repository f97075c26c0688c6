"""Synthetic Mask2Former-shaped inputs for MSDeformAttn.

All tensors are drawn from a CPU ``torch.Generator`` so that CPU checkers and the
GPU kernels see identical bits.  No dataset, no weights.

Shapes follow the callers of the op in the reference:
  * level order res5 -> res3 (smallest level first), spatial_shapes int64 (L,2)=(H,W),
    level_start_index int64 (L)      mask2former/modeling/pixel_decoder/msdeformattn.py:61-89
  * reference points = pixel centres (i+0.5)/dim, identical for every level
    (valid_ratios == 1)              msdeformattn.py:141-153
  * sampling offsets start at the 8 compass directions scaled by (p+1)
                                     ops/modules/ms_deform_attn.py:66-74
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Tuple

import torch


@dataclass(frozen=True)
class Workload:
    cfg: int
    name: str
    levels: Tuple[Tuple[int, int], ...]   # (H, W), res5 -> res3
    batch: int                            # images (frames) in the job
    dtype: str                            # "f32" | "bf16"
    mode: str                             # "fwd" | "fwd+bwd"
    n_heads: int = 8
    head_dim: int = 32
    n_points: int = 4
    n_layers: int = 6                     # TRANSFORMER_ENC_LAYERS

    @property
    def S(self) -> int:
        return sum(h * w for h, w in self.levels)

    @property
    def L(self) -> int:
        return len(self.levels)


# BASELINE.json configs[0..4] -> cfg 1..5
WORKLOADS = {
    1: Workload(1, "ade20k_r50_512", ((16, 16), (32, 32), (64, 64)), 1, "f32", "fwd"),
    2: Workload(2, "coco_panoptic_r50_1024_lsj", ((32, 32), (64, 64), (128, 128)), 16, "f32", "fwd+bwd"),
    3: Workload(3, "ade20k_swinl_640", ((20, 20), (40, 40), (80, 80)), 16, "bf16", "fwd"),
    4: Workload(4, "cityscapes_swinl_1024x2048", ((32, 64), (64, 128), (128, 256)), 16, "f32", "fwd+bwd"),
    5: Workload(5, "ytvis21_r50_T2_384x640", ((12, 20), (24, 40), (48, 80)), 32, "f32", "fwd+bwd"),
}


def level_tensors(levels, device="cpu"):
    """spatial_shapes (L,2) int64 and level_start_index (L) int64, as the caller builds them."""
    shapes = torch.tensor(list(levels), dtype=torch.long, device=device)
    start = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    return shapes, start


def reference_points(levels, batch: int) -> torch.Tensor:
    """(N, S, L, 2) pixel-centre reference points in (x, y) order, valid_ratios == 1."""
    pts = []
    for h, w in levels:
        ys = (torch.arange(h, dtype=torch.float32) + 0.5) / h
        xs = (torch.arange(w, dtype=torch.float32) + 0.5) / w
        yy, xx = torch.meshgrid(ys, xs, indexing="ij")
        pts.append(torch.stack((xx.reshape(-1), yy.reshape(-1)), -1))
    ref = torch.cat(pts, 0)                                   # (S, 2)
    return ref[None, :, None, :].expand(batch, -1, len(levels), -1).contiguous()


def compass_offset_bias(n_heads: int, n_levels: int, n_points: int) -> torch.Tensor:
    """(M, L, P, 2) initial sampling-offset bias in pixels: head m points along angle
    2*pi*m/M, normalised to the unit square, point p at distance p+1."""
    ang = torch.arange(n_heads, dtype=torch.float32) * (2.0 * math.pi / n_heads)
    d = torch.stack((ang.cos(), ang.sin()), -1)
    d = d / d.abs().max(-1, keepdim=True)[0]
    scale = torch.arange(1, n_points + 1, dtype=torch.float32)
    return (d[:, None, None, :] * scale[None, None, :, None]).expand(-1, n_levels, -1, -1).contiguous()


def make_inputs(levels, batch: int, *, n_heads=8, head_dim=32, n_points=4, seed=0,
                dist="model", n_query=None, with_grad_out=True):
    """Returns dict(value, shapes, start, loc, attn, grad_out) of fp32 CPU tensors.

    dist="model":   loc = pixel-centre ref + (compass bias + N(0,1) px) / (W_l, H_l)
    dist="uniform": loc ~ U(-0.25, 1.25)  (edge / worst-locality stress, parity only)
    n_query:        only for dist="uniform"; default Lq = S (encoder self-attention).
    """
    g = torch.Generator(device="cpu").manual_seed(seed)
    L = len(levels)
    S = sum(h * w for h, w in levels)
    shapes, start = level_tensors(levels)
    value = torch.randn(batch, S, n_heads, head_dim, generator=g)
    if dist == "model":
        Lq = S
        ref = reference_points(levels, batch)                                  # (N,S,L,2)
        off = compass_offset_bias(n_heads, L, n_points)[None, None] + torch.randn(
            batch, Lq, n_heads, L, n_points, 2, generator=g)
        norm = torch.stack((shapes[:, 1], shapes[:, 0]), -1).float()           # (L,2) = (W,H)
        loc = ref[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    elif dist == "uniform":
        Lq = S if n_query is None else n_query
        loc = torch.rand(batch, Lq, n_heads, L, n_points, 2, generator=g) * 1.5 - 0.25
    else:
        raise ValueError(dist)
    attn = torch.softmax(torch.randn(batch, Lq, n_heads, L * n_points, generator=g), -1)
    attn = attn.view(batch, Lq, n_heads, L, n_points)
    out = dict(value=value, shapes=shapes, start=start, loc=loc.contiguous(), attn=attn.contiguous())
    if with_grad_out:
        out["grad_out"] = torch.randn(batch, Lq, n_heads * head_dim, generator=g)
    return out


def workload_inputs(cfg: int, batch=None, dist="model"):
    w = WORKLOADS[cfg]
    return make_inputs(w.levels, w.batch if batch is None else batch, n_heads=w.n_heads,
                       head_dim=w.head_dim, n_points=w.n_points, seed=1234 + cfg, dist=dist)


# ---- algorithmic byte counts per image-layer ---------------------------
def hbm_bytes(S: int, mode: str, dtype: str = "f32", M=8, D=32, L=3, P=4) -> int:
    e = 4 if dtype == "f32" else 2
    C = M * D
    fwd = S * C * e + S * M * L * P * 2 * 4 + S * M * L * P * 4 + S * C * e
    if mode == "fwd":
        return fwd
    bwd = fwd - S * C * e + 3 * S * C * e + S * M * L * P * 3 * 4   # +grad_out, gvalue, zero-fill, gloc, gattn
    return fwd + bwd if mode == "fwd+bwd" else bwd


def gather_bytes(S: int, mode: str, dtype: str = "f32", M=8, D=32, L=3, P=4) -> int:
    e = 4 if dtype == "f32" else 2
    g = S * M * L * P * 4 * D * e
    return {"fwd": g, "bwd": 2 * g, "fwd+bwd": 3 * g}[mode]
