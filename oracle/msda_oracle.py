"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

ctypes front-end of the CPU oracle (oracle/msda_oracle.c) for multi-scale
deformable attention, plus `torch_port_forward`, a restatement of the
reference's only CPU path.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
import this module.  Nothing under bm2f_b200/ does.

Reference lines followed (relative to /root/reference/mask2former/modeling/
pixel_decoder/ops/):
  * C oracle ........ src/cuda/ms_deform_im2col_cuda.cuh:38-164,242-304
  * torch_port_* .... functions/ms_deform_attn_func.py:52-72
                      (ms_deform_attn_core_pytorch: per level, view value as an
                      image batch, F.grid_sample(bilinear, zeros,
                      align_corners=False) at 2*loc-1, weight, sum over L*P)

Parity pin: tests/test_oracle_golden.py checks both against tests/golden/*.npz,
which oracle/gen_golden.py produced by importing the reference itself.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmsda_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C oracle with the committed Makefile (gcc only)."""
    src = [os.path.join(_HERE, f) for f in ("msda_oracle.c", "msda_oracle_impl.h", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "libmsda_oracle.so"])
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_LIB_PATH)
        i64p = ctypes.POINTER(ctypes.c_int64)
        for suffix, ct in (("f64", ctypes.c_double), ("f32", ctypes.c_float)):
            rp = ctypes.POINTER(ct)
            f = getattr(lib, "msda_oracle_forward_" + suffix)
            f.restype = None
            f.argtypes = [rp, i64p, i64p, rp, rp] + [ctypes.c_int] * 7 + [rp]
            g = getattr(lib, "msda_oracle_backward_" + suffix)
            g.restype = None
            g.argtypes = [rp, i64p, i64p, rp, rp, rp] + [ctypes.c_int] * 7 + [rp, rp, rp]
        _lib = lib
    return _lib


def _prep(value, shapes, start, loc, attn, dtype):
    value = np.ascontiguousarray(value, dtype=dtype)
    loc = np.ascontiguousarray(loc, dtype=dtype)
    attn = np.ascontiguousarray(attn, dtype=dtype)
    shapes = np.ascontiguousarray(shapes, dtype=np.int64)
    start = np.ascontiguousarray(start, dtype=np.int64)
    N, S, M, D = value.shape
    _, Lq, M2, L, P, two = loc.shape
    assert two == 2 and M2 == M and shapes.shape == (L, 2) and start.shape == (L,)
    assert attn.shape == (N, Lq, M, L, P)
    return value, shapes, start, loc, attn, (N, S, M, D, L, Lq, P)


def _ptr(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def forward(value, shapes, start, loc, attn, dtype=np.float64):
    """Oracle forward.  Returns out (N, Lq, M*D) in `dtype`."""
    lib = _load()
    value, shapes, start, loc, attn, dims = _prep(value, shapes, start, loc, attn, dtype)
    N, S, M, D, L, Lq, P = dims
    ct = ctypes.c_double if dtype == np.float64 else ctypes.c_float
    out = np.empty((N, Lq, M * D), dtype=dtype)
    fn = lib.msda_oracle_forward_f64 if dtype == np.float64 else lib.msda_oracle_forward_f32
    fn(_ptr(value, ct), _ptr(shapes, ctypes.c_int64), _ptr(start, ctypes.c_int64), _ptr(loc, ct),
       _ptr(attn, ct), N, S, M, D, L, Lq, P, _ptr(out, ct))
    return out


def backward(value, shapes, start, loc, attn, grad_out, dtype=np.float64):
    """Oracle backward.  Returns (grad_value, grad_loc, grad_attn)."""
    lib = _load()
    value, shapes, start, loc, attn, dims = _prep(value, shapes, start, loc, attn, dtype)
    N, S, M, D, L, Lq, P = dims
    grad_out = np.ascontiguousarray(grad_out, dtype=dtype).reshape(N, Lq, M * D)
    ct = ctypes.c_double if dtype == np.float64 else ctypes.c_float
    gv = np.empty_like(value)
    gl = np.zeros_like(loc)
    ga = np.zeros_like(attn)
    fn = lib.msda_oracle_backward_f64 if dtype == np.float64 else lib.msda_oracle_backward_f32
    fn(_ptr(value, ct), _ptr(shapes, ctypes.c_int64), _ptr(start, ctypes.c_int64), _ptr(loc, ct),
       _ptr(attn, ct), _ptr(grad_out, ct), N, S, M, D, L, Lq, P, _ptr(gv, ct), _ptr(gl, ct),
       _ptr(ga, ct))
    return gv, gl, ga


# --------------------------------------------------------------------------------------
# Port of the reference's CPU path (the thing `bench.py --impl reference` times).
# --------------------------------------------------------------------------------------
def torch_port_forward(value, shapes, loc, attn):
    """The reference's pure-torch formulation (functions/ms_deform_attn_func.py:52-72)
    restated: every level is viewed as a batch of N*M images with D channels and
    sampled with F.grid_sample; the L*P samples are then blended with the
    attention weights.  Differentiable, runs on any device torch supports.

    value (N,S,M,D)  shapes [(H,W)]*L  loc (N,Lq,M,L,P,2) in [0,1]  attn (N,Lq,M,L,P)
    returns (N, Lq, M*D)
    """
    import torch
    import torch.nn.functional as F

    N, S, M, D = value.shape
    Lq, L, P = loc.shape[1], loc.shape[3], loc.shape[4]
    hw = [(int(h), int(w)) for h, w in (shapes.tolist() if hasattr(shapes, "tolist") else shapes)]
    # (N,S,M,D) -> (N*M, D, S): one "image stack" per (batch, head)
    planes = value.permute(0, 2, 3, 1).reshape(N * M, D, S)
    # grid_sample wants coordinates in [-1,1] and (N*M, Lq, P, 2) per level
    grid = (loc * 2.0 - 1.0).permute(0, 2, 3, 1, 4, 5).reshape(N * M, L, Lq, P, 2)
    blend = attn.permute(0, 2, 1, 3, 4).reshape(N * M, 1, Lq, L, P)
    acc = None
    offset = 0
    for lvl, (h, w) in enumerate(hw):
        img = planes[:, :, offset:offset + h * w].reshape(N * M, D, h, w)
        offset += h * w
        sampled = F.grid_sample(img, grid[:, lvl], mode="bilinear", padding_mode="zeros",
                                align_corners=False)            # (N*M, D, Lq, P)
        term = (sampled * blend[:, :, :, lvl]).sum(-1)           # (N*M, D, Lq)
        acc = term if acc is None else acc + term
    return acc.reshape(N, M * D, Lq).transpose(1, 2).contiguous()


def torch_port_forward_backward(value, shapes, loc, attn, grad_out):
    """fwd + autograd bwd through the torch port; returns (out, gv, gl, ga)."""
    import torch

    value = value.detach().clone().requires_grad_(True)
    loc = loc.detach().clone().requires_grad_(True)
    attn = attn.detach().clone().requires_grad_(True)
    out = torch_port_forward(value, shapes, loc, attn)
    out.backward(grad_out.reshape(out.shape))
    return out.detach(), value.grad, loc.grad, attn.grad


# --------------------------------------------------------------------------------------
# Fused op (softmax + location arithmetic + sampling): float64 numpy on top of the C oracle.
# Restates ops/modules/ms_deform_attn.py:101-109 and their derivatives.
# --------------------------------------------------------------------------------------
def _fused_prologue(shapes, ref, offsets, logits):
    shapes = np.asarray(shapes, dtype=np.int64)
    ref = np.asarray(ref, dtype=np.float64)
    offsets = np.asarray(offsets, dtype=np.float64)
    logits = np.asarray(logits, dtype=np.float64)
    N, Lq, M, L, P, _ = offsets.shape
    z = logits.reshape(N, Lq, M, L * P)
    z = z - z.max(-1, keepdims=True)
    e = np.exp(z)
    attn = (e / e.sum(-1, keepdims=True)).reshape(N, Lq, M, L, P)
    norm = np.stack((shapes[:, 1], shapes[:, 0]), -1).astype(np.float64)          # (L,2) = (W,H)
    loc = ref[:, :, None, :, None, :] + offsets / norm[None, None, None, :, None, :]
    return loc, attn, norm


def fused_forward(value, shapes, start, ref, offsets, logits):
    loc, attn, _ = _fused_prologue(shapes, ref, offsets, logits)
    return forward(value, shapes, start, loc, attn)


def fused_backward(value, shapes, start, ref, offsets, logits, grad_out):
    """Returns (grad_value, grad_offsets, grad_logits)."""
    loc, attn, norm = _fused_prologue(shapes, ref, offsets, logits)
    gv, gl, ga = backward(value, shapes, start, loc, attn, grad_out)
    g_off = gl / norm[None, None, None, :, None, :]
    N, Lq, M, L, P = attn.shape
    a = attn.reshape(N, Lq, M, L * P)
    g = ga.reshape(N, Lq, M, L * P)
    g_logits = (a * (g - (a * g).sum(-1, keepdims=True))).reshape(N, Lq, M, L, P)
    return gv, g_off, g_logits


def sine_position_embedding(height, width, num_pos_feats=128, temperature=10000.0, scale=2 * np.pi, normalize=True):
    """numpy restatement of PositionEmbeddingSine.forward for an all-False mask
    (reference: transformer_decoder/position_encoding.py:29-52), float32 arithmetic in the reference's order.
    Returns (2 * num_pos_feats, height, width) like the reference's `pos[0]`."""
    f32 = np.float32
    y_embed = np.cumsum(np.ones((height, width), dtype=f32), axis=0, dtype=f32)
    x_embed = np.cumsum(np.ones((height, width), dtype=f32), axis=1, dtype=f32)
    if normalize:
        eps = f32(1e-6)
        y_embed = y_embed / (y_embed[-1:, :] + eps) * f32(scale)
        x_embed = x_embed / (x_embed[:, -1:] + eps) * f32(scale)
    dim_t = np.arange(num_pos_feats, dtype=f32)
    dim_t = np.power(f32(temperature), (2 * np.floor(dim_t / 2) / f32(num_pos_feats)).astype(f32)).astype(f32)
    pos_x = x_embed[:, :, None] / dim_t
    pos_y = y_embed[:, :, None] / dim_t
    pos_x = np.stack((np.sin(pos_x[:, :, 0::2]), np.cos(pos_x[:, :, 1::2])), axis=3).reshape(height, width, -1)
    pos_y = np.stack((np.sin(pos_y[:, :, 0::2]), np.cos(pos_y[:, :, 1::2])), axis=3).reshape(height, width, -1)
    return np.concatenate((pos_y, pos_x), axis=2).transpose(2, 0, 1).astype(f32)
