"""Shared test helpers (CPU and GPU suites)."""
import numpy as np


def smooth_mask(loc, shapes, eps=1e-4):
    """True where grad_sampling_loc is well defined in finite precision.

    Bilinear interpolation is piecewise linear: d out / d loc jumps when a pixel coordinate
    (x*W-0.5 or y*H-0.5) crosses an integer, and at the -1 / H in-range borders.  At such kinks
    float32 and float64 may legitimately land on different sides (e.g. 0.8333333f*3-0.5 vs 2.0),
    so those entries are excluded from grad_loc comparisons.  out, grad_value and grad_attn are
    continuous there and are always compared everywhere.
    loc: (N,Lq,M,L,P,2) float64;  returns bool (N,Lq,M,L,P,1)."""
    loc = np.asarray(loc, dtype=np.float64)
    shapes = np.asarray(shapes)
    W = shapes[:, 1].astype(np.float64)[None, None, None, :, None]
    H = shapes[:, 0].astype(np.float64)[None, None, None, :, None]
    w_im = loc[..., 0] * W - 0.5
    h_im = loc[..., 1] * H - 0.5
    ok = (np.abs(w_im - np.round(w_im)) > eps) & (np.abs(h_im - np.round(h_im)) > eps)
    return ok[..., None]


def rel_err(got, ref):
    """max |got-ref| relative to max |ref| (the north_star's 'relative' for gradients)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30))
