"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference):

    python oracle/gen_golden.py

What is recorded: for each seeded case, the inputs and the float64 outputs of the
reference's `ms_deform_attn_core_pytorch` (ops/functions/ms_deform_attn_func.py:52-72)
plus its autograd gradients for a seeded grad_output.  This is the reference's own
correctness anchor: ops/test.py:34-63 compares the CUDA op against exactly this
function, and `gradcheck` (test.py:66-81) differentiates it numerically.

The reference holds no golden vectors of its own, so these files are
the pin for oracle/msda_oracle.c and, through it, for the CUDA kernels.

The reference imports a module named `MultiScaleDeformableAttention` at import time
(func.py:21-29); a stub is registered because only the pure-torch function is used.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

REF_OPS_PARENT = "/root/reference/mask2former/modeling/pixel_decoder"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def import_reference():
    if not os.path.isdir(REF_OPS_PARENT):
        raise SystemExit("reference tree not present: " + REF_OPS_PARENT)
    sys.modules.setdefault("MultiScaleDeformableAttention", types.ModuleType("MultiScaleDeformableAttention"))
    sys.path.insert(0, REF_OPS_PARENT)
    from ops.functions.ms_deform_attn_func import ms_deform_attn_core_pytorch  # noqa
    return ms_deform_attn_core_pytorch


def level_tensors(levels):
    shapes = torch.tensor(list(levels), dtype=torch.long)
    start = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    return shapes, start


def run_case(ref_fn, name, levels, N, M, D, Lq, P, seed, loc_kind):
    g = torch.Generator().manual_seed(seed)
    shapes, start = level_tensors(levels)
    L = len(levels)
    S = int(shapes.prod(1).sum())
    # inputs are drawn in float32 (so they can be stored as float32 and fed bit-identically to
    # the fp32 kernels) and promoted to float64 for the reference run, like test.py:41 does.
    f32 = dict(generator=g, dtype=torch.float32)
    value = (torch.rand(N, S, M, D, **f32) - 0.5).double()
    if loc_kind == "unit":          # ops/test.py style: strictly inside [0,1)
        loc = torch.rand(N, Lq, M, L, P, 2, **f32).double()
    elif loc_kind == "wide":        # 1/3 of coordinates outside the image
        loc = (torch.rand(N, Lq, M, L, P, 2, **f32) * 1.5 - 0.25).double()
    elif loc_kind == "edges":       # exact pixel centres, borders, -1/H corner cases
        loc = (torch.rand(N, Lq, M, L, P, 2, **f32) * 1.5 - 0.25).double()
        flat = loc.view(-1, 2)
        H0, W0 = levels[0]
        special = torch.tensor([
            [0.0, 0.0], [1.0, 1.0], [0.5 / W0, 0.5 / H0], [1.0 - 0.5 / W0, 1.0 - 0.5 / H0],
            [-0.5 / W0, 0.3], [0.3, -0.5 / H0], [1.0 + 0.5 / W0, 0.7], [0.7, 1.0 + 0.5 / H0],
            [-0.4999 / W0, 0.5], [1.0 + 0.4999 / W0, 0.5], [0.25, 0.75], [2.0, 2.0], [-1.0, 0.5],
        ], dtype=torch.float32).double()
        flat[: special.shape[0]] = special
    else:
        raise ValueError(loc_kind)
    attn = torch.rand(N, Lq, M, L, P, **f32) + 1e-5
    attn = (attn / attn.sum(-1, keepdim=True).sum(-2, keepdim=True)).double()
    grad_out = (torch.rand(N, Lq, M * D, **f32) - 0.5).double()

    v = value.clone().requires_grad_(True)
    lo = loc.clone().requires_grad_(True)
    a = attn.clone().requires_grad_(True)
    out = ref_fn(v, shapes, lo, a)
    out.backward(grad_out)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        value=value.float().numpy(), shapes=shapes.numpy(), start=start.numpy(),
        loc=loc.float().numpy(), attn=attn.float().numpy(), grad_out=grad_out.float().numpy(),
        out=out.detach().numpy(),
        grad_value=v.grad.numpy(), grad_loc=lo.grad.numpy(), grad_attn=a.grad.numpy(),
    )
    print(f"{name}: N={N} S={S} M={M} D={D} L={L} Lq={Lq} P={P} out|max|={out.abs().max():.4f}")


CASES = [
    # name,                 levels,                     N, M, D,  Lq, P, seed, loc
    ("testpy_shape_d2",     ((6, 4), (3, 2)),           1, 2, 2,   2, 2, 3,  "unit"),   # ops/test.py:24-31
    ("testpy_shape_d32",    ((6, 4), (3, 2)),           1, 2, 32,  2, 2, 4,  "unit"),
    ("testpy_shape_d30",    ((6, 4), (3, 2)),           1, 2, 30,  2, 2, 5,  "unit"),   # test.py:88 channel list
    ("testpy_shape_d71",    ((6, 4), (3, 2)),           1, 2, 71,  2, 2, 6,  "unit"),
    ("m2f_tiny_d32_wide",   ((2, 2), (4, 4), (8, 8)),   2, 8, 32, 84, 4, 7, "wide"),     # Lq == S, M2F head layout
    ("m2f_tiny_d32_edges",  ((2, 3), (4, 6), (8, 12)),  1, 8, 32, 126, 4, 8, "edges"),
    ("ragged_levels_d32",   ((3, 5), (7, 2), (1, 9), (6, 6)), 2, 4, 32, 37, 4, 9, "wide"),  # L=4, odd shapes, Lq != S
    ("one_pixel_level",     ((1, 1), (2, 3)),           2, 3, 8,   5, 3, 10, "wide"),   # 1x1 level, P=3, D=8
    ("d64_p1",              ((5, 5),),                  2, 2, 64,  9, 1, 11, "wide"),   # single level, P=1
]


def main():
    os.makedirs(OUT, exist_ok=True)
    ref_fn = import_reference()
    for c in CASES:
        run_case(ref_fn, *c)


if __name__ == "__main__":
    main()
