"""CPU: pins the oracle (oracle/msda_oracle.c and the torch port) to the golden vectors that
oracle/gen_golden.py produced by running the reference's own ms_deform_attn_core_pytorch."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import msda_oracle as O
from tests.helpers import smooth_mask

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def test_golden_files_present():
    assert len(GOLDEN) >= 9


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_c_oracle_f64_matches_reference(path):
    z = np.load(path)
    out = O.forward(z["value"], z["shapes"], z["start"], z["loc"], z["attn"])
    gv, gl, ga = O.backward(z["value"], z["shapes"], z["start"], z["loc"], z["attn"], z["grad_out"])
    assert np.abs(out - z["out"]).max() < 1e-13
    assert np.abs(gv - z["grad_value"]).max() < 1e-13
    assert np.abs(gl - z["grad_loc"]).max() < 1e-12   # scaled by H, W
    assert np.abs(ga - z["grad_attn"]).max() < 1e-13


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_c_oracle_f32_close(path):
    z = np.load(path)
    out = O.forward(z["value"], z["shapes"], z["start"], z["loc"], z["attn"], dtype=np.float32)
    gv, gl, ga = O.backward(z["value"], z["shapes"], z["start"], z["loc"], z["attn"], z["grad_out"],
                            dtype=np.float32)
    assert np.abs(out - z["out"]).max() < 1e-5
    ok = smooth_mask(z["loc"], z["shapes"])          # grad_loc is one-sided at pixel-grid kinks
    for got, ref in ((gv, z["grad_value"]), (gl * ok, z["grad_loc"] * ok), (ga, z["grad_attn"])):
        assert np.abs(got - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_torch_port_matches_reference(path):
    z = np.load(path)
    t = {k: torch.from_numpy(z[k]).double() for k in ("value", "loc", "attn", "grad_out")}
    out, gv, gl, ga = O.torch_port_forward_backward(t["value"], z["shapes"], t["loc"], t["attn"], t["grad_out"])
    assert (out.numpy() - z["out"]).__abs__().max() < 1e-13
    assert (gv.numpy() - z["grad_value"]).__abs__().max() < 1e-13
    assert (gl.numpy() - z["grad_loc"]).__abs__().max() < 1e-12
    assert (ga.numpy() - z["grad_attn"]).__abs__().max() < 1e-13


def test_oracle_skipped_points_leave_zero_grads():
    # a location far outside the image must contribute nothing and receive zero gradients
    value = np.random.default_rng(0).standard_normal((1, 4, 1, 4))
    shapes = np.array([[2, 2]], dtype=np.int64)
    start = np.array([0], dtype=np.int64)
    loc = np.full((1, 1, 1, 1, 1, 2), 3.0)
    attn = np.ones((1, 1, 1, 1, 1))
    out = O.forward(value, shapes, start, loc, attn)
    gv, gl, ga = O.backward(value, shapes, start, loc, attn, np.ones((1, 1, 4)))
    assert not out.any() and not gv.any() and not gl.any() and not ga.any()


def test_oracle_pixel_centre_reproduces_value():
    # sampling exactly at a pixel centre with weight 1 returns that pixel
    rng = np.random.default_rng(1)
    H, W = 3, 5
    value = rng.standard_normal((1, H * W, 2, 8))
    shapes = np.array([[H, W]], dtype=np.int64)
    start = np.array([0], dtype=np.int64)
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    loc = np.stack(((xs.reshape(-1) + 0.5) / W, (ys.reshape(-1) + 0.5) / H), -1)
    loc = np.broadcast_to(loc[None, :, None, None, None, :], (1, H * W, 2, 1, 1, 2)).copy()
    attn = np.ones((1, H * W, 2, 1, 1))
    out = O.forward(value, shapes, start, loc, attn)
    assert np.abs(out.reshape(value.shape) - value).max() < 1e-12
