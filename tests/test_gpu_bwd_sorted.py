"""GPU parity tests of the anchor-sorted backward (bm2f_b200/csrc/msda_bwd_sorted.cuh, tuning.bwd = 2), called through
the C ABI, against the CPU oracle, the golden vectors recorded from the reference and the per-corner kernel.

It replaces ms_deformable_col2im_gpu_kernel_shm_blocksize_aware_reduce_v1 + ms_deform_attn_col2im_bilinear
(/root/reference/mask2former/modeling/pixel_decoder/ops/src/cuda/ms_deform_im2col_cuda.cuh:92-164, 306-408) for encoder
self-attention shapes.  Tolerances as in test_gpu_parity.py: gradients <= 1e-4 relative to the largest reference entry.
"""
import glob
import os

import numpy as np
import pytest
import torch

from bm2f_b200 import cabi
from bm2f_b200 import workloads as W
from oracle import msda_oracle as O
from tests.helpers import rel_err, smooth_mask
from tests.test_gpu_parity import (SMALL_LEVELS, _bwd, _dev, _fused_inputs, _fwd, check_f32, oracle_ref, run_cabi)

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def _pyramid(z):
    """golden case usable by the sorted kernel: D = 32, M = 8, P = 4 and queries == pixels"""
    return z["value"].shape[2:] == (8, 32) and z["loc"].shape[4] == 4 and z["loc"].shape[1] == z["value"].shape[1]


_CONFIG_PROBLEMS = {}


def _config_problem(cfg, batch, dist):
    """inputs + CPU-oracle results of a BASELINE config shape, computed once per (cfg, batch, dist) for all kernel variants"""
    key = (cfg, batch, dist)
    if key not in _CONFIG_PROBLEMS:
        inp = W.workload_inputs(cfg, batch=batch, dist=dist)
        _CONFIG_PROBLEMS[key] = (inp, oracle_ref(inp))
    return _CONFIG_PROBLEMS[key]


@pytest.fixture(scope="module")
def small_problem():
    inp = W.make_inputs(SMALL_LEVELS, 3, seed=77)
    return inp, oracle_ref(inp)


# margin 1: nearly every point leaves the window and takes the in-kernel per-corner path; 64: windows = whole levels
@pytest.mark.parametrize("lanes", [8, 4])
@pytest.mark.parametrize("margin", [0, 1, 3, 64])
def test_sorted_backward_vs_oracle(lanes, margin, small_problem, built):
    inp, ref = small_problem
    res = run_cabi(inp["value"], inp["shapes"], inp["start"], inp["loc"], inp["attn"], inp["grad_out"],
                   tuning=cabi.make_tuning(bwd=2, bwd_lanes=lanes, bwd_margin=margin))
    check_f32(res, ref, inp["loc"].numpy(), inp["shapes"].numpy(), f"lanes={lanes} margin={margin}")


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_sorted_backward_golden(path, built):
    z = np.load(path)
    if not _pyramid(z):
        pytest.skip("not an encoder self-attention shape (the library uses the per-corner / generic kernels)")
    res = run_cabi(z["value"], z["shapes"], z["start"], z["loc"], z["attn"], z["grad_out"], tuning=cabi.make_tuning(bwd=2))
    check_f32(res, z, z["loc"], z["shapes"], path)


@pytest.mark.parametrize("levels", [((7, 9),), ((3, 5), (7, 9)), ((3, 5), (7, 9), (16, 16), (5, 33)),
                                    ((1, 1), (3, 2), (25, 38)), ((2, 70), (4, 140), (1, 35))],
                         ids=["L1", "L2", "L4", "one_pixel_level", "wide_strips"])
@pytest.mark.parametrize("dist", ["model", "uniform"])
def test_sorted_backward_level_counts_and_ragged_shapes(levels, dist, built):
    inp = W.make_inputs(levels, 2, seed=300 + len(levels), dist=dist)
    ref = oracle_ref(inp)
    res = run_cabi(inp["value"], inp["shapes"], inp["start"], inp["loc"], inp["attn"], inp["grad_out"],
                   tuning=cabi.make_tuning(bwd=2))
    check_f32(res, ref, inp["loc"].numpy(), inp["shapes"].numpy(), f"{levels} {dist}")


@pytest.mark.parametrize("variant", [0, 6, 8, 9, 10, 11, 12, 13, 14, 15, 16])
@pytest.mark.parametrize("cfg,batch,dist", [(1, 1, "model"), (1, 1, "uniform"), (2, 2, "model"), (4, 1, "model"),
                                            (5, 4, "model")])
def test_sorted_backward_config_shapes_vs_oracle(cfg, batch, dist, variant, built):
    inp, ref = _config_problem(cfg, batch, dist)
    res = run_cabi(inp["value"], inp["shapes"], inp["start"], inp["loc"], inp["attn"], inp["grad_out"],
                   tuning=cabi.make_tuning(bwd=2, variant=variant))
    check_f32(res, ref, inp["loc"].numpy(), inp["shapes"].numpy(), f"cfg{cfg} variant {variant}")


@pytest.mark.parametrize("variant", [6, 8, 9, 10, 11, 12, 13, 14])
@pytest.mark.parametrize("margin", [0, 1, 3, 64])
def test_pixel_owner_backward_vs_oracle(variant, margin, small_problem, built):
    inp, ref = small_problem
    res = run_cabi(inp["value"], inp["shapes"], inp["start"], inp["loc"], inp["attn"], inp["grad_out"],
                   tuning=cabi.make_tuning(bwd=2, bwd_margin=margin, variant=variant))
    check_f32(res, ref, inp["loc"].numpy(), inp["shapes"].numpy(), f"variant={variant} margin={margin}")


@pytest.mark.parametrize("levels", [((1, 1), (3, 2), (25, 38)), ((2, 70), (4, 140), (1, 35)), ((33, 31), (5, 9), (17, 64))],
                         ids=["one_pixel_level", "wide_strips", "unordered_sizes"])
@pytest.mark.parametrize("dist", ["model", "uniform"])
@pytest.mark.parametrize("variant", [8, 10, 13, 14])
def test_pixel_owner_backward_ragged_shapes(levels, dist, variant, built):
    inp = W.make_inputs(levels, 2, seed=400 + levels[0][0], dist=dist)
    ref = oracle_ref(inp)
    res = run_cabi(inp["value"], inp["shapes"], inp["start"], inp["loc"], inp["attn"], inp["grad_out"],
                   tuning=cabi.make_tuning(bwd=2, variant=variant))
    check_f32(res, ref, inp["loc"].numpy(), inp["shapes"].numpy(), f"{levels} {dist}")


def test_sorted_backward_rejects_what_it_does_not_cover(built):
    inp = W.make_inputs(SMALL_LEVELS, 1, seed=5, dist="uniform", n_query=77)          # Lq != S
    with pytest.raises(cabi.MSDAError, match="anchor-sorted"):
        run_cabi(inp["value"], inp["shapes"], inp["start"], inp["loc"], inp["attn"], inp["grad_out"],
                 tuning=cabi.make_tuning(bwd=2))
    inp = W.make_inputs(SMALL_LEVELS, 1, seed=5)
    with pytest.raises(cabi.MSDAError, match="anchor-sorted"):
        run_cabi(inp["value"], inp["shapes"], inp["start"], inp["loc"], inp["attn"], inp["grad_out"],
                 tuning=cabi.make_tuning(bwd=2, order=1))


@pytest.mark.parametrize("levels,batch,null_ref", [(SMALL_LEVELS, 3, False), (SMALL_LEVELS, 3, True),
                                                   (((5, 7), (9, 13)), 2, True), (W.WORKLOADS[1].levels, 1, True),
                                                   (W.WORKLOADS[2].levels, 1, False)])
@pytest.mark.parametrize("lanes", [8, 4])
def test_sorted_fused_backward_vs_oracle(levels, batch, null_ref, lanes, built):
    base, ref, offsets, logits = _fused_inputs(levels, batch, 900 + len(levels))
    dev = _dev()
    sh, st = base["shapes"].to(dev), base["start"].to(dev)
    v, r, o, lg, go = (t.to(dev).contiguous() for t in (base["value"], ref, offsets, logits, base["grad_out"]))
    N, S, M, D = v.shape
    L = len(levels)
    dims = (N, S, M, D, L, S, 4)
    gv, goff, glog = torch.full_like(v, float("nan")), torch.full_like(o, float("nan")), torch.full_like(lg, float("nan"))
    stream = torch.cuda.current_stream().cuda_stream
    cabi.fused_backward(v.data_ptr(), sh.data_ptr(), st.data_ptr(), 0 if null_ref else r.data_ptr(), o.data_ptr(),
                        lg.data_ptr(), go.data_ptr(), gv.data_ptr(), goff.data_ptr(), glog.data_ptr(), dims,
                        cabi.DTYPE_F32, cabi.make_tuning(bwd=2, bwd_lanes=lanes), stream)
    torch.cuda.synchronize()
    a = [t.numpy() for t in (base["value"], base["shapes"], base["start"], ref, offsets, logits)]
    rgv, rgo, rgl = O.fused_backward(*a, base["grad_out"].numpy())
    loc = ref.numpy()[:, :, None, :, None, :] + offsets.numpy() / np.stack(
        (base["shapes"].numpy()[:, 1], base["shapes"].numpy()[:, 0]), -1)[None, None, None, :, None, :]
    ok = smooth_mask(loc, base["shapes"].numpy(), eps=1e-3)
    assert rel_err(gv.cpu().numpy(), rgv) <= 1e-4
    assert rel_err(glog.cpu().numpy(), rgl) <= 1e-4
    assert rel_err(goff.cpu().numpy() * ok, rgo * ok) <= 1e-4


# ----------------------------------------------------------------------------------------------
# full BASELINE size (cfg 2, N = 16): the two backward kernels against each other, adjoint identity
# ----------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def full_problem():
    inp = W.workload_inputs(2)
    dev = _dev()
    return {k: v.to(dev) for k, v in inp.items()}


@pytest.mark.parametrize("lanes", [8, 4])
def test_full_size_sorted_equals_per_corner(lanes, full_problem, built):
    d = full_problem
    a = _bwd(d, d["grad_out"], tuning=cabi.make_tuning(bwd=2, bwd_lanes=lanes))
    b = _bwd(d, d["grad_out"], tuning=cabi.make_tuning(bwd=1))
    for x, y, name in zip(a, b, ("grad_value", "grad_loc", "grad_attn")):
        assert torch.isfinite(x).all(), name
        assert (x - y).abs().max().item() <= 2e-5 * y.abs().max().item(), name


def test_full_size_sorted_adjoint_identity(full_problem, built):
    d = full_problem
    g = d["grad_out"]
    out = _fwd(d)
    gv, gl, ga = _bwd(d, g, tuning=cabi.make_tuning(bwd=2))
    lhs = (out.double() * g.double()).sum().item()
    rhs = (d["value"].double() * gv.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-6 * max(abs(lhs), 1.0) + 1e-2
    mid = (ga.double() * d["attn"].double()).sum().item()
    assert abs(lhs - mid) <= 1e-6 * max(abs(lhs), 1.0) + 1e-2
