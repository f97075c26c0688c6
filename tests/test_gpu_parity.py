"""GPU parity tests: the sm_100a kernels, called through the C ABI (bm2f_b200.cabi -> libbm2f_msda.so),
against the CPU oracle and the golden vectors recorded from the reference.

Tolerances (BASELINE.json north_star): forward max-abs <= 1e-5 in fp32 (<= 1e-2 relative in bf16),
gradients <= 1e-4 relative in fp32 (relative to the largest reference entry), atomic order is free.
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

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
FWD_ABS_F32 = 1e-5
GRAD_REL_F32 = 1e-4

TORCH_DT = {cabi.DTYPE_F32: torch.float32, cabi.DTYPE_F64: torch.float64, cabi.DTYPE_BF16: torch.bfloat16}


def _dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def run_cabi(value, shapes, start, loc, attn, grad_out=None, dtype=cabi.DTYPE_F32, tuning=None):
    """CPU tensors/arrays in -> device -> C ABI -> CPU float64 numpy out.  bf16: value-side tensors
    are bf16, loc/attn stay fp32 (include/bm2f_msda.h)."""
    dev = _dev()
    vt = TORCH_DT[dtype]
    lt = torch.float64 if dtype == cabi.DTYPE_F64 else torch.float32
    t = lambda a, d: torch.as_tensor(np.asarray(a)).to(device=dev, dtype=d).contiguous()
    v, lo, at = t(value, vt), t(loc, lt), t(attn, lt)
    sh = torch.as_tensor(np.asarray(shapes)).to(dev, torch.long).contiguous()
    st = torch.as_tensor(np.asarray(start)).to(dev, torch.long).contiguous()
    N, S, M, D = v.shape
    Lq, L, P = lo.shape[1], lo.shape[3], lo.shape[4]
    dims = (N, S, M, D, L, Lq, P)
    out = torch.full((N, Lq, M * D), float("nan"), device=dev, dtype=vt)   # must be fully overwritten
    stream = torch.cuda.current_stream().cuda_stream
    cabi.forward(v.data_ptr(), sh.data_ptr(), st.data_ptr(), lo.data_ptr(), at.data_ptr(), out.data_ptr(), dims,
                 dtype, tuning, stream)
    res = {"out": out.double().cpu().numpy()}
    if grad_out is not None:
        go = t(grad_out, vt).reshape(N, Lq, M * D)
        # library must zero-fill; bf16 values accumulate grad_value in fp32 (include/bm2f_msda.h)
        gv = torch.full(v.shape, float("nan"), device=dev, dtype=torch.float64 if dtype == cabi.DTYPE_F64 else torch.float32)
        gl = torch.full_like(lo, float("nan"))     # library must fully overwrite
        ga = torch.full_like(at, float("nan"))
        cabi.backward(v.data_ptr(), sh.data_ptr(), st.data_ptr(), lo.data_ptr(), at.data_ptr(), go.data_ptr(),
                      gv.data_ptr(), gl.data_ptr(), ga.data_ptr(), dims, dtype, tuning, stream)
        res.update(grad_value=gv.double().cpu().numpy(), grad_loc=gl.double().cpu().numpy(),
                   grad_attn=ga.double().cpu().numpy())
    torch.cuda.synchronize()
    return res


def check_f32(res, ref, loc, shapes, tag=""):
    assert np.isfinite(res["out"]).all(), tag
    assert np.abs(res["out"] - ref["out"]).max() <= FWD_ABS_F32 * max(1.0, np.abs(ref["out"]).max()), tag
    if "grad_value" in res:
        ok = smooth_mask(loc, shapes)
        assert rel_err(res["grad_value"], ref["grad_value"]) <= GRAD_REL_F32, tag
        assert rel_err(res["grad_attn"], ref["grad_attn"]) <= GRAD_REL_F32, tag
        assert rel_err(res["grad_loc"] * ok, ref["grad_loc"] * ok) <= GRAD_REL_F32, tag
        assert np.isfinite(res["grad_loc"]).all() and np.isfinite(res["grad_value"]).all(), tag


def oracle_ref(inp):
    a = {k: v.numpy() if isinstance(v, torch.Tensor) else v for k, v in inp.items()}
    out = O.forward(a["value"], a["shapes"], a["start"], a["loc"], a["attn"])
    gv, gl, ga = O.backward(a["value"], a["shapes"], a["start"], a["loc"], a["attn"], a["grad_out"])
    return dict(out=out, grad_value=gv, grad_loc=gl, grad_attn=ga)


# ----------------------------------------------------------------------------------------------
# 1. golden vectors recorded from the reference (tests/golden, oracle/gen_golden.py)
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_golden_f32(path, built):
    z = np.load(path)
    res = run_cabi(z["value"], z["shapes"], z["start"], z["loc"], z["attn"], z["grad_out"])
    check_f32(res, z, z["loc"], z["shapes"], path)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_golden_f64(path, built):
    # float64 goes through the any-shape kernels, like the reference's own double test (test.py:34-47)
    z = np.load(path)
    res = run_cabi(z["value"], z["shapes"], z["start"], z["loc"], z["attn"], z["grad_out"], dtype=cabi.DTYPE_F64)
    assert np.abs(res["out"] - z["out"]).max() < 1e-12
    assert np.abs(res["grad_value"] - z["grad_value"]).max() < 1e-12
    assert np.abs(res["grad_attn"] - z["grad_attn"]).max() < 1e-12
    assert np.abs(res["grad_loc"] - z["grad_loc"]).max() < 1e-11


@pytest.mark.parametrize("path", [p for p in GOLDEN if "m2f_tiny" in p],
                         ids=[os.path.basename(p)[:-4] for p in GOLDEN if "m2f_tiny" in p])
def test_golden_f32_generic_kernel(path, built):
    z = np.load(path)
    res = run_cabi(z["value"], z["shapes"], z["start"], z["loc"], z["attn"], z["grad_out"],
                   tuning=cabi.make_tuning(force_generic=1))
    check_f32(res, z, z["loc"], z["shapes"], path)


# ----------------------------------------------------------------------------------------------
# 2. every fast-path variant against the oracle on a Mask2Former-shaped problem
# ----------------------------------------------------------------------------------------------
SMALL_LEVELS = ((6, 10), (12, 20), (24, 40))     # cfg-5 aspect ratio, widths not multiples of 16/32


@pytest.fixture(scope="module")
def small_problem():
    inp = W.make_inputs(SMALL_LEVELS, 3, seed=77)
    return inp, oracle_ref(inp)


VARIANTS = [dict(vec=v, staging=s, strip_w=sw, ctas_per_sm=c, merge=1)
            for v in (8, 4, 2, 1) for s in (1, 2) for sw in (8, 16, 32) for c in (1, 2)]
VARIANTS += [dict(vec=4, staging=s, strip_w=sw, ctas_per_sm=1, merge=2) for s in (1, 2) for sw in (16, 32)]


@pytest.mark.parametrize("var", VARIANTS,
                         ids=lambda d: "v{vec}_st{staging}_sw{strip_w}_c{ctas_per_sm}_m{merge}".format(**d))
def test_fast_variants_vs_oracle(var, small_problem, built):
    if (var["vec"] in (1, 2) or var["strip_w"] == 8) and b"sweep" not in cabi.lib().bm2f_msda_build_info():
        pytest.skip("sweep-only variant (BM2F_SWEEP=1 python -m bm2f_b200.build)")
    inp, ref = small_problem
    res = run_cabi(inp["value"], inp["shapes"], inp["start"], inp["loc"], inp["attn"], inp["grad_out"],
                   tuning=cabi.make_tuning(**var))
    check_f32(res, ref, inp["loc"].numpy(), inp["shapes"].numpy(), str(var))


@pytest.mark.parametrize("rows", [1, 4, 7, 64])
@pytest.mark.parametrize("order", [0, 1])
def test_job_shapes_do_not_change_results(rows, order, small_problem, built):
    inp, ref = small_problem
    res = run_cabi(inp["value"], inp["shapes"], inp["start"], inp["loc"], inp["attn"], inp["grad_out"],
                   tuning=cabi.make_tuning(rows=rows, order=order))
    check_f32(res, ref, inp["loc"].numpy(), inp["shapes"].numpy(), f"rows={rows} order={order}")


@pytest.mark.parametrize("L", [1, 2, 4])
@pytest.mark.parametrize("staging", [1, 2])
def test_other_level_counts(L, staging, built):
    levels = ((3, 5), (7, 9), (16, 16), (5, 33))[:L]
    inp = W.make_inputs(levels, 2, seed=100 + L, dist="uniform", n_query=123)       # Lq != S: 1-D order
    ref = oracle_ref(inp)
    res = run_cabi(inp["value"], inp["shapes"], inp["start"], inp["loc"], inp["attn"], inp["grad_out"],
                   tuning=cabi.make_tuning(staging=staging))
    check_f32(res, ref, inp["loc"].numpy(), inp["shapes"].numpy(), f"L={L}")


# ----------------------------------------------------------------------------------------------
# 3. BASELINE.json config shapes at sizes the oracle finishes in seconds
# ----------------------------------------------------------------------------------------------
# cfg 2 at batch 2 and cfg 5 at its frames-as-batch 4: image b > 0 of the headline shapes meets the oracle (cross-image
# indexing); cfg 4 (Cityscapes, S = 43008, W up to 256 = 8 strips per row) in both distributions
@pytest.mark.parametrize("cfg,batch,dist", [(1, 1, "model"), (5, 4, "model"), (2, 2, "model"), (1, 1, "uniform"),
                                            (3, 1, "model"), (4, 1, "model"), (4, 1, "uniform")])
def test_config_shapes_vs_oracle(cfg, batch, dist, built):
    inp = W.workload_inputs(cfg, batch=batch, dist=dist)
    ref = oracle_ref(inp)
    res = run_cabi(inp["value"], inp["shapes"], inp["start"], inp["loc"], inp["attn"], inp["grad_out"])
    check_f32(res, ref, inp["loc"].numpy(), inp["shapes"].numpy(), f"cfg{cfg}")


def test_bf16_forward_and_backward_vs_oracle(built):
    # config 3 is bf16 inference; oracle runs in float64 on the bf16-rounded value
    inp = W.workload_inputs(3, batch=2)
    inp["value"] = inp["value"].bfloat16().float()
    inp["grad_out"] = inp["grad_out"].bfloat16().float()
    ref = oracle_ref(inp)
    res = run_cabi(inp["value"], inp["shapes"], inp["start"], inp["loc"], inp["attn"], inp["grad_out"],
                   dtype=cabi.DTYPE_BF16)
    assert rel_err(res["out"], ref["out"]) <= 1e-2
    ok = smooth_mask(inp["loc"].numpy(), inp["shapes"].numpy())
    assert rel_err(res["grad_attn"], ref["grad_attn"]) <= 1e-4      # fp32 outputs
    assert rel_err(res["grad_loc"] * ok, ref["grad_loc"] * ok) <= 1e-4
    assert rel_err(res["grad_value"], ref["grad_value"]) <= 1e-4    # fp32 accumulation of bf16 products


# ----------------------------------------------------------------------------------------------
# 4. full BASELINE size (cfg 2, N = 16): size-independent properties, no oracle
# ----------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def full_problem():
    inp = W.workload_inputs(2)
    dev = _dev()
    return {k: v.to(dev) for k, v in inp.items()}


def _fwd(d, value=None, tuning=None):
    v = d["value"] if value is None else value
    N, S, M, D = v.shape
    Lq = d["loc"].shape[1]
    out = torch.empty(N, Lq, M * D, device=v.device)
    cabi.forward(v.data_ptr(), d["shapes"].data_ptr(), d["start"].data_ptr(), d["loc"].data_ptr(),
                 d["attn"].data_ptr(), out.data_ptr(), (N, S, M, D, 3, Lq, 4), cabi.DTYPE_F32, tuning,
                 torch.cuda.current_stream().cuda_stream)
    return out


def _bwd(d, grad_out, tuning=None):
    v = d["value"]
    N, S, M, D = v.shape
    Lq = d["loc"].shape[1]
    gv, gl, ga = torch.empty_like(v), torch.empty_like(d["loc"]), torch.empty_like(d["attn"])
    cabi.backward(v.data_ptr(), d["shapes"].data_ptr(), d["start"].data_ptr(), d["loc"].data_ptr(),
                  d["attn"].data_ptr(), grad_out.data_ptr(), gv.data_ptr(), gl.data_ptr(), ga.data_ptr(),
                  (N, S, M, D, 3, Lq, 4), cabi.DTYPE_F32, tuning, torch.cuda.current_stream().cuda_stream)
    return gv, gl, ga


def test_full_size_linearity_in_value(full_problem, built):
    d = full_problem
    v2 = torch.randn_like(d["value"])
    lhs = _fwd(d, 0.5 * d["value"] + 2.0 * v2)
    rhs = 0.5 * _fwd(d) + 2.0 * _fwd(d, v2)
    assert (lhs - rhs).abs().max().item() <= 1e-4


def test_full_size_constant_value_gives_coverage(full_problem, built):
    # value == 1 everywhere: out = sum of attention weights x bilinear coverage, in [0, 1]; equal
    # to 1 wherever every sampling point lies at least one pixel inside its level
    d = full_problem
    out = _fwd(d, torch.ones_like(d["value"]))
    assert out.min().item() >= -1e-6 and out.max().item() <= 1 + 1e-5
    shapes = d["shapes"].float()
    wh = torch.stack((shapes[:, 1], shapes[:, 0]), -1)[None, None, None, :, None, :]
    px = d["loc"] * wh - 0.5
    inside = ((px >= 0) & (px <= wh - 1)).all(-1).all(-1).all(-1)            # (N, Lq, M)
    full = out.view(*inside.shape, 32)[inside]
    assert full.numel() > 0 and (full - 1).abs().max().item() <= 1e-5


def test_full_size_adjoint_identity(full_problem, built):
    # forward is linear in value, so <A v, g> == <v, A^T g>: ties the scatter (backward) to the gather
    d = full_problem
    g = d["grad_out"]
    out = _fwd(d)
    gv, gl, ga = _bwd(d, g)
    lhs = (out.double() * g.double()).sum().item()
    rhs = (d["value"].double() * gv.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-6 * max(abs(lhs), 1.0) + 1e-2
    # grad_attn is the directional derivative in the attention weights: <grad_attn, attn> == <out, g>
    mid = (ga.double() * d["attn"].double()).sum().item()
    assert abs(lhs - mid) <= 1e-6 * max(abs(lhs), 1.0) + 1e-2


def test_full_size_fast_equals_generic(full_problem, built):
    # two independent kernel families on the full problem
    d = full_problem
    gen = cabi.make_tuning(force_generic=1)
    out_f, out_g = _fwd(d), _fwd(d, tuning=gen)
    assert (out_f - out_g).abs().max().item() <= 1e-5 * max(1.0, out_g.abs().max().item())
    gf, gg = _bwd(d, d["grad_out"]), _bwd(d, d["grad_out"], tuning=gen)
    for a, b in zip(gf[:1] + gf[2:], gg[:1] + gg[2:]):
        assert (a - b).abs().max().item() <= 1e-4 * b.abs().max().item()


def test_full_size_staging_variants_agree_bitwise_forward(full_problem, built):
    # TMA-staged and directly loaded sampling metadata feed the same arithmetic: forward is bit-equal
    d = full_problem
    a = _fwd(d, tuning=cabi.make_tuning(staging=1))
    b = _fwd(d, tuning=cabi.make_tuning(staging=2))
    assert torch.equal(a, b)
    # 256-bit gathers change the summation order only
    c = _fwd(d, tuning=cabi.make_tuning(vec=8))
    assert (a - c).abs().max().item() <= 1e-5 * max(1.0, a.abs().max().item())


def test_full_size_merged_scatter_matches_unmerged(full_problem, built):
    # warp-aggregated REDs (equal-pixel corners merged) vs one RED per corner
    d = full_problem
    a = _bwd(d, d["grad_out"], tuning=cabi.make_tuning(merge=1))
    b = _bwd(d, d["grad_out"], tuning=cabi.make_tuning(merge=2))
    assert (a[0] - b[0]).abs().max().item() <= 1e-4 * b[0].abs().max().item()
    assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])       # grad_loc / grad_attn untouched by merging


# ----------------------------------------------------------------------------------------------
# 5. host-buffer entry
# ----------------------------------------------------------------------------------------------
def test_host_entry_matches_device_entry(built):
    inp = W.workload_inputs(5, batch=5)                    # odd batch: ragged last chunk
    pin = {k: v.contiguous().pin_memory() for k, v in inp.items()}
    N, S, M, D = inp["value"].shape
    Lq = inp["loc"].shape[1]
    out = torch.empty(N, Lq, M * D).pin_memory()
    gv, gl, ga = (torch.empty_like(inp[k]).pin_memory() for k in ("value", "loc", "attn"))
    cabi.forward_backward_host(pin["value"].data_ptr(), pin["shapes"].data_ptr(), pin["start"].data_ptr(),
                               pin["loc"].data_ptr(), pin["attn"].data_ptr(), pin["grad_out"].data_ptr(),
                               out.data_ptr(), gv.data_ptr(), gl.data_ptr(), ga.data_ptr(),
                               (N, S, M, D, 3, Lq, 4))
    ref = run_cabi(inp["value"], inp["shapes"], inp["start"], inp["loc"], inp["attn"], inp["grad_out"])
    assert np.abs(out.double().numpy() - ref["out"]).max() <= 1e-6
    assert rel_err(gv.numpy(), ref["grad_value"]) <= 1e-5
    assert rel_err(gl.numpy(), ref["grad_loc"]) <= 1e-5
    assert rel_err(ga.numpy(), ref["grad_attn"]) <= 1e-5
    # the workspace survives between calls; releasing it and calling again re-creates it
    assert cabi.lib().bm2f_msda_release_host_workspace() == 0
    out2 = torch.empty_like(out)
    cabi.forward_backward_host(pin["value"].data_ptr(), pin["shapes"].data_ptr(), pin["start"].data_ptr(),
                               pin["loc"].data_ptr(), pin["attn"].data_ptr(), 0, out2.data_ptr(), 0, 0, 0,
                               (N, S, M, D, 3, Lq, 4))
    assert torch.equal(out, out2)
    assert cabi.lib().bm2f_msda_release_host_workspace() == 0


# ----------------------------------------------------------------------------------------------
# 6. fused entry points (softmax + location arithmetic inside the kernels)
# ----------------------------------------------------------------------------------------------
def _fused_inputs(levels, batch, seed):
    g = torch.Generator().manual_seed(seed)
    base = W.make_inputs(levels, batch, seed=seed)
    L = len(levels)
    S = base["value"].shape[1]
    ref = W.reference_points(levels, batch)                                       # (N,S,L,2)
    offsets = (W.compass_offset_bias(8, L, 4)[None, None] + torch.randn(batch, S, 8, L, 4, 2, generator=g)).contiguous()
    logits = (torch.randn(batch, S, 8, L, 4, generator=g) * 1.5).contiguous()
    return base, ref, offsets, logits


@pytest.mark.parametrize("levels,batch,staging", [(SMALL_LEVELS, 3, 1), (SMALL_LEVELS, 3, 2),
                                                   (((5, 7), (9, 13)), 2, 1), (W.WORKLOADS[1].levels, 1, 1)])
def test_fused_vs_oracle(levels, batch, staging, built):
    base, ref, offsets, logits = _fused_inputs(levels, batch, 900 + len(levels))
    dev = _dev()
    sh, st = base["shapes"].to(dev), base["start"].to(dev)
    v, r, o, lg, go = (t.to(dev).contiguous() for t in (base["value"], ref, offsets, logits, base["grad_out"]))
    N, S, M, D = v.shape
    L = len(levels)
    dims = (N, S, M, D, L, S, 4)
    out = torch.full((N, S, M * D), float("nan"), device=dev)
    gv, goff, glog = torch.full_like(v, float("nan")), torch.full_like(o, float("nan")), torch.full_like(lg, float("nan"))
    stream = torch.cuda.current_stream().cuda_stream
    tun = cabi.make_tuning(staging=staging)
    cabi.fused_forward(v.data_ptr(), sh.data_ptr(), st.data_ptr(), r.data_ptr(), o.data_ptr(), lg.data_ptr(),
                       out.data_ptr(), dims, cabi.DTYPE_F32, tun, stream)
    cabi.fused_backward(v.data_ptr(), sh.data_ptr(), st.data_ptr(), r.data_ptr(), o.data_ptr(), lg.data_ptr(),
                        go.data_ptr(), gv.data_ptr(), goff.data_ptr(), glog.data_ptr(), dims, cabi.DTYPE_F32, tun, stream)
    torch.cuda.synchronize()
    a = [t.numpy() for t in (base["value"], base["shapes"], base["start"], ref, offsets, logits)]
    ref_out = O.fused_forward(*a)
    rgv, rgo, rgl = O.fused_backward(*a, base["grad_out"].numpy())
    assert np.abs(out.cpu().numpy() - ref_out).max() <= 2e-5 * max(1.0, np.abs(ref_out).max())
    loc = ref.numpy()[:, :, None, :, None, :] + offsets.numpy() / np.stack(
        (base["shapes"].numpy()[:, 1], base["shapes"].numpy()[:, 0]), -1)[None, None, None, :, None, :]
    ok = smooth_mask(loc, base["shapes"].numpy(), eps=1e-3)
    assert rel_err(gv.cpu().numpy(), rgv) <= 1e-4
    assert rel_err(glog.cpu().numpy(), rgl) <= 1e-4
    assert rel_err(goff.cpu().numpy() * ok, rgo * ok) <= 1e-4


def test_fused_rejects_unsupported_shapes(built):
    assert cabi.lib().bm2f_msda_fused_supported(8, 32, 3, 4, cabi.DTYPE_F32) == 1
    assert cabi.lib().bm2f_msda_fused_supported(8, 32, 3, 4, cabi.DTYPE_BF16) == 1
    assert cabi.lib().bm2f_msda_fused_supported(8, 64, 3, 4, cabi.DTYPE_F32) == 0
    assert cabi.lib().bm2f_msda_fused_supported(4, 32, 3, 4, cabi.DTYPE_F32) == 0
    assert cabi.lib().bm2f_msda_fused_supported(8, 32, 5, 4, cabi.DTYPE_F32) == 0
    assert cabi.lib().bm2f_msda_fused_supported(8, 32, 3, 4, cabi.DTYPE_F64) == 0


def test_fused_bf16_forward_vs_oracle(built):
    levels = W.WORKLOADS[3].levels
    base, ref, offsets, logits = _fused_inputs(levels, 1, 931)
    dev = _dev()
    sh, st = base["shapes"].to(dev), base["start"].to(dev)
    v = base["value"].to(dev).bfloat16().contiguous()
    r, o, lg = (t.to(dev).contiguous() for t in (ref, offsets, logits))
    N, S, M, D = v.shape
    out = torch.empty(N, S, M * D, device=dev, dtype=torch.bfloat16)
    cabi.fused_forward(v.data_ptr(), sh.data_ptr(), st.data_ptr(), r.data_ptr(), o.data_ptr(), lg.data_ptr(),
                       out.data_ptr(), (N, S, M, D, 3, S, 4), cabi.DTYPE_BF16, None,
                       torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    want = O.fused_forward(v.float().cpu().numpy(), base["shapes"].numpy(), base["start"].numpy(), ref.numpy(),
                           offsets.numpy(), logits.numpy())
    assert rel_err(out.float().cpu().numpy(), want) <= 1e-2


@pytest.mark.parametrize("levels,batch", [(SMALL_LEVELS, 2), (((1, 1), (3, 2), (25, 38)), 2), (((5, 7), (9, 13)), 1),
                                          (W.WORKLOADS[1].levels, 1)])
def test_fused_analytic_reference_points_bit_identical(levels, batch, built):
    """reference_points == NULL: the kernel derives the encoder's pixel-centre reference points (valid ratios 1,
    msdeformattn.py:141-153) from the query index.  Same fp32 division as torch, so forward output and the
    deterministic gradients are bit-identical to passing the tensor torch builds; grad_value (atomic order) to 1e-6."""
    from bm2f_b200.encoder import MSDeformAttnTransformerEncoder
    base, _, offsets, logits = _fused_inputs(levels, batch, 950 + len(levels))
    dev = _dev()
    L = len(levels)
    ref = MSDeformAttnTransformerEncoder.get_reference_points(
        list(levels), torch.ones(batch, L, 2, device=dev), dev).contiguous()             # what the encoder passes
    assert torch.equal(ref.cpu(), W.reference_points(levels, batch))
    sh, st = base["shapes"].to(dev), base["start"].to(dev)
    v, o, lg, go = (t.to(dev).contiguous() for t in (base["value"], offsets, logits, base["grad_out"]))
    N, S, M, D = v.shape
    dims = (N, S, M, D, L, S, 4)
    stream = torch.cuda.current_stream().cuda_stream
    res = []
    for r in (ref.data_ptr(), 0):
        out = torch.full((N, S, M * D), float("nan"), device=dev)
        gv, goff, glog = torch.full_like(v, float("nan")), torch.full_like(o, float("nan")), torch.full_like(lg, float("nan"))
        cabi.fused_forward(v.data_ptr(), sh.data_ptr(), st.data_ptr(), r, o.data_ptr(), lg.data_ptr(), out.data_ptr(),
                           dims, cabi.DTYPE_F32, None, stream)
        cabi.fused_backward(v.data_ptr(), sh.data_ptr(), st.data_ptr(), r, o.data_ptr(), lg.data_ptr(), go.data_ptr(),
                            gv.data_ptr(), goff.data_ptr(), glog.data_ptr(), dims, cabi.DTYPE_F32, None, stream)
        torch.cuda.synchronize()
        res.append((out, gv, goff, glog))
    (out_t, gv_t, goff_t, glog_t), (out_a, gv_a, goff_a, glog_a) = res
    assert torch.equal(out_t, out_a) and torch.equal(goff_t, goff_a) and torch.equal(glog_t, glog_a)
    assert rel_err(gv_a.cpu().numpy(), gv_t.cpu().numpy()) <= 1e-6
    # NULL reference points only make sense for self-attention over the same pyramid
    with pytest.raises(cabi.MSDAError, match="num_query == spatial_size"):
        cabi.fused_forward(v.data_ptr(), sh.data_ptr(), st.data_ptr(), 0, o.data_ptr(), lg.data_ptr(), out.data_ptr(),
                           (N, S, M, D, L, S - 1, 4), cabi.DTYPE_F32, None, stream)


@pytest.mark.parametrize("levels,batch", [(SMALL_LEVELS, 3), (((1, 1), (3, 2), (25, 38)), 2), (((5, 7), (9, 13)), 1),
                                          (W.WORKLOADS[1].levels, 1), (((32, 32), (64, 64), (16, 16), (8, 8)), 2),
                                          (((5, 7), (1, 9)), 2), (((4, 4), (6, 1)), 2), (((3, 3), (1, 1)), 1)])
def test_geometry_warp_forward_bit_identical(levels, batch, built):
    """Geometry-warp forward kernels: per-point footprints / weights / corner offsets computed once by geometry warps and
    handed to the consumers through shared-memory records.  The default (lean records: unconditional gathers, weight 0
    for dropped corners, 28 consumer warps), tuning.geo = 1 (predicated records) and the consumer-lane kernel of round 1
    (tuning.geo = 2) use the same operands and FMA order, so their outputs are bit-identical — plain and fused entry
    points, strip-walked and raster-walked levels, a 1 x 1 level, L = 2, 3, 4, a third of the points outside."""
    base, _, offsets, logits = _fused_inputs(levels, batch, 970 + len(levels))
    dev = _dev()
    L = len(levels)
    sh, st = base["shapes"].to(dev), base["start"].to(dev)
    v, lc, at, o, lg = (t.to(dev).contiguous() for t in (base["value"], base["loc"], base["attn"], offsets, logits))
    N, S, M, D = v.shape
    dims = (N, S, M, D, L, S, 4)
    stream = torch.cuda.current_stream().cuda_stream
    geo = cabi.make_tuning(geo=1)
    res = {}
    for name, tun in (("default", None), ("geo", geo), ("lanes", cabi.make_tuning(geo=2)), ("lean32", cabi.make_tuning(geo=11))):
        out_p = torch.full((N, S, M * D), float("nan"), device=dev)
        out_f = torch.full((N, S, M * D), float("nan"), device=dev)
        cabi.forward(v.data_ptr(), sh.data_ptr(), st.data_ptr(), lc.data_ptr(), at.data_ptr(), out_p.data_ptr(), dims,
                     cabi.DTYPE_F32, tun, stream)
        cabi.fused_forward(v.data_ptr(), sh.data_ptr(), st.data_ptr(), 0, o.data_ptr(), lg.data_ptr(), out_f.data_ptr(),
                           dims, cabi.DTYPE_F32, tun, stream)
        torch.cuda.synchronize()
        res[name] = (out_p, out_f)
    assert not torch.isnan(res["geo"][0]).any() and not torch.isnan(res["geo"][1]).any()
    assert torch.equal(res["lanes"][0], res["geo"][0])
    assert torch.equal(res["lanes"][1], res["geo"][1])
    assert torch.equal(res["lanes"][0], res["lean32"][0])
    # the default's 16-byte records keep a left-border column's weight as 1 - hw instead of lw (2^-25 apart at most):
    # everything else is bit-identical, so the outputs agree to one rounding of one term
    assert not torch.isnan(res["default"][0]).any()
    scale0 = max(res["lanes"][0].abs().max().item(), 1.0)
    assert (res["default"][0] - res["lanes"][0]).abs().max().item() <= 2e-7 * scale0
    assert (res["default"][0] != res["lanes"][0]).float().mean().item() <= 0.05
    assert torch.equal(res["default"][1], res["lanes"][1])
    res["default"] = res["lanes"]
    if L == 3:
        # geo = 3: the same records consumed with 256-bit gathers (8 corners per warp instruction); the partial sums
        # are split differently across lanes, so equality is to rounding
        out_w = torch.full((N, S, M * D), float("nan"), device=dev)
        cabi.forward(v.data_ptr(), sh.data_ptr(), st.data_ptr(), lc.data_ptr(), at.data_ptr(), out_w.data_ptr(), dims,
                     cabi.DTYPE_F32, cabi.make_tuning(geo=3), stream)
        torch.cuda.synchronize()
        scale = res["default"][0].abs().max().item()
        assert (out_w - res["default"][0]).abs().max().item() <= 2e-6 * max(scale, 1.0)
        # geo = 1 with two CTAs per SM (47 registers per thread): bit-identical again
        out_2 = torch.full((N, S, M * D), float("nan"), device=dev)
        cabi.forward(v.data_ptr(), sh.data_ptr(), st.data_ptr(), lc.data_ptr(), at.data_ptr(), out_2.data_ptr(), dims,
                     cabi.DTYPE_F32, cabi.make_tuning(geo=1, ctas_per_sm=2), stream)
        torch.cuda.synchronize()
        assert torch.equal(out_2, res["default"][0])


def test_default_forward_skips_non_finite_locations(built):
    """The reference's range test (`h_im > -1 && w_im > -1 && h_im < H && w_im < W`, ms_deform_im2col_cuda.cuh:264) is false
    for NaN / infinite sampling locations, so such points contribute nothing.  The default forward (unconditional gathers
    with zero weights) must not turn them into 0 * NaN: same bits as with a finite out-of-range location."""
    inp = W.make_inputs(SMALL_LEVELS, 2, seed=5)
    dev = _dev()
    N, S, M, D = inp["value"].shape
    L = len(SMALL_LEVELS)
    dims = (N, S, M, D, L, S, 4)
    stream = torch.cuda.current_stream().cuda_stream
    sh, st = inp["shapes"].to(dev), inp["start"].to(dev)
    v, at = inp["value"].to(dev).contiguous(), inp["attn"].to(dev).contiguous()
    outs = []
    for bad in ((5.0, 5.0), (float("inf"), 0.3), (0.4, float("nan")), (float("-inf"), float("inf"))):
        loc = inp["loc"].clone()
        flat = loc.view(-1, 2)
        flat[3::97] = torch.tensor(bad)
        lc = loc.to(dev).contiguous()
        out = torch.full((N, S, M * D), float("nan"), device=dev)
        cabi.forward(v.data_ptr(), sh.data_ptr(), st.data_ptr(), lc.data_ptr(), at.data_ptr(), out.data_ptr(), dims,
                     cabi.DTYPE_F32, None, stream)
        torch.cuda.synchronize()
        outs.append(out)
    assert not torch.isnan(outs[0]).any()
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
