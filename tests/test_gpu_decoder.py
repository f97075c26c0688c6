"""GPU: pixel-decoder glue — token transpose, token GroupNorm, sine position embedding, the fused
`input_proj` function — and the whole `MSDeformAttnPixelDecoder` mirror against golden vectors produced by the
UNMODIFIED reference classes on the CPU (oracle/gen_golden_decoder.py -> tests/golden/decoder/pixel_decoder_tiny.npz;
reference: msdeformattn.py:165-358, ops/modules/ms_deform_attn.py, position_encoding.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLD = os.path.join(ROOT, "tests", "golden", "decoder", "pixel_decoder_tiny.npz")


@pytest.fixture(scope="module")
def msda(built):
    import bm2f_b200
    return bm2f_b200.load_extension()


@pytest.fixture
def fp32_convs():
    """torch's convolutions default to TF32 on this GPU (`torch.backends.cudnn.allow_tf32 = True`), for the reference as
    for this repo: the FPN tail runs on the library convolutions and the 1x1 `input_proj` GEMMs follow the same flag
    (one TF32 pass instead of three).  The golden vectors / float64 references need float32-grade arithmetic."""
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old


@pytest.mark.parametrize("shape", [(1, 1, 1), (2, 320, 6), (3, 33, 65), (2, 512, 4096), (1, 7, 1000)])
def test_transpose_batched(msda, shape):
    torch.manual_seed(0)
    x = torch.randn(*shape, device=DEV)
    assert torch.equal(msda.transpose_batched(x), x.transpose(1, 2).contiguous())


@pytest.mark.parametrize("h,w", [(2, 3), (7, 5), (32, 32), (25, 38)])
def test_sine_position_embedding(msda, h, w):
    import msda_oracle
    like = torch.empty(1, device=DEV)
    got = msda.sine_position_embedding(like, h, w, 128, 10000.0, 2 * np.pi, True)       # (h*w, 256)
    got = got.t().reshape(256, h, w).cpu().numpy()
    want = msda_oracle.sine_position_embedding(h, w)
    assert np.abs(got - want).max() <= 1e-5            # north_star forward tolerance (abs, fp32)
    z = np.load(GOLD)
    if f"pos_{h}x{w}" in z.files:                      # the reference class's own output
        assert np.abs(got - z[f"pos_{h}x{w}"]).max() <= 1e-5


@pytest.mark.parametrize("batch,tokens,offset,total", [(1, 1, 0, 1), (2, 6, 0, 126), (2, 96, 30, 126), (3, 1000, 17, 1100),
                                                       (2, 16384, 5120, 21504)])
def test_groupnorm_tokens_vs_torch_fp64(msda, batch, tokens, offset, total):
    torch.manual_seed(tokens)
    y = (torch.randn(batch, tokens, 256, device=DEV) * 1.7 + 0.8)
    gamma = torch.randn(256, device=DEV) * 0.3 + 1
    beta = torch.randn(256, device=DEV) * 0.3
    out = torch.full((batch, total, 256), 7.0, device=DEV)
    mean, rstd = msda.groupnorm_tokens_forward(y, gamma, beta, 1e-5, out, offset)
    yd = y.double().transpose(1, 2).reshape(batch, 256, tokens, 1).requires_grad_(True)
    gd, bd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    ref = F.group_norm(yd, 32, gd, bd, 1e-5)
    want = ref.reshape(batch, 256, tokens).transpose(1, 2)
    assert (out[:, offset:offset + tokens] - want).abs().max().item() <= 1e-5
    # rows outside the level's slice are untouched
    assert torch.all(out[:, :offset] == 7.0) and torch.all(out[:, offset + tokens:] == 7.0)
    go = torch.randn(batch, total, 256, device=DEV)
    dy, dgamma, dbeta = msda.groupnorm_tokens_backward(go, offset, y, mean, rstd, gamma)
    gslice = go[:, offset:offset + tokens].double().transpose(1, 2).reshape(batch, 256, tokens, 1)
    ref.backward(gslice)
    want_dy = yd.grad.reshape(batch, 256, tokens).transpose(1, 2)
    assert rel_err(dy.cpu().numpy(), want_dy.cpu().numpy()) <= 1e-4 if tokens > 1 else True
    assert rel_err(dgamma.cpu().numpy(), gd.grad.cpu().numpy()) <= 1e-4
    assert rel_err(dbeta.cpu().numpy(), bd.grad.cpu().numpy()) <= 1e-4


def _input_proj(chans, seed):
    torch.manual_seed(seed)
    proj = torch.nn.ModuleList([torch.nn.Sequential(torch.nn.Conv2d(c, 256, 1), torch.nn.GroupNorm(32, 256)) for c in chans])
    with torch.no_grad():
        for p in proj:
            p[0].bias.normal_(0, 0.2); p[1].weight.normal_(1, 0.2); p[1].bias.normal_(0, 0.2)
    return proj.to(DEV)


@pytest.mark.parametrize("channels_last", [False, True])
def test_input_proj_flatten_vs_torch_fp64(msda, fp32_convs, channels_last):
    """conv1x1 + GroupNorm + flatten + cat in one function vs the reference op sequence (msdeformattn.py:66-82, 321) in
    fp64; 512 / 256 channels run on the tcgen05 GEMMs, 320 on the library GEMM, all on the token GroupNorm kernels."""
    from bm2f_b200.ops.functions import glue_func
    chans, sizes, n = [320, 512, 256], [(3, 5), (6, 10), (12, 20)], 2
    proj = _input_proj(chans, 1)
    xs = [torch.randn(n, c, h, w, device=DEV) for c, (h, w) in zip(chans, sizes)]
    if channels_last:
        xs = [x.contiguous(memory_format=torch.channels_last) for x in xs]
    xs = [x.requires_grad_(True) for x in xs]
    assert glue_func.supported(xs, proj)
    out = glue_func.input_proj_flatten(xs, proj)
    go = torch.randn_like(out)
    out.backward(go)
    got = [out.detach()] + [x.grad.clone() for x in xs] + [p.grad.clone() for p in proj.parameters()]
    for t in list(xs) + list(proj.parameters()):
        t.grad = None
    proj64 = _input_proj(chans, 1).double()
    xs64 = [x.detach().double().requires_grad_(True) for x in xs]
    ref = torch.cat([proj64[i](x).flatten(2).transpose(1, 2) for i, x in enumerate(xs64)], 1)
    ref.backward(go.double())
    want = [ref.detach()] + [x.grad for x in xs64] + [p.grad for p in proj64.parameters()]
    assert rel_err(got[0].cpu().numpy(), want[0].cpu().numpy()) <= 1e-5      # |values| reach ~4: relative to the output scale
    for a, b in zip(got[1:], want[1:]):
        assert a.shape == b.shape
        assert rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 1e-4


def _decoder(fused):
    import gen_golden_decoder as G
    from bm2f_b200.pixel_decoder import MSDeformAttnPixelDecoder, ShapeSpec
    shapes = {k: ShapeSpec(channels=c, stride=s) for k, (c, s) in G.CASE["input_shape"].items()}
    dec = G.fill_state_dict(MSDeformAttnPixelDecoder(shapes, **G.CASE["kwargs"]), G.CASE["seed"])
    dec = G.load_conditioned(dec, np.load(GOLD)).to(DEV).train()      # biases moved away from ReLU / bilinear kinks
    dec.fused = fused
    for m in dec.modules():
        if hasattr(m, "fused") and m is not dec:
            m.fused = fused
    return dec, G


def _run_decoder(dec, G):
    feats = {k: v.requires_grad_(True) for k, v in G.make_features(G.CASE, DEV).items()}
    mask_features, out0, multi = dec.forward_features(feats)
    outputs = [mask_features, out0] + list(multi)
    probes = G.make_probes(G.CASE, outputs, DEV)
    sum((p * o).sum() for p, o in zip(probes, outputs)).backward()
    res = {"mask_features": mask_features, "out0": out0}
    res.update({f"multi_scale_{i}": m for i, m in enumerate(multi)})
    res.update({f"grad_feature_{k}": v.grad for k, v in feats.items()})
    params = dict(dec.named_parameters())
    res.update({"grad_param::" + k: params[k].grad for k in G.GRAD_KEYS})
    return {k: v.detach().cpu().numpy() for k, v in res.items()}


@pytest.mark.parametrize("fused", [True, False])
def test_pixel_decoder_matches_reference_golden(msda, fp32_convs, fused):
    """Whole decoder (input_proj, position embedding, 2 encoder layers, FPN tail), forward and backward, against the
    reference's own classes run on the CPU in float32.  fused=False runs the reference op sequence around the sm_100a
    attention op, fused=True the glue / encoder kernels.  The fixture's biases keep every ReLU input and sampling
    coordinate away from its kink (oracle/gen_golden_decoder.py:condition), so the gradients are comparable entry by
    entry.  Tolerances: forward 5e-5 of the output scale (two fp32 implementations of the whole network; measured
    <= 9e-6), gradients 1e-4 relative (north_star; measured <= 2.5e-5)."""
    dec, G = _decoder(fused)
    launches = msda.launch_count()
    got = _run_decoder(dec, G)
    assert msda.launch_count() > launches
    z = np.load(GOLD)
    worst = {}
    for k in got:
        assert got[k].shape == z[k].shape, k
        worst[k] = rel_err(got[k], z[k])
    fwd = {k: v for k, v in worst.items() if not k.startswith("grad")}
    bwd = {k: v for k, v in worst.items() if k.startswith("grad")}
    assert max(fwd.values()) <= 5e-5, fwd
    assert max(bwd.values()) <= 1e-4, bwd


def test_pixel_decoder_fused_equals_reference_sequence_channels_last(msda, fp32_convs):
    """Same weights, channels_last backbone tensors: fused path (no transpose at all) vs the reference sequence."""
    dec, G = _decoder(True)
    feats = {k: v.contiguous(memory_format=torch.channels_last) for k, v in G.make_features(G.CASE, DEV).items()}
    with torch.no_grad():
        a = dec.forward_features(feats)
        dec.fused = False
        b = dec.forward_features(feats)
    for x, y in zip([a[0], a[1]] + list(a[2]), [b[0], b[1]] + list(b[2])):
        assert rel_err(x.cpu().numpy(), y.cpu().numpy()) <= 5e-5


def test_input_proj_follows_the_conv_tf32_flag(msda):
    """`torch.backends.cudnn.allow_tf32` (default True) selects one TF32 pass for the 1x1 input_proj GEMMs, as it does
    for the reference's nn.Conv2d; switched off, the GEMMs use the three-term split (fp32-grade)."""
    from bm2f_b200.ops.functions import glue_func
    chans, sizes, n = [512, 256], [(6, 10), (12, 20)], 2
    proj = _input_proj(chans, 2)
    xs = [torch.randn(n, c, h, w, device=DEV) for c, (h, w) in zip(chans, sizes)]
    proj64 = _input_proj(chans, 2).double()
    ref = torch.cat([proj64[i](x.double()).flatten(2).transpose(1, 2) for i, x in enumerate(xs)], 1)
    old = torch.backends.cudnn.allow_tf32
    try:
        errs = {}
        for flag in (True, False):
            torch.backends.cudnn.allow_tf32 = flag
            with torch.no_grad():
                errs[flag] = rel_err(glue_func.input_proj_flatten(xs, proj).cpu().numpy(), ref.cpu().numpy())
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert errs[False] <= 1e-5
    assert 2e-5 <= errs[True] <= 5e-3, errs


def test_fused_fpn_tail_runs_on_own_kernels(msda):
    """with the fused path on, forward_features + backward launch no cuDNN convolution and no ATen
    GroupNorm / bilinear-interpolation kernel (the FPN tail runs on csrc/fpn_kernels.cuh and the tcgen05 GEMMs), and the
    results agree with the reference op sequence on the same weights (default flags: single TF32 pass on both sides)."""
    from torch.profiler import profile, ProfilerActivity
    from bm2f_b200.pixel_decoder import MSDeformAttnPixelDecoder, ShapeSpec
    torch.manual_seed(5)
    chans = {"res2": 256, "res3": 512, "res4": 1024, "res5": 2048}
    shapes = {k: ShapeSpec(channels=c, stride=4 * 2 ** i) for i, (k, c) in enumerate(chans.items())}
    dec = MSDeformAttnPixelDecoder(shapes, transformer_dropout=0.0, transformer_nheads=8, transformer_dim_feedforward=1024,
                                   transformer_enc_layers=1, conv_dim=256, mask_dim=256, norm="GN",
                                   transformer_in_features=["res3", "res4", "res5"], common_stride=4).to(DEV).train()
    sizes = {"res5": (3, 4), "res4": (6, 8), "res3": (12, 16), "res2": (24, 32)}
    feats = {k: torch.randn(2, chans[k], *sizes[k], device=DEV, requires_grad=True) for k in chans}

    def run():
        for v in feats.values():
            v.grad = None
        dec.zero_grad(set_to_none=True)
        mf, out0, multi = dec.forward_features(feats)
        (mf.square().mean() + sum(m.square().mean() for m in multi)).backward()
        return mf.detach().clone(), feats["res2"].grad.clone(), dec.layer_1.weight.grad.clone(), dec.adapter_1.weight.grad.clone()

    run()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        got = run()
        torch.cuda.synchronize()
    names = [e.key for e in prof.key_averages()]
    bad = [k for k in names if not k.startswith(("bm2f::", "void bm2f::")) and
           any(s in k.lower() for s in ("cudnn", "implicit_convolve", "upsample", "group_norm", "xmma", "conv2d", "cutlass"))]
    assert not bad, bad
    assert any("linear_dw_tma_kernel" in k for k in names) and any("fpn_merge_forward_kernel" in k for k in names)
    dec.fused = False                      # FPN tail and glue on torch's library kernels, same weights
    want = run()
    # both sides single-pass TF32 with different summation orders; a ReLU input within that noise of zero takes the other
    # branch, so single entries of the gradients may differ by O(1): compare in the L2 norm
    for a, b, tol in zip(got, want, (5e-3, 2e-2, 2e-2, 2e-2)):
        assert ((a - b).norm() / b.norm()).item() <= tol


def test_fused_fpn_tail_two_levels_and_unsupported_configuration(msda, fp32_convs):
    """Two top-down FPN levels (transformer on res4 / res5 only): the second level upsamples the first level's token rows.
    Compared with the reference op sequence on the same weights in fp32-grade arithmetic, forward and gradients.  A
    decoder without GroupNorm (norm = "", conv biases) is outside the fused tail and must run the reference sequence."""
    from bm2f_b200.pixel_decoder import MSDeformAttnPixelDecoder, ShapeSpec
    from bm2f_b200.ops.functions import fpn_func
    torch.manual_seed(11)
    chans = {"res2": 256, "res3": 512, "res4": 1024, "res5": 2048}
    shapes = {k: ShapeSpec(channels=c, stride=4 * 2 ** i) for i, (k, c) in enumerate(chans.items())}
    sizes = {"res5": (3, 4), "res4": (6, 8), "res3": (12, 16), "res2": (24, 32)}

    def build(norm):
        torch.manual_seed(12)
        return MSDeformAttnPixelDecoder(shapes, transformer_dropout=0.0, transformer_nheads=8, transformer_dim_feedforward=1024,
                                        transformer_enc_layers=1, conv_dim=256, mask_dim=256, norm=norm,
                                        transformer_in_features=["res4", "res5"], common_stride=4).to(DEV).train()

    dec = build("GN")
    assert dec.num_fpn_levels == 2
    feats = {k: torch.randn(2, chans[k], *sizes[k], device=DEV, requires_grad=True) for k in chans}
    assert fpn_func.supported([feats["res3"], feats["res2"]], dec.lateral_convs, dec.output_convs, dec.mask_features)

    def run(d):
        for v in feats.values():
            v.grad = None
        d.zero_grad(set_to_none=True)
        mf, out0, multi = d.forward_features(feats)
        assert len(multi) == 3 and multi[2].shape[-2:] == sizes["res3"]          # the first FPN level is the third scale
        (mf.square().mean() + sum(m.square().mean() for m in multi)).backward()
        return [mf.detach().clone(), multi[2].detach().clone(), feats["res2"].grad.clone(), feats["res3"].grad.clone(),
                feats["res5"].grad.clone(), d.layer_1.weight.grad.clone(), d.layer_2.weight.grad.clone(),
                d.adapter_2.weight.grad.clone(), d.mask_features.bias.grad.clone()]

    got = run(dec)
    dec.fused = False
    want = run(dec)
    for a, b in zip(got, want):
        assert a.shape == b.shape
        assert ((a - b).norm() / b.norm()).item() <= 2e-4          # tf32x3 at K = 2304 on one side, cuDNN fp32 on the other
    plain = build("")
    assert not fpn_func.supported([feats["res3"], feats["res2"]], plain.lateral_convs, plain.output_convs, plain.mask_features)
    mf, _, _ = plain.forward_features(feats)
    assert mf.shape == (2, 256, 24, 32)
