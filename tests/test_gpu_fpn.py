"""GPU: kernels of the FPN tail (reference msdeformattn.py:341-358) against torch in float64:
3x3 convolution on tcgen05 over zero-haloed token images (forward, input gradient, weight gradient; tf32x3 and single
TF32 pass), GroupNorm + bilinear-upsample merge and its adjoint, GroupNorm + ReLU and its backward."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

@pytest.fixture(scope="module")
def msda(built):
    import bm2f_b200
    return bm2f_b200.load_extension()


SHAPES = [(2, 5, 7), (1, 16, 16), (2, 33, 40), (3, 1, 1), (1, 12, 130)]


def _halo(x_nhwc):
    return F.pad(x_nhwc, (0, 0, 1, 1, 1, 1)).contiguous()


def _rel(a, ref):
    return ((a.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()


@pytest.mark.parametrize("batch,h,w", SHAPES)
@pytest.mark.parametrize("split,tol", [(3, 5e-5), (1, 3e-3)])     # K = 2304: the fp32 accumulation chain in TMEM is 9x a projection's
def test_conv3x3_forward_and_gradients(msda, batch, h, w, split, tol):
    dev = torch.device("cuda:0")
    torch.manual_seed(batch * 1000 + h * 10 + w)
    x = torch.randn(batch, h, w, 256, device=dev)
    wgt = torch.randn(256, 256, 3, 3, device=dev) / 48.0
    g = torch.randn(batch, h, w, 256, device=dev)
    xd = x.double().permute(0, 3, 1, 2).requires_grad_(True)
    wd = wgt.double().requires_grad_(True)
    ref = F.conv2d(xd, wd, padding=1)
    ref.backward(g.double().permute(0, 3, 1, 2))
    y = msda.conv3x3_tokens_forward(_halo(x), wgt, split)
    assert y.shape == (batch, h, w, 256)
    assert _rel(y.permute(0, 3, 1, 2), ref.detach()) < tol
    gx = msda.conv3x3_tokens_backward_input(_halo(g), wgt, split)
    assert _rel(gx.permute(0, 3, 1, 2), xd.grad) < tol
    gw = msda.conv3x3_tokens_backward_weight(_halo(g), _halo(x), split)
    assert gw.shape == (256, 256, 3, 3)
    assert _rel(gw, wd.grad) < tol


def test_conv3x3_halo_is_what_pads(msda):
    """The kernel never predicates on image borders: a non-zero halo must change border outputs (i.e. the halo really is
    the padding) and must not change interior ones."""
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    x = torch.randn(1, 6, 9, 256, device=dev)
    wgt = torch.randn(256, 256, 3, 3, device=dev) / 48.0
    xh = _halo(x)
    y0 = msda.conv3x3_tokens_forward(xh, wgt, 3)
    xh2 = xh.clone()
    xh2[:, 0] = 1.0
    y1 = msda.conv3x3_tokens_forward(xh2, wgt, 3)
    assert torch.equal(y0[:, 1:], y1[:, 1:])
    assert not torch.equal(y0[:, 0], y1[:, 0])


@pytest.mark.parametrize("batch,h,w,eh,ew", [(2, 8, 12, 4, 6), (1, 32, 32, 16, 16), (2, 9, 7, 5, 3), (1, 6, 6, 6, 6),
                                              (2, 10, 10, 3, 4)])
def test_fpn_merge_forward_and_upsample_adjoint(msda, batch, h, w, eh, ew):
    dev = torch.device("cuda:0")
    torch.manual_seed(h * 100 + w)
    lat = torch.randn(batch, h, w, 256, device=dev) * 2 + 0.3
    gamma = torch.randn(256, device=dev)
    beta = torch.randn(256, device=dev)
    pre, post = 5, 3                               # the level sits inside a longer (batch, S, 256) encoder output
    enc_all = torch.randn(batch, pre + eh * ew + post, 256, device=dev)
    mean, rstd = msda.groupnorm_tokens_stats(lat, 1e-5)
    y = msda.fpn_merge_forward(lat, mean, rstd, gamma, beta, enc_all[:, pre:pre + eh * ew], eh, ew)
    assert y.shape == (batch, h + 2, w + 2, 256)
    assert y[:, 0].abs().max() == 0 and y[:, -1].abs().max() == 0 and y[:, :, 0].abs().max() == 0 and y[:, :, -1].abs().max() == 0
    enc = enc_all[:, pre:pre + eh * ew].reshape(batch, eh, ew, 256)
    encd = enc.double().permute(0, 3, 1, 2).requires_grad_(True)
    ref = F.group_norm(lat.double().permute(0, 3, 1, 2), 32, gamma.double(), beta.double(), 1e-5) + \
        F.interpolate(encd, size=(h, w), mode="bilinear", align_corners=False)
    assert _rel(y[:, 1:-1, 1:-1].permute(0, 3, 1, 2), ref.detach()) < 1e-5
    # fp32 torch: same operation order in the upsample
    up32 = F.interpolate(enc.permute(0, 3, 1, 2), size=(h, w), mode="bilinear", align_corners=False)
    gn32 = F.group_norm(lat.permute(0, 3, 1, 2), 32, gamma, beta, 1e-5)
    assert (y[:, 1:-1, 1:-1].permute(0, 3, 1, 2) - (gn32 + up32)).abs().max().item() < 2e-5
    g = torch.randn(batch, h, w, 256, device=dev)
    ref.backward(g.double().permute(0, 3, 1, 2))
    ge = msda.fpn_upsample_backward(g, eh, ew)
    assert ge.shape == (batch, eh * ew, 256)
    assert _rel(ge.reshape(batch, eh, ew, 256).permute(0, 3, 1, 2), encd.grad) < 1e-5


@pytest.mark.parametrize("batch,h,w", [(2, 5, 7), (1, 16, 16), (3, 9, 20)])
def test_groupnorm_relu_tokens(msda, batch, h, w):
    dev = torch.device("cuda:0")
    torch.manual_seed(h + w)
    y = torch.randn(batch, h, w, 256, device=dev) * 1.7 - 0.2
    gamma = torch.randn(256, device=dev)
    beta = torch.randn(256, device=dev) * 0.5
    g = torch.randn(batch, h, w, 256, device=dev)
    yd = y.double().permute(0, 3, 1, 2).requires_grad_(True)
    gd, bd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    ref = F.relu(F.group_norm(yd, 32, gd, bd, 1e-5))
    ref.backward(g.double().permute(0, 3, 1, 2))
    mean, rstd = msda.groupnorm_tokens_stats(y, 1e-5)
    out = msda.groupnorm_relu_tokens_apply(y, mean, rstd, gamma, beta)
    assert _rel(out.permute(0, 3, 1, 2), ref.detach()) < 1e-5
    gh, dgamma, dbeta = msda.groupnorm_relu_tokens_backward(g, y, mean, rstd, gamma, beta)
    assert gh.shape == (batch, h + 2, w + 2, 256)
    assert gh[:, 0].abs().max() == 0 and gh[:, :, -1].abs().max() == 0
    # entries whose pre-activation is within rounding of zero may take the other ReLU branch in fp32
    pre = F.group_norm(yd.detach(), 32, gd.detach(), bd.detach(), 1e-5)
    ok = (pre.abs() > 1e-4)
    frac_ok = ok.double().mean().item()
    assert frac_ok > 0.999
    assert _rel(gh[:, 1:-1, 1:-1].permute(0, 3, 1, 2), yd.grad) < 1e-3     # a flipped branch changes the group sums slightly
    assert _rel(dgamma, gd.grad) < 1e-3 and _rel(dbeta, bd.grad) < 1e-3


@pytest.mark.parametrize("batch,h,w", [(2, 5, 7), (2, 33, 40), (1, 12, 130)])
def test_conv3x3_tile_pair_variant_is_bit_identical(msda, batch, h, w):
    """The default single-pass kernel shares each weight k-block between two row tiles (RT = 2); per tile the MMA order is
    the one of the one-tile kernel, so the results are identical — also when the last pair has only one tile."""
    dev = torch.device("cuda:0")
    torch.manual_seed(w)
    xh = _halo(torch.randn(batch, h, w, 256, device=dev))
    wgt = torch.randn(256, 256, 3, 3, device=dev) / 48.0
    y2 = msda.conv3x3_tokens_forward(xh, wgt, 1)
    msda.conv3x3_set_variant(0)
    try:
        y1 = msda.conv3x3_tokens_forward(xh, wgt, 1)
    finally:
        msda.conv3x3_set_variant(1)
    assert torch.equal(y1, y2)
