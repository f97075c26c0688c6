"""GPU: encoder-layer pieces next to the hot path — fused residual+LayerNorm, FFN GEMMs
with fused ReLU, and the encoder mirror (reference: msdeformattn.py:22-161) fused vs the reference op sequence."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from bm2f_b200 import workloads as W
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def msda(built):
    import bm2f_b200
    return bm2f_b200.load_extension()


@pytest.mark.parametrize("rows", [1, 77, 4096, 21504])
def test_add_layernorm_matches_torch(msda, rows):
    from bm2f_b200.ops.functions.encoder_func import add_layernorm
    torch.manual_seed(rows)
    norm = torch.nn.LayerNorm(256).to(DEV)
    with torch.no_grad():
        norm.weight.normal_(1.0, 0.2); norm.bias.normal_(0, 0.2)
    x = torch.randn(rows, 256, device=DEV, requires_grad=True)
    r = torch.randn(rows, 256, device=DEV, requires_grad=True)
    go = torch.randn(rows, 256, device=DEV)
    y = add_layernorm(x, r, norm)
    y.backward(go)
    got = (y.detach(), x.grad.clone(), r.grad.clone(), norm.weight.grad.clone(), norm.bias.grad.clone())
    for p in (x, r, norm.weight, norm.bias):
        p.grad = None
    ref = F.layer_norm((x.double() + r.double()), (256,), norm.weight.double(), norm.bias.double(), norm.eps)
    ref.backward(go.double())
    want = (ref.detach(), x.grad, r.grad, norm.weight.grad, norm.bias.grad)
    for a, b in zip(got, want):
        assert rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 2e-5


@pytest.mark.parametrize("rows", [300, 5376])
def test_ffn_matches_fp64(msda, rows):
    from bm2f_b200.ops.functions.encoder_func import ffn
    torch.manual_seed(rows)
    l1 = torch.nn.Linear(256, 1024).to(DEV)
    l2 = torch.nn.Linear(1024, 256).to(DEV)
    x = torch.randn(rows, 256, device=DEV, requires_grad=True)
    go = torch.randn(rows, 256, device=DEV)
    y = ffn(x, l1, l2)
    y.backward(go)
    got = [t.detach().clone() for t in (y, x.grad, l1.weight.grad, l1.bias.grad, l2.weight.grad, l2.bias.grad)]
    h32 = msda.linear_relu_tf32x3(x.detach(), l1.weight.detach(), l1.bias.detach(), 3)
    # float64 reference with the SAME ReLU mask as the fp32 forward: a pre-activation within 1e-6 of zero may land on
    # either side in fp32 vs fp64, and the gradient is discontinuous there (same situation as the bilinear kinks)
    xd = x.detach().double().requires_grad_(True)
    w1, b1, w2, b2 = (t.detach().double().requires_grad_(True) for t in (l1.weight, l1.bias, l2.weight, l2.bias))
    pre = xd @ w1.t() + b1
    hd = pre * (h32 > 0).double()
    assert rel_err(h32.cpu().numpy(), hd.detach().cpu().numpy()) <= 5e-6
    yd = hd @ w2.t() + b2
    yd.backward(go.double())
    want = [yd, xd.grad, w1.grad, b1.grad, w2.grad, b2.grad]
    tol = [2e-5, 2e-5, 5e-5, 5e-5, 5e-5, 5e-5]           # y / grad_x reduce over 1024 (128 accumulation steps)
    for a, b, t in zip(got, want, tol):
        assert rel_err(a.cpu().numpy(), b.detach().cpu().numpy()) <= t


def _layer_pair(levels, seed=0):
    from bm2f_b200.encoder import MSDeformAttnTransformerEncoderLayer
    torch.manual_seed(seed)
    layer = MSDeformAttnTransformerEncoderLayer(256, 1024, 0.0, "relu", len(levels), 8, 4).to(DEV)
    with torch.no_grad():
        layer.self_attn.sampling_offsets.weight.normal_(0, 0.01)
        layer.self_attn.attention_weights.weight.normal_(0, 0.05)
    return layer


def test_encoder_layer_fused_equals_reference_sequence(msda):
    levels = ((8, 8), (16, 16), (32, 32))
    layer = _layer_pair(levels)
    shapes, start = W.level_tensors(levels, DEV)
    S = sum(h * w for h, w in levels)
    src = torch.randn(2, S, 256, device=DEV)
    pos = torch.randn(2, S, 256, device=DEV) * 0.1
    ref_pts = W.reference_points(levels, 2).to(DEV)
    mask = torch.zeros(2, S, dtype=torch.bool, device=DEV)
    go = torch.randn(2, S, 256, device=DEV)

    def run(fused):
        layer.fused = fused
        layer.self_attn.fuse_prologue = fused
        layer.self_attn.tcgen05_linear = fused
        layer.zero_grad()
        s = src.clone().requires_grad_(True)
        p = pos.clone().requires_grad_(True)
        out = layer(s, p, ref_pts, shapes, start, mask)
        out.backward(go)
        return [out.detach(), s.grad, p.grad] + [q.grad.clone() for q in layer.parameters()]

    a, b = run(True), run(False)
    assert (a[0] - b[0]).abs().max().item() <= 5e-5 * max(1.0, b[0].abs().max().item())
    for x, y in zip(a[1:], b[1:]):
        assert rel_err(x.cpu().numpy(), y.cpu().numpy()) <= 1e-3


def test_encoder_only_end_to_end(msda):
    from bm2f_b200.encoder import MSDeformAttnTransformerEncoderOnly
    torch.manual_seed(1)
    enc = MSDeformAttnTransformerEncoderOnly(256, 8, 2, 1024, 0.0, "relu", 3, 4).to(DEV)
    srcs = [torch.randn(2, 256, h, w, device=DEV) for h, w in ((4, 6), (8, 12), (16, 24))]
    poss = [torch.randn_like(s) * 0.1 for s in srcs]
    mem, shapes, start = enc(srcs, poss)
    assert mem.shape == (2, 24 + 96 + 384, 256) and shapes.tolist() == [[4, 6], [8, 12], [16, 24]]
    assert start.tolist() == [0, 24, 120] and torch.isfinite(mem).all()
    for m in enc.modules():
        if hasattr(m, "fused"):
            m.fused = False
        if hasattr(m, "fuse_prologue"):
            m.fuse_prologue = False
            m.tcgen05_linear = False
    mem2, _, _ = enc(srcs, poss)
    assert (mem - mem2).abs().max().item() <= 1e-4 * max(1.0, mem2.abs().max().item())


def test_self_attention_entry_equals_module_forward(msda):
    """`MSDeformAttn.forward_self_attention(src, pos, ...)` (one autograd node for the three input projections, gradient
    branches summed in GEMM epilogues) against `forward(src + pos, ref, src, ...)` of the same module: outputs and all
    gradients, with a shared (1, S, C) position table and with a per-image one, and with a padding mask."""
    from bm2f_b200.encoder import MSDeformAttnTransformerEncoder
    from bm2f_b200.ops.modules import MSDeformAttn
    torch.manual_seed(3)
    levels = ((6, 9), (12, 18), (24, 36))
    n, S = 2, sum(h * w for h, w in levels)
    attn = MSDeformAttn(256, 3, 8, 4).to(DEV)
    with torch.no_grad():
        attn.sampling_offsets.weight.normal_(0, 0.02); attn.attention_weights.weight.normal_(0, 0.05)
    shapes = torch.as_tensor(levels, dtype=torch.long, device=DEV)
    start = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    ref = MSDeformAttnTransformerEncoder.get_reference_points(list(levels), torch.ones(n, 3, 2, device=DEV), DEV)
    mask = torch.zeros(n, S, dtype=torch.bool, device=DEV); mask[1, -40:] = True
    go = torch.randn(n, S, 256, device=DEV)
    for pos_batch, pad in ((1, None), (n, None), (1, mask)):
        src = torch.randn(n, S, 256, device=DEV, requires_grad=True)
        pos = torch.randn(pos_batch, S, 256, device=DEV, requires_grad=True)
        assert attn.self_attention_supported(src, pos, ref)
        res = []
        for fast in (True, False):
            for t in [src, pos] + list(attn.parameters()):
                t.grad = None
            out = attn.forward_self_attention(src, pos, ref, shapes, start, pad) if fast \
                else attn(src + pos, ref, src, shapes, start, pad)
            out.backward(go)
            res.append([out.detach()] + [t.grad.clone() for t in [src, pos] + list(attn.parameters())])
        for a, b in zip(*res):
            assert a.shape == b.shape
            assert rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 2e-5


def test_module_gemms_follow_the_matmul_tf32_flag(msda):
    """`torch.backends.cuda.matmul.allow_tf32` (default False; Mask2Former leaves it off) selects one TF32 pass for the
    projections / FFN exactly where the reference's nn.Linear would run TF32; off = three-term split, fp32-grade."""
    from bm2f_b200.encoder import MSDeformAttnTransformerEncoderLayer, MSDeformAttnTransformerEncoder
    torch.manual_seed(11)
    levels = ((8, 8), (16, 16), (32, 32))
    n, S = 2, sum(h * w for h, w in levels)
    layer = MSDeformAttnTransformerEncoderLayer(256, 1024, 0.0, "relu", 3, 8, 4).to(DEV)
    shapes = torch.as_tensor(levels, dtype=torch.long, device=DEV)
    start = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    ref_pts = MSDeformAttnTransformerEncoder.get_reference_points(list(levels), torch.ones(n, 3, 2, device=DEV), DEV)
    src, pos = torch.randn(n, S, 256, device=DEV), torch.randn(1, S, 256, device=DEV) * 0.1
    old = torch.backends.cuda.matmul.allow_tf32
    try:
        with torch.no_grad():
            torch.backends.cuda.matmul.allow_tf32 = False
            exact = layer(src, pos, ref_pts, shapes, start)
            layer.fused = False
            for m in layer.modules():
                if hasattr(m, "tcgen05_linear"): m.tcgen05_linear = False; m.fuse_prologue = False
            want = layer(src, pos, ref_pts, shapes, start)          # cuBLAS fp32 + torch ops around the sampling op
            layer.fused = True
            for m in layer.modules():
                if hasattr(m, "tcgen05_linear"): m.tcgen05_linear = True; m.fuse_prologue = True
            torch.backends.cuda.matmul.allow_tf32 = True
            fast = layer(src, pos, ref_pts, shapes, start)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    e_exact = rel_err(exact.cpu().numpy(), want.cpu().numpy())
    e_fast = rel_err(fast.cpu().numpy(), want.cpu().numpy())
    assert e_exact <= 2e-5, e_exact
    assert 2e-5 < e_fast <= 1e-2, e_fast
