"""GPU: the packed offsets||logits path of encoder self-attention (reference
ops/modules/ms_deform_attn.py:101-102): one 256 -> 288 projection, the fused attention op reading / writing column blocks of
one matrix.  The packed op must be bit-identical to the unpacked one (same kernels, other row stride); the module must give
the same outputs and gradients with `packed_projections` on and off."""
import pytest
import torch

from bm2f_b200 import workloads as W

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


@pytest.fixture(scope="module")
def msda(built):
    import bm2f_b200
    return bm2f_b200.load_extension()


@pytest.mark.parametrize("levels,batch,with_ref", [(((5, 7), (9, 13), (16, 16)), 2, True), (W.WORKLOADS[1].levels, 1, False),
                                                   (W.WORKLOADS[2].levels, 4, False), (((6, 10), (12, 20)), 3, True)])
def test_packed_fused_op_is_bit_identical(msda, levels, batch, with_ref):
    """small launches run the per-corner backward, 4 images of cfg 2 the anchor-sorted one"""
    torch.manual_seed(len(levels) * 10 + batch)
    L, M, D, P = len(levels), 8, 32, 4
    S = sum(h * w for h, w in levels)
    shapes, start = W.level_tensors(levels, DEV)
    value = torch.randn(batch, S, M, D, device=DEV)
    offsets = (W.compass_offset_bias(M, L, P)[None, None].to(DEV) + torch.randn(batch, S, M, L, P, 2, device=DEV)).contiguous()
    logits = torch.randn(batch, S, M, L * P, device=DEV)
    ref = W.reference_points(levels, batch).to(DEV).contiguous() if with_ref else None
    go = torch.randn(batch, S, M * D, device=DEV)
    oa = torch.cat((offsets.reshape(batch, S, -1), logits.reshape(batch, S, -1)), -1).contiguous()
    assert oa.shape[-1] == M * L * P * 3
    out_u = msda.ms_deform_attn_fused_forward(value, shapes, start, ref, offsets, logits)
    out_p = msda.ms_deform_attn_fused_forward_packed(value, shapes, start, ref, oa, P)
    assert torch.equal(out_u, out_p)
    gv_u, goff_u, glog_u = msda.ms_deform_attn_fused_backward(value, shapes, start, ref, offsets, logits, go)
    poison = torch.full_like(oa, float("nan"))
    del poison
    gv_p, goa_p = msda.ms_deform_attn_fused_backward_packed(value, shapes, start, ref, oa, P, go)
    assert not torch.isnan(goa_p).any()
    n_off = M * L * P * 2
    assert torch.equal(goa_p[..., :n_off].reshape_as(goff_u), goff_u)
    assert torch.equal(goa_p[..., n_off:].reshape_as(glog_u), glog_u)
    # grad_value is a float sum of reductions in scheduling order: equal to rounding
    assert (gv_p - gv_u).abs().max().item() <= 1e-5 * gv_u.abs().max().item()


def test_module_packed_projections_match_two_projections(msda):
    from bm2f_b200.ops.modules import MSDeformAttn
    torch.manual_seed(3)
    levels = ((8, 12), (16, 24), (32, 48))
    S = sum(h * w for h, w in levels)
    mod = MSDeformAttn(256, 3, 8, 4).to(DEV)
    with torch.no_grad():
        mod.sampling_offsets.weight.normal_(0, 0.02)
        mod.attention_weights.weight.normal_(0, 0.05)
    shapes, start = W.level_tensors(levels, DEV)
    ref = W.reference_points(levels, 2).to(DEV)
    src = torch.randn(2, S, 256, device=DEV, requires_grad=True)
    pos = torch.randn(1, S, 256, device=DEV, requires_grad=True)
    mask = torch.zeros(2, S, dtype=torch.bool, device=DEV)
    mask[1, -37:] = True
    g = torch.randn(2, S, 256, device=DEV)
    res = {}
    for packed in (True, False):
        mod.packed_projections = packed
        for t in (src, pos):
            t.grad = None
        mod.zero_grad(set_to_none=True)
        assert mod.self_attention_supported(src, pos, ref)
        out = mod.forward_self_attention(src, pos, ref, shapes, start, mask)
        out.backward(g)
        res[packed] = [out.detach().clone(), src.grad.clone(), pos.grad.clone()] + [p.grad.clone() for p in mod.parameters()]
    for a, b in zip(res[True], res[False]):
        assert a.shape == b.shape
        scale = b.abs().max().item()
        assert (a - b).abs().max().item() <= 2e-5 * max(scale, 1e-6)      # one K = 288 chain vs two GEMMs summed: fp32 rounding
