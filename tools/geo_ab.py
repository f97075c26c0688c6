"""A/B of the forward kernels at cfg 2: default fast kernel vs geometry-warp kernel (tuning.geo = 1), plain and fused
entry points; outputs must be bit-identical.   python tools/geo_ab.py [batch] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bm2f_b200 import cabi
from bm2f_b200 import workloads as W

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda:0"); torch.manual_seed(0)
wl = W.WORKLOADS[2]; levels = list(wl.levels); L, M, D, P = len(levels), 8, 32, 4
S = sum(h * w for h, w in levels)
shapes, start = W.level_tensors(levels); shapes, start = shapes.to(dev), start.to(dev)
value = torch.randn(batch, S, M, D, device=dev)
ref = W.reference_points(levels, batch).to(dev).contiguous()
offsets = (W.compass_offset_bias(M, L, P)[None, None].to(dev) + torch.randn(batch, S, M, L, P, 2, device=dev)).contiguous()
logits = torch.randn(batch, S, M, L * P, device=dev)
norm = torch.stack((shapes[:, 1], shapes[:, 0]), -1).float()
loc = (ref[:, :, None, :, None, :] + offsets / norm[None, None, None, :, None, :]).contiguous()
attn = torch.softmax(logits, -1).view(batch, S, M, L, P).contiguous()
outs = {k: torch.empty(batch, S, M * D, device=dev) for k in ("plain", "plain_geo", "plain_geo_2cta", "plain_geo_wide", "fused", "fused_geo", "geo_20w", "geo_24w", "geo_28w", "lean_16w", "lean_24w", "lean_28w", "lean_28w_3g", "lean_26w_5g", "lean_24w_6g", "lean_24w_4g", "lean_20w_8g", "lean_16w_4g", "rec16_28w_3g", "rec16_26w_5g", "fused_rec16_28_3", "fused_rec16_26_5", "fused_rec16_24_7")}
dims = (batch, S, M, D, L, S, P); st = torch.cuda.current_stream().cuda_stream
geo = cabi.make_tuning(geo=1); geow = cabi.make_tuning(geo=3); geo2 = cabi.make_tuning(geo=1, ctas_per_sm=2)
P_ = lambda t: t.data_ptr()
fns = {
    "plain": lambda: cabi.forward(P_(value), P_(shapes), P_(start), P_(loc), P_(attn), P_(outs["plain"]), dims, 0, None, st),
    "plain_geo": lambda: cabi.forward(P_(value), P_(shapes), P_(start), P_(loc), P_(attn), P_(outs["plain_geo"]), dims, 0, geo, st),
    "plain_geo_2cta": lambda: cabi.forward(P_(value), P_(shapes), P_(start), P_(loc), P_(attn), P_(outs["plain_geo_2cta"]), dims, 0, geo2, st),
    "plain_geo_wide": lambda: cabi.forward(P_(value), P_(shapes), P_(start), P_(loc), P_(attn), P_(outs["plain_geo_wide"]), dims, 0, geow, st),
    "geo_20w": lambda: cabi.forward(P_(value), P_(shapes), P_(start), P_(loc), P_(attn), P_(outs["geo_20w"]), dims, 0, cabi.make_tuning(geo=5), st),
    "geo_24w": lambda: cabi.forward(P_(value), P_(shapes), P_(start), P_(loc), P_(attn), P_(outs["geo_24w"]), dims, 0, cabi.make_tuning(geo=6), st),
    "geo_28w": lambda: cabi.forward(P_(value), P_(shapes), P_(start), P_(loc), P_(attn), P_(outs["geo_28w"]), dims, 0, cabi.make_tuning(geo=7), st),
    "lean_16w": lambda: cabi.forward(P_(value), P_(shapes), P_(start), P_(loc), P_(attn), P_(outs["lean_16w"]), dims, 0, cabi.make_tuning(geo=8), st),
    "lean_24w": lambda: cabi.forward(P_(value), P_(shapes), P_(start), P_(loc), P_(attn), P_(outs["lean_24w"]), dims, 0, cabi.make_tuning(geo=9), st),
    "lean_28w": lambda: cabi.forward(P_(value), P_(shapes), P_(start), P_(loc), P_(attn), P_(outs["lean_28w"]), dims, 0, cabi.make_tuning(geo=10), st),
    "lean_28w_3g": lambda: cabi.forward(P_(value), P_(shapes), P_(start), P_(loc), P_(attn), P_(outs["lean_28w_3g"]), dims, 0, cabi.make_tuning(geo=11), st),
    "lean_26w_5g": lambda: cabi.forward(P_(value), P_(shapes), P_(start), P_(loc), P_(attn), P_(outs["lean_26w_5g"]), dims, 0, cabi.make_tuning(geo=12), st),
    "lean_24w_6g": lambda: cabi.forward(P_(value), P_(shapes), P_(start), P_(loc), P_(attn), P_(outs["lean_24w_6g"]), dims, 0, cabi.make_tuning(geo=13), st),
    "lean_24w_4g": lambda: cabi.forward(P_(value), P_(shapes), P_(start), P_(loc), P_(attn), P_(outs["lean_24w_4g"]), dims, 0, cabi.make_tuning(geo=14), st),
    "lean_20w_8g": lambda: cabi.forward(P_(value), P_(shapes), P_(start), P_(loc), P_(attn), P_(outs["lean_20w_8g"]), dims, 0, cabi.make_tuning(geo=15), st),
    "lean_16w_4g": lambda: cabi.forward(P_(value), P_(shapes), P_(start), P_(loc), P_(attn), P_(outs["lean_16w_4g"]), dims, 0, cabi.make_tuning(geo=16), st),
    "rec16_28w_3g": lambda: cabi.forward(P_(value), P_(shapes), P_(start), P_(loc), P_(attn), P_(outs["rec16_28w_3g"]), dims, 0, cabi.make_tuning(geo=17), st),
    "rec16_26w_5g": lambda: cabi.forward(P_(value), P_(shapes), P_(start), P_(loc), P_(attn), P_(outs["rec16_26w_5g"]), dims, 0, cabi.make_tuning(geo=17), st),
    "fused": lambda: cabi.fused_forward(P_(value), P_(shapes), P_(start), 0, P_(offsets), P_(logits), P_(outs["fused"]), dims, 0, None, st),
    "fused_rec16_28_3": lambda: cabi.fused_forward(P_(value), P_(shapes), P_(start), 0, P_(offsets), P_(logits), P_(outs["fused_rec16_28_3"]), dims, 0, cabi.make_tuning(geo=21), st),
    "fused_rec16_26_5": lambda: cabi.fused_forward(P_(value), P_(shapes), P_(start), 0, P_(offsets), P_(logits), P_(outs["fused_rec16_26_5"]), dims, 0, cabi.make_tuning(geo=22), st),
    "fused_rec16_24_7": lambda: cabi.fused_forward(P_(value), P_(shapes), P_(start), 0, P_(offsets), P_(logits), P_(outs["fused_rec16_24_7"]), dims, 0, cabi.make_tuning(geo=23), st),
    "fused_geo": lambda: cabi.fused_forward(P_(value), P_(shapes), P_(start), 0, P_(offsets), P_(logits), P_(outs["fused_geo"]), dims, 0, geo, st),
}
def t(fn):
    fn(); fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best
for k, fn in list(fns.items()):
    try:
        print(f"{k:10s} {t(fn):7.3f} ms", flush=True)
    except Exception as e:          # sweep-only variants (BM2F_SWEEP build) fall through to the default kernel or raise
        print(f"{k:10s} not in this build: {e}", flush=True)
print("geo 20 / 24 / 28 consumer warps == plain:", [bool(torch.equal(outs["plain"], outs[k])) for k in ("geo_20w", "geo_24w", "geo_28w")])
print("lean variants == plain:", [bool(torch.equal(outs["plain"], outs[k])) for k in outs if k.startswith("lean")])
print("16-byte records: max |diff| vs plain", [float((outs[k] - outs["plain"]).abs().max()) for k in ("rec16_28w_3g", "rec16_26w_5g")],
      "fraction of entries that differ", float((outs["rec16_28w_3g"] != outs["plain"]).float().mean()))
print("fused 16-byte records: max |diff| vs fused", [float((outs[k] - outs["fused"]).abs().max()) for k in ("fused_rec16_28_3", "fused_rec16_26_5", "fused_rec16_24_7")])
print("plain_geo_2cta == plain:", bool(torch.equal(outs["plain"], outs["plain_geo_2cta"])))
print("plain_geo == plain:", bool(torch.equal(outs["plain"], outs["plain_geo"])),
      " fused_geo == fused:", bool(torch.equal(outs["fused"], outs["fused_geo"])),
      " max |wide - plain|", float((outs["plain_geo_wide"] - outs["plain"]).abs().max()),
      " max |fused - plain|", float((outs["fused"] - outs["plain"]).abs().max()))
