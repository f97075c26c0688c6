"""A/B of the fused-prologue sampling kernels against the plain ones on IDENTICAL sampling positions (cfg 2 shape):
   offsets / logits -> (a) torch softmax + loc arithmetic, then bm2f_msda_forward/backward; (b) bm2f_msda_fused_*.
   python tools/fused_ab.py [batch] [noise_px]      (noise 1.0 = bench.py's distribution, 0 = freshly initialised module)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bm2f_b200 import cabi
from bm2f_b200 import workloads as W

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
noise = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev = torch.device("cuda:0"); torch.manual_seed(0)
wl = W.WORKLOADS[2]; levels = list(wl.levels); L, M, D, P = len(levels), 8, 32, 4
S = sum(h * w for h, w in levels)
shapes, start = W.level_tensors(levels); shapes, start = shapes.to(dev), start.to(dev)
value = torch.randn(batch, S, M, D, device=dev)
ref = W.reference_points(levels, batch).to(dev).contiguous()                       # (N, S, L, 2)
offsets = (W.compass_offset_bias(M, L, P)[None, None].to(dev) + noise * torch.randn(batch, S, M, L, P, 2, device=dev)).contiguous()
logits = torch.randn(batch, S, M, L * P, device=dev)
norm = torch.stack((shapes[:, 1], shapes[:, 0]), -1).float()
loc = (ref[:, :, None, :, None, :] + offsets / norm[None, None, None, :, None, :]).contiguous()
attn = torch.softmax(logits, -1).view(batch, S, M, L, P).contiguous()
go = torch.randn(batch, S, M * D, device=dev)
out_a, out_b = torch.empty(batch, S, M * D, device=dev), torch.empty(batch, S, M * D, device=dev)
gv_a, gv_b = torch.empty_like(value), torch.empty_like(value)
gl, ga = torch.empty_like(loc), torch.empty_like(attn)
goff, glog = torch.empty_like(offsets), torch.empty_like(logits)
dims = (batch, S, M, D, L, S, P); st = torch.cuda.current_stream().cuda_stream
P_ = lambda t: t.data_ptr()

def f_plain(): cabi.forward(P_(value), P_(shapes), P_(start), P_(loc), P_(attn), P_(out_a), dims, 0, None, st)
def b_plain(): cabi.backward(P_(value), P_(shapes), P_(start), P_(loc), P_(attn), P_(go), P_(gv_a), P_(gl), P_(ga), dims, 0, None, st)
def f_fused(): cabi.fused_forward(P_(value), P_(shapes), P_(start), P_(ref), P_(offsets), P_(logits), P_(out_b), dims, 0, None, st)
def b_fused(): cabi.fused_backward(P_(value), P_(shapes), P_(start), P_(ref), P_(offsets), P_(logits), P_(go), P_(gv_b), P_(goff), P_(glog), dims, 0, None, st)

out_c = torch.empty_like(out_b); gv_c = torch.empty_like(gv_b); goff_c, glog_c = torch.empty_like(goff), torch.empty_like(glog)
def f_analytic(): cabi.fused_forward(P_(value), P_(shapes), P_(start), 0, P_(offsets), P_(logits), P_(out_c), dims, 0, None, st)
def b_analytic(): cabi.fused_backward(P_(value), P_(shapes), P_(start), 0, P_(offsets), P_(logits), P_(go), P_(gv_c), P_(goff_c), P_(glog_c), dims, 0, None, st)

def t(fn):
    fn(); fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best

print(f"batch {batch} noise {noise} px")
for name, fn in (("forward plain", f_plain), ("forward fused", f_fused), ("forward fused, ref in kernel", f_analytic),
                 ("backward plain", b_plain), ("backward fused", b_fused), ("backward fused, ref in kernel", b_analytic)):
    print(f"{name:30s} {t(fn):7.3f} ms", flush=True)
print("analytic == tensor reference points:", bool(torch.equal(out_b, out_c)), bool(torch.equal(goff, goff_c)), bool(torch.equal(glog, glog_c)))
print("max |out fused - plain|", float((out_a - out_b).abs().max()), " max |grad_value diff|", float((gv_a - gv_b).abs().max()))
