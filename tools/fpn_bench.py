"""FPN tail of the pixel decoder (msdeformattn.py:341-358) at a BASELINE config, R50 widths, fwd+bwd:
token-row sm_100a path (ops/functions/fpn_func.py) vs the reference op sequence on torch's library kernels (cuDNN
convolutions, ATen GroupNorm / interpolate).  --profile prints the kernel table of the fused path; the script also checks
that the fused path launches no library convolution / normalisation / interpolation kernel."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from bm2f_b200 import workloads as W
from bm2f_b200.pixel_decoder import MSDeformAttnPixelDecoder, ShapeSpec
from bm2f_b200.ops.functions import fpn_func

ap = argparse.ArgumentParser(); ap.add_argument("--cfg", type=int, default=2); ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--reps", type=int, default=5); ap.add_argument("--profile", action="store_true")
ap.add_argument("--tf32", type=int, default=1, help="torch.backends.cudnn.allow_tf32 (default 1 = torch's default)")
args = ap.parse_args()
torch.backends.cudnn.allow_tf32 = bool(args.tf32)
wl = W.WORKLOADS[args.cfg]; dev = torch.device("cuda:0"); torch.manual_seed(0)
chans = {"res2": 256, "res3": 512, "res4": 1024, "res5": 2048}
shapes = {k: ShapeSpec(channels=c, stride=4 * 2 ** i) for i, (k, c) in enumerate(chans.items())}
dec = MSDeformAttnPixelDecoder(shapes, transformer_dropout=0.0, transformer_nheads=8, transformer_dim_feedforward=1024,
                               transformer_enc_layers=1, conv_dim=256, mask_dim=256, norm="GN",
                               transformer_in_features=["res3", "res4", "res5"], common_stride=4).to(dev).train()
n = args.batch
ph, pw = wl.levels[-1]
H, Wd = 2 * ph, 2 * pw
x = torch.randn(n, 256, H, Wd, device=dev, requires_grad=True)
enc = torch.randn(n, wl.S, 256, device=dev, requires_grad=True)          # encoder output; finest level = last slice
off = wl.S - ph * pw
g_mf = torch.randn(n, 256, H, Wd, device=dev)
lat, outc, mfc = dec.lateral_convs[0], dec.output_convs[0], dec.mask_features
assert fpn_func.supported([x], [lat], [outc], mfc)

def fused():
    tok = fpn_func.fpn_level(x, enc[:, off:], ph, pw, lat, outc)
    return fpn_func.mask_features_tokens(tok, mfc)

def reference():
    prev = enc[:, off:].transpose(1, 2).reshape(n, 256, ph, pw)
    cur = lat(x)
    y = cur + F.interpolate(prev, size=cur.shape[-2:], mode="bilinear", align_corners=False)
    return mfc(outc(y))

def step(fn):
    for t in (x, enc): t.grad = None
    dec.zero_grad(set_to_none=True)
    fn().backward(g_mf)

def timeit(fn):
    step(fn); step(fn); torch.cuda.synchronize()
    best = 1e9
    for _ in range(args.reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step(fn); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best

# agreement of the two paths (same weights, same inputs)
step(fused); gf = [x.grad.clone(), enc.grad.clone(), outc.weight.grad.clone(), lat.weight.grad.clone()]
with torch.no_grad(): of = fused()
step(reference); gr = [x.grad.clone(), enc.grad.clone(), outc.weight.grad.clone(), lat.weight.grad.clone()]
with torch.no_grad(): orf = reference()
rel = lambda a, b: ((a - b).norm() / b.norm()).item()       # L2: single entries flip ReLU branches under TF32 noise
print(f"cfg {args.cfg} batch {n}: res2 {H}x{Wd}, cudnn.allow_tf32={bool(args.tf32)}")
print("fused vs reference sequence: mask_features %.2e  grad res2 %.2e  grad enc %.2e  grad W3x3 %.2e  grad Wlat %.2e"
      % (rel(of, orf), *[rel(a, b) for a, b in zip(gf, gr)]))
tf, tr = timeit(fused), timeit(reference)
print(f"FPN tail fwd+bwd: fused sm_100a {tf:.2f} ms | reference sequence (cuDNN / ATen) {tr:.2f} ms | x{tr / tf:.2f}")

# the 3x3 convolution alone: two row tiles per weight k-block (default) vs one
import bm2f_b200
MSDA = bm2f_b200.load_extension()
xh = torch.randn(n, H + 2, Wd + 2, 256, device=dev)
def t_kernel(fn):
    fn(); fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(args.reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best
flop = 2.0 * n * H * Wd * 256 * 256 * 9
for sp in (1, 3):
    t2 = t_kernel(lambda: MSDA.conv3x3_tokens_forward(xh, outc.weight.detach(), sp))
    MSDA.conv3x3_set_variant(0)
    t1 = t_kernel(lambda: MSDA.conv3x3_tokens_forward(xh, outc.weight.detach(), sp))
    MSDA.conv3x3_set_variant(1)
    print(f"conv3x3 forward split {sp}: default {t2:.3f} ms ({flop / t2 * 1e-9:.0f} TFLOP/s) | one row tile per weight block {t1:.3f} ms")
tw = t_kernel(lambda: MSDA.conv3x3_tokens_backward_weight(xh, xh, 1))
print(f"conv3x3 grad_W split 1 (TMA, MN-major): {tw:.3f} ms ({flop / tw * 1e-9:.0f} TFLOP/s)")
xn = torch.randn(n, 256, H, Wd, device=dev)
tc = t_kernel(lambda: F.conv2d(xn, outc.weight.detach(), padding=1))
xcl = xn.contiguous(memory_format=torch.channels_last)
tcl = t_kernel(lambda: F.conv2d(xcl, outc.weight.detach(), padding=1))
print(f"cuDNN conv2d forward (allow_tf32={bool(args.tf32)}): NCHW {tc:.3f} ms | channels_last {tcl:.3f} ms")

from torch.profiler import profile, ProfilerActivity
for name, fn in (("fused", fused), ("reference", reference)):
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step(fn); torch.cuda.synchronize()
    ev = prof.key_averages()
    lib = [e.key for e in ev if not e.key.startswith(("bm2f::", "void bm2f::")) and
           any(s in e.key.lower() for s in ("cudnn", "cutlass", "implicit", "upsample", "group_norm", "xmma", "sm90", "sm100_", "nchw"))]
    print(f"{name}: {len(ev)} distinct kernels, library conv / norm / interpolate kernels: {lib if lib else 'none'}")
    if args.profile or name == "fused":
        print(ev.table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
