"""Encoder-level timing: the 6-layer deformable encoder fwd+bwd at a BASELINE config,
fused sm_100a path vs the reference op sequence in torch on our sampling kernels vs on the reference's CUDA op."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bm2f_b200 import workloads as W
from bm2f_b200.encoder import MSDeformAttnTransformerEncoderOnly

ap = argparse.ArgumentParser(); ap.add_argument("--cfg", type=int, default=2); ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--tf32", action="store_true", help="torch.backends.cuda.matmul.allow_tf32 = True for both arms")
args = ap.parse_args()
torch.backends.cuda.matmul.allow_tf32 = bool(args.tf32)
wl = W.WORKLOADS[args.cfg]; dev = torch.device("cuda:0"); torch.manual_seed(0)
enc = MSDeformAttnTransformerEncoderOnly(256, 8, 6, 1024, 0.0, "relu", wl.L, 4).to(dev)
srcs = [torch.randn(args.batch, 256, h, w, device=dev, requires_grad=True) for h, w in wl.levels]
poss = [torch.randn(args.batch, 256, h, w, device=dev) * 0.1 for h, w in wl.levels]
gout = torch.randn(args.batch, wl.S, 256, device=dev)

def set_mode(fused):
    for m in enc.modules():
        if hasattr(m, "fused"): m.fused = fused
        if hasattr(m, "fuse_prologue"): m.fuse_prologue = fused; m.tcgen05_linear = fused

def run():
    mem, _, _ = enc(srcs, poss); mem.backward(gout); return mem

res = {}
for name, fused in (("fused_sm100a", True), ("fused_sm100a, separate projection nodes", True), ("fused_sm100a", True),
                    ("reference_sequence_torch", False)):
    set_mode(fused)
    for m in enc.modules():
        if hasattr(m, "fuse_projections"): m.fuse_projections = "separate" not in name
    run(); run(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(args.reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = run(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    res[name] = best
    print(f"{name:42s} {best:9.2f} ms / encoder fwd+bwd (6 layers, batch {args.batch}) -> {args.batch / (best * 1e-3):8.1f} images/s, "
          f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"encoder_bench_cfg{args.cfg}.json"), "w"))

# alternating A/B of the projection variants (the lines above run one after the other on a warming, power-capped GPU)
set_mode(True)
variants = (("one node, packed 256 -> 288 projection", True, True), ("one node, two projections", True, False),
            ("separate projection nodes", False, False))
acc = {n: [] for n, _, _ in variants}
for rnd in range(4):
    for n, fuse, packed in variants:
        for m in enc.modules():
            if hasattr(m, "fuse_projections"): m.fuse_projections = fuse
            if hasattr(m, "packed_projections"): m.packed_projections = packed
        run(); torch.cuda.synchronize()
        best = 1e9
        for _ in range(2):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); run(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
        acc[n].append(best)
for n, v in acc.items():
    print(f"A/B {n:42s} " + "  ".join(f"{x:7.2f}" for x in v) + f"   median {sorted(v)[len(v) // 2]:7.2f} ms", flush=True)
