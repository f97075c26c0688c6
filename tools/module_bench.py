"""Module-level timing: MSDeformAttn fwd+bwd for one encoder layer at a BASELINE config,
   (a) this repository: fused prologue + tcgen05 forward GEMMs, (b) same with torch Linear, (c) unfused op +
   torch Linear (the reference module's op sequence on our kernels), (d) the same sequence on the reference's
   own CUDA op when oracle/_ref exists.   python tools/module_bench.py [--cfg 2] [--batch 16]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from bm2f_b200 import workloads as W
from bm2f_b200.ops.modules import MSDeformAttn


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", type=int, default=2)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    wl = W.WORKLOADS[args.cfg]
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    mod = MSDeformAttn(256, wl.L, 8, 4).to(dev)
    with torch.no_grad():
        mod.sampling_offsets.weight.normal_(0, 0.01)
        mod.attention_weights.weight.normal_(0, 0.05)
    shapes, start = W.level_tensors(wl.levels, dev)
    src = torch.randn(args.batch, wl.S, 256, device=dev)
    pos = torch.randn(args.batch, wl.S, 256, device=dev) * 0.1
    ref_pts = W.reference_points(wl.levels, args.batch).to(dev)
    mask = torch.zeros(args.batch, wl.S, dtype=torch.bool, device=dev)      # Mask2Former passes an all-False mask
    gout = torch.randn(args.batch, wl.S, 256, device=dev)

    REF = None
    try:
        from oracle import build_ref
        REF = build_ref.load()
    except Exception:
        pass

    class RefFn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, v, sh, st, loc, aw, step):
            ctx.save_for_backward(v, sh, st, loc, aw)
            ctx.step = step
            return REF.ms_deform_attn_forward(v, sh, st, loc, aw, step)

        @staticmethod
        def backward(ctx, g):
            v, sh, st, loc, aw = ctx.saved_tensors
            a, b, c = REF.ms_deform_attn_backward(v, sh, st, loc, aw, g.contiguous(), ctx.step)
            return a, None, None, b, c, None

    def run(kind):
        q = (src + pos).requires_grad_(True)
        x = src.clone().requires_grad_(True)
        if kind == "reference_op":
            import torch.nn.functional as F
            value = mod.value_proj(x).masked_fill(mask[..., None], 0.0).view(args.batch, wl.S, 8, 32)
            off = mod.sampling_offsets(q).view(args.batch, wl.S, 8, wl.L, 4, 2)
            aw = F.softmax(mod.attention_weights(q).view(args.batch, wl.S, 8, wl.L * 4), -1).view(args.batch, wl.S, 8, wl.L, 4)
            norm = torch.stack([shapes[..., 1], shapes[..., 0]], -1)
            loc = ref_pts[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
            y = mod.output_proj(RefFn.apply(value, shapes, start, loc, aw, 128))
        else:
            y = mod(q, ref_pts, x, shapes, start, mask)
        y.backward(gout)
        return y

    variants = {"fused+tcgen05": (True, True), "fused+torch_linear": (True, False), "unfused+torch_linear": (False, False)}
    res = {}
    outs = {}
    for name, cfg in list(variants.items()) + ([("reference_op", None)] if REF is not None else []):
        if name != "reference_op":
            mod.fuse_prologue, mod.tcgen05_linear = cfg
        kind = "reference_op" if name == "reference_op" else "ours"
        for _ in range(2):
            outs[name] = run(kind).detach()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(args.reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); run(kind); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        res[name] = best
        print(f"{name:24s} {best:8.3f} ms / layer (batch {args.batch}, fwd+bwd)  -> {args.batch / (6 * best * 1e-3):8.1f} images/s for 6 layers", flush=True)
    base = outs["unfused+torch_linear"]
    for k, v in outs.items():
        print(f"max |out - unfused| {k:24s} {(v - base).abs().max().item():.3e} (|out| max {base.abs().max().item():.2f})")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"module_bench_cfg{args.cfg}.json"), "w"))


if __name__ == "__main__":
    main()
