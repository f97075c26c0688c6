"""Pixel-decoder timing at a BASELINE config with R50 backbone widths (res2..res5 = 256, 512,
1024, 2048 channels at strides 4..32): (1) the encoder-input glue alone — input_proj (1x1 conv + GroupNorm), position
embedding + level_embed, flatten / cat — fused sm_100a path vs the reference op sequence in torch; (2) the whole
`forward_features` fwd+bwd, fused vs reference sequence (both on our sampling kernels)."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bm2f_b200 import workloads as W
from bm2f_b200.pixel_decoder import MSDeformAttnPixelDecoder, ShapeSpec
from bm2f_b200.ops.functions import glue_func

ap = argparse.ArgumentParser(); ap.add_argument("--cfg", type=int, default=2); ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--reps", type=int, default=3); ap.add_argument("--layers", type=int, default=6)
ap.add_argument("--channels-last", action="store_true"); ap.add_argument("--split", type=int, default=None, help="TF32 terms of the input_proj GEMMs; default: follow torch.backends.cudnn.allow_tf32")
ap.add_argument("--profile", action="store_true"); args = ap.parse_args()
wl = W.WORKLOADS[args.cfg]; dev = torch.device("cuda:0"); torch.manual_seed(0)
levels = list(wl.levels)                       # lowest resolution first: res5, res4, res3
chans = {"res2": 256, "res3": 512, "res4": 1024, "res5": 2048}
shapes = {k: ShapeSpec(channels=c, stride=4 * 2 ** i) for i, (k, c) in enumerate(chans.items())}
dec = MSDeformAttnPixelDecoder(shapes, transformer_dropout=0.0, transformer_nheads=8, transformer_dim_feedforward=1024,
                               transformer_enc_layers=args.layers, conv_dim=256, mask_dim=256, norm="GN",
                               transformer_in_features=["res3", "res4", "res5"], common_stride=4).to(dev).train()
h3, w3 = levels[-1]
sizes = {"res5": levels[0], "res4": levels[1], "res3": levels[2], "res2": (2 * h3, 2 * w3)}
feats = {k: torch.randn(args.batch, chans[k], *sizes[k], device=dev) for k in chans}
if args.channels_last:
    feats = {k: v.contiguous(memory_format=torch.channels_last) for k, v in feats.items()}
feats = {k: v.requires_grad_(True) for k, v in feats.items()}

def set_mode(fused):
    dec.fused = fused
    for m in dec.modules():
        if m is not dec and hasattr(m, "fused"): m.fused = fused
        if hasattr(m, "fuse_prologue"): m.fuse_prologue = fused; m.tcgen05_linear = fused

def glue(fused):
    xs = [feats[f] for f in dec.transformer_in_features[::-1]]
    if fused:
        src = glue_func.input_proj_flatten(xs, dec.input_proj, args.split)
        pos = torch.cat([dec.pe_layer.tokens(h, w, xs[0]) + dec.transformer.level_embed[l].view(1, -1)
                         for l, (h, w) in enumerate(levels)], 0)[None]
    else:   # msdeformattn.py:319-322 + 66-82
        srcs = [dec.input_proj[i](x) for i, x in enumerate(xs)]
        poss = [dec.pe_layer(x) for x in xs]
        src = torch.cat([s.flatten(2).transpose(1, 2) for s in srcs], 1)
        pos = torch.cat([p.flatten(2).transpose(1, 2) + dec.transformer.level_embed[l].view(1, 1, -1)
                         for l, p in enumerate(poss)], 1)
    return src, pos

gsrc = torch.randn(args.batch, wl.S, 256, device=dev); gpos_n = torch.randn(args.batch, wl.S, 256, device=dev)

def run_glue(fused):
    src, pos = glue(fused)
    # the encoder's q = src + pos sends a batch-summed gradient to a (1, S, C) table and a per-image one to (N, S, C)
    torch.autograd.backward([src, pos], [gsrc, gpos_n[:pos.shape[0]]])

def run_full(fused):
    mf, out0, multi = dec.forward_features(feats)
    (mf.sum() + sum(m.square().mean() for m in multi)).backward()

def timeit(fn, fused):
    set_mode(fused)
    fn(fused); fn(fused); torch.cuda.synchronize()
    best = 1e9
    for _ in range(args.reps):
        for v in feats.values(): v.grad = None
        dec.zero_grad(set_to_none=True)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(fused); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best

if args.profile:
    from torch.profiler import profile, ProfilerActivity
    set_mode(True); run_glue(True); run_glue(True); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        run_glue(True); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
    sys.exit(0)

res = {}
print(f"cfg {args.cfg} batch {args.batch} levels {levels} channels_last={args.channels_last} encoder layers {args.layers}")
for label, fn in (("glue fwd+bwd (input_proj + pos + flatten)", run_glue), ("forward_features fwd+bwd", run_full)):
    for name, fused in (("fused_sm100a", True), ("reference_sequence_torch", False)):
        t = timeit(fn, fused)
        res[f"{label}|{name}"] = t
        print(f"{label:44s} {name:26s} {t:9.2f} ms  peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"decoder_bench_cfg{args.cfg}.json"), "w"))
