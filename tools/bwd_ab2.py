"""Backward A/B through the C ABI on one config: per-corner REDs (bwd = 1), anchor-sorted (bwd = 2, lanes 4) and the
balanced pixel-owner kernel (bwd = 2, variant 10/11) at several window margins; device-resident inputs, CUDA events.
   python tools/bwd_ab2.py [--cfg 2] [--batch N] [--reps 5]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from bm2f_b200 import cabi
from bm2f_b200 import workloads as W


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", type=int, default=2)
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--dist", default="model")
    ap.add_argument("--set", default="owner", help="owner: pixel-owner variants; cells: cell-strided phase 2 (variants 13-16 of the sorted kernel)")
    args = ap.parse_args()
    wl = W.WORKLOADS[args.cfg]
    dev = torch.device("cuda:0")
    inp = {k: v.to(dev) for k, v in W.workload_inputs(args.cfg, batch=args.batch, dist=args.dist).items()}
    N, S, M, D = inp["value"].shape
    dims = (N, S, M, D, wl.L, inp["loc"].shape[1], 4)
    st = torch.cuda.current_stream().cuda_stream
    p = {k: v.data_ptr() for k, v in inp.items()}

    def bwd(t):
        gv, gl, ga = torch.empty_like(inp["value"]), torch.empty_like(inp["loc"]), torch.empty_like(inp["attn"])
        def run():
            cabi.backward(p["value"], p["shapes"], p["start"], p["loc"], p["attn"], p["grad_out"], gv.data_ptr(),
                          gl.data_ptr(), ga.data_ptr(), dims, 0, t, st)
        run(); run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); run(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        return ts[len(ts) // 2], (gv, gl, ga)

    print(f"# cfg{args.cfg} dist={args.dist} batch {N} S={S}")
    med, ref = bwd(cabi.make_tuning(bwd=1))
    print(f"per-corner (bwd=1)              {med:7.3f} ms")
    cases = [("sorted lanes=4 margin 6", dict(bwd=2, bwd_lanes=4, bwd_margin=6))]
    if args.set == "owner":
        for v in (10, 11):
            for m in (4, 5, 6):
                cases.append((f"owner variant {v} margin {m}", dict(bwd=2, bwd_margin=m, variant=v)))
    else:
        # sorted kernel, phase 2 by cell ownership: 13 = 2 CTAs x 8 warps, 14 = + prefetch, 15 = 1 CTA x 16 warps, 16 = + prefetch
        for v, lanes in ((17, 4), (13, 4), (13, 8), (14, 4), (15, 4), (16, 4)):
            for m in ((6, 5, 8) if v in (13, 14) and lanes == 4 else (6,)):
                cases.append((f"cell-strided v{v} lanes {lanes} margin {m}", dict(bwd=2, bwd_lanes=lanes, bwd_margin=m, variant=v)))
    for name, kw in cases:
        med, got = bwd(cabi.make_tuning(**kw))
        errs = [((a - b).abs().max() / b.abs().max()).item() for a, b in zip(got, ref)]
        print(f"{name:30s}  {med:7.3f} ms   max rel diff vs per-corner: gv {errs[0]:.1e} gl {errs[1]:.1e} ga {errs[2]:.1e}")


if __name__ == "__main__":
    main()
