"""Anchor-sorted backward: A/B of the launch variants (tuning.reserved[0], msda_bwd_sorted.cu:pick) and the per-phase
cycle profile recorded by thread 0 of every CTA (bm2f_msda_debug_phase_profile).
  python tools/bwd_phases.py [--cfg 2] [--batch 16] [--ncu VARIANT,LANES]   (--ncu: run that variant 3x and exit)"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from bm2f_b200 import cabi
from bm2f_b200 import workloads as W

PHASES = ("wait loc/attn", "phase 1", "scan", "scatter", "wait grad_out", "phase 2", "phase 3")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", type=int, default=2)
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--margin", type=int, default=6)
    ap.add_argument("--ncu", default="")
    ap.add_argument("--variants", default="0,6,7,8,9")
    args = ap.parse_args()
    wl = W.WORKLOADS[args.cfg]
    dev = torch.device("cuda:0")
    inp = {k: v.to(dev) for k, v in W.workload_inputs(args.cfg, batch=args.batch).items()}
    N, S, M, D = inp["value"].shape
    Lq = inp["loc"].shape[1]
    dims = (N, S, M, D, wl.L, Lq, 4)
    st = torch.cuda.current_stream().cuda_stream
    p = {k: v.data_ptr() for k, v in inp.items()}
    gv, gl, ga = torch.empty_like(inp["value"]), torch.empty_like(inp["loc"]), torch.empty_like(inp["attn"])

    def run(t):
        cabi.backward(p["value"], p["shapes"], p["start"], p["loc"], p["attn"], p["grad_out"], gv.data_ptr(),
                      gl.data_ptr(), ga.data_ptr(), dims, 0, t, st)

    if args.ncu:
        v, lanes = (int(x) for x in args.ncu.split(","))
        t = cabi.make_tuning(bwd=2, bwd_lanes=lanes, bwd_margin=args.margin, variant=v) if v >= 0 else cabi.make_tuning(bwd=1)
        for _ in range(3):
            run(t)
        torch.cuda.synchronize()
        return

    sms = torch.cuda.get_device_properties(0).multi_processor_count
    prof = torch.zeros(2 * sms * 8, dtype=torch.int64, device=dev)
    print(f"# cfg{args.cfg} batch {N} S={S} margin {args.margin}; cycles of thread 0 per CTA, mean over CTAs, in k-cycles")
    print(f"{'variant':>8} {'lanes':>5} {'ms':>7} | " + " ".join(f"{n:>13}" for n in PHASES) + f" {'total':>9} {'pts/CTA':>9}")
    for v in (int(x) for x in args.variants.split(",")):
        for lanes in ((8, 4) if v < 6 else (4,) if v >= 13 else (0,)):
            t = cabi.make_tuning(bwd=2, bwd_lanes=lanes, bwd_margin=args.margin, variant=v)
            cabi.lib().bm2f_msda_debug_phase_profile(0)
            run(t); run(t)
            torch.cuda.synchronize()
            ts = []
            for _ in range(args.reps):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); run(t); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            ts.sort()
            prof.zero_()
            cabi.lib().bm2f_msda_debug_phase_profile(prof.data_ptr())
            run(t)
            torch.cuda.synchronize()
            cabi.lib().bm2f_msda_debug_phase_profile(0)
            pr = prof.view(-1, 8).double()
            pr = pr[pr[:, :7].sum(1) > 0]
            mean = pr.mean(0)
            print(f"{v:>8} {lanes:>5} {ts[len(ts) // 2]:7.3f} | " + " ".join(f"{mean[i].item() / 1e3:13.1f}" for i in range(7)) +
                  f" {mean[:7].sum().item() / 1e3:9.1f} {mean[7].item():9.0f}   ({pr.shape[0]} CTAs)")


if __name__ == "__main__":
    main()
