"""Backward A/B through the C ABI: per-corner REDs (tuning.bwd = 1) vs anchor-sorted (bwd = 2; lanes 8 / 4, window
margins), device-resident inputs, CUDA events, L2 flushed between repetitions by the problem size itself (1.1 GB of
inputs per launch).   python tools/bwd_ab.py [--cfg 2] [--batch 16] [--reps 7]"""
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
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--dist", default="model")
    args = ap.parse_args()
    wl = W.WORKLOADS[args.cfg]
    dev = torch.device("cuda:0")
    inp = {k: v.to(dev) for k, v in W.workload_inputs(args.cfg, batch=args.batch, dist=args.dist).items()}
    N, S, M, D = inp["value"].shape
    Lq = inp["loc"].shape[1]
    dims = (N, S, M, D, wl.L, Lq, 4)
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
        return ts[len(ts) // 2], ts[0], (gv, gl, ga)

    print(f"# cfg{args.cfg} dist={args.dist} batch {N} S={S}  algorithmic HBM bytes bwd {W.hbm_bytes(S, 'bwd') * N / 1e9:.3f} GB")
    med, best, ref = bwd(cabi.make_tuning(bwd=1))
    print(f"per-corner (bwd=1)                 median {med:7.3f} ms  best {best:7.3f} ms")
    for lanes in (8, 4):
        for margin in (4, 6, 8, 12):
            med, best, got = bwd(cabi.make_tuning(bwd=2, bwd_lanes=lanes, bwd_margin=margin))
            errs = [((a - b).abs().max() / b.abs().max()).item() for a, b in zip(got, ref)]
            print(f"sorted lanes={lanes} margin={margin:2d}           median {med:7.3f} ms  best {best:7.3f} ms   "
                  f"max rel diff vs per-corner: gv {errs[0]:.1e} gl {errs[1]:.1e} ga {errs[2]:.1e}")


if __name__ == "__main__":
    main()
