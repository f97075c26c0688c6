"""Variant sweep of the D=32 fast kernels through the C ABI (device-resident inputs, CUDA events).
Prints a table and writes gpurun_out/sweep_<tag>.json.   python tools/sweep.py [--cfg 2] [--batch 16]"""
import argparse
import itertools
import json
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
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--tag", default="r01")
    args = ap.parse_args()
    wl = W.WORKLOADS[args.cfg]
    dev = torch.device("cuda:0")
    inp = {k: v.to(dev) for k, v in W.workload_inputs(args.cfg, batch=args.batch).items()}
    N, S, M, D = inp["value"].shape
    Lq = inp["loc"].shape[1]
    dims = (N, S, M, D, wl.L, Lq, 4)
    out = torch.empty(N, Lq, M * D, device=dev)
    gv, gl, ga = torch.empty_like(inp["value"]), torch.empty_like(inp["loc"]), torch.empty_like(inp["attn"])
    st = torch.cuda.current_stream().cuda_stream
    p = {k: v.data_ptr() for k, v in inp.items()}

    def fwd(t):
        cabi.forward(p["value"], p["shapes"], p["start"], p["loc"], p["attn"], out.data_ptr(), dims, 0, t, st)

    def bwd(t):
        cabi.backward(p["value"], p["shapes"], p["start"], p["loc"], p["attn"], p["grad_out"], gv.data_ptr(),
                      gl.data_ptr(), ga.data_ptr(), dims, 0, t, st)

    def timeit(fn, t):
        fn(t); fn(t)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(args.reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(t); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        return best

    grid = dict(vec=(8, 4), staging=(1, 2), strip_w=(16, 32), ctas_per_sm=(1, 2), rows=(0,), order=(0,), merge=(1, 2))
    if args.quick:
        grid = dict(vec=(8, 4), staging=(1,), strip_w=(32,), ctas_per_sm=(1,), rows=(0,), order=(0,), merge=(1, 2))
    res = []
    gb_f = W.gather_bytes(S, "fwd") * N / 1e9
    print(f"# cfg{args.cfg} batch {N} S={S}; gather bytes fwd {gb_f:.2f} GB; hbm fwd {W.hbm_bytes(S,'fwd')*N/1e9:.3f} GB "
          f"bwd {W.hbm_bytes(S,'bwd')*N/1e9:.3f} GB")
    print(f"{'vec':>3} {'stg':>3} {'sw':>3} {'cps':>3} {'rows':>4} {'ord':>3} {'mrg':>3} | {'fwd ms':>8} {'bwd ms':>8} {'fwd lineTB/s':>12} {'img/s(6L)':>10}")
    for vals in itertools.product(*grid.values()):
        kw = dict(zip(grid.keys(), vals))
        t = cabi.make_tuning(**kw)
        try:
            f, b = timeit(fwd, t), timeit(bwd, t)
        except cabi.MSDAError as e:
            print(kw, "ERR", e); continue
        r = dict(kw, fwd_ms=f, bwd_ms=b, img_s=N / (6 * (f + b) * 1e-3))
        res.append(r)
        print(f"{kw['vec']:>3} {kw['staging']:>3} {kw['strip_w']:>3} {kw['ctas_per_sm']:>3} {kw['rows']:>4} {kw['order']:>3} {kw['merge']:>3} | "
              f"{f:8.3f} {b:8.3f} {gb_f / f:12.2f} {r['img_s']:10.1f}", flush=True)
    gen = cabi.make_tuning(force_generic=1)
    f, b = timeit(fwd, gen), timeit(bwd, gen)
    print(f"generic kernel: fwd {f:.3f} ms bwd {b:.3f} ms")
    res.append(dict(generic=1, fwd_ms=f, bwd_ms=b))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"sweep_{args.tag}_cfg{args.cfg}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
