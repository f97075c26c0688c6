"""Small fixed program for ncu: the persistent tcgen05 GEMM at cfg-2 row count, 256 -> 256 and 1024 -> 256, plus a clock /
power sample under a 2-second GEMM loop (python tools/ncu_gemm_target.py clocks)."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bm2f_b200
MSDA = bm2f_b200.load_extension()
dev = torch.device("cuda:0"); torch.manual_seed(0)
split = int(os.environ.get("BM2F_SPLIT", "3"))
rows = 16 * 21504
xs = {k: torch.randn(rows, k, device=dev) for k in (256, 1024)}
ws = {k: torch.randn(256, k, device=dev) / k ** 0.5 for k in (256, 1024)}
b = torch.randn(256, device=dev)
for _ in range(3):
    for k in (256, 1024):
        y = MSDA.linear_tf32x3(xs[k], ws[k], b, split)
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
if len(sys.argv) > 1 and sys.argv[1] == "clocks":
    q = "clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu,clocks_throttle_reasons.active"
    for k in (256, 1024):
        t0 = time.time(); n = 0; samples = []
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        while time.time() - t0 < 3.0:
            for _ in range(50):
                MSDA.linear_tf32x3(xs[k], ws[k], b, 3); n += 1
            samples.append(subprocess.run(["nvidia-smi", "--query-gpu=" + q, "--format=csv,noheader"],
                                          capture_output=True, text=True).stdout.strip())
        e.record(); torch.cuda.synchronize()
        print(f"K={k}: {a.elapsed_time(e) / n:.3f} ms / GEMM over {n} launches (incl. nvidia-smi pauses)")
        for s_ in samples[:2] + samples[-3:]:
            print("   ", s_)
