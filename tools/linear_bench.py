"""tcgen05 projection GEMM vs torch F.linear (fp32 cuBLAS, TF32 off / on) at the cfg-2 row count."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
import bm2f_b200
MSDA = bm2f_b200.load_extension()
dev = torch.device("cuda:0")
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 16 * 21504
x = torch.randn(rows, 256, device=dev)
def t(fn, reps=10):
    fn(); fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best
for n in (256, 288, 192, 96):
    w = torch.randn(n, 256, device=dev) / 16; b = torch.randn(n, device=dev)
    ref = x.double() @ w.double().t() + b.double()
    torch.backends.cuda.matmul.allow_tf32 = False
    t_fp32 = t(lambda: F.linear(x, w, b)); e_fp32 = ((F.linear(x, w, b).double() - ref).abs().max() / ref.abs().max()).item()
    torch.backends.cuda.matmul.allow_tf32 = True
    t_tf32 = t(lambda: F.linear(x, w, b)); e_tf32 = ((F.linear(x, w, b).double() - ref).abs().max() / ref.abs().max()).item()
    torch.backends.cuda.matmul.allow_tf32 = False
    t3 = t(lambda: MSDA.linear_tf32x3(x, w, b, 3)); e3 = ((MSDA.linear_tf32x3(x, w, b, 3).double() - ref).abs().max() / ref.abs().max()).item()
    MSDA.linear_set_tuning(1, 0)          # A/B variants: bm2f_linear_tuning_t.variant
    t13 = t(lambda: MSDA.linear_tf32x3(x, w, b, 3))
    MSDA.linear_set_tuning(2, 0)
    t23 = t(lambda: MSDA.linear_tf32x3(x, w, b, 3)); e23 = ((MSDA.linear_tf32x3(x, w, b, 3).double() - ref).abs().max() / ref.abs().max()).item()
    MSDA.linear_set_tuning(0, 0)
    t1 = t(lambda: MSDA.linear_tf32x3(x, w, b, 1)); e1 = ((MSDA.linear_tf32x3(x, w, b, 1).double() - ref).abs().max() / ref.abs().max()).item()
    gb = rows * (256 + n) * 4 / 1e9
    fl = 2.0 * rows * 256 * n / 1e12
    print(f"N={n:3d} rows={rows}: cuBLAS fp32 {t_fp32:.3f} ms (err {e_fp32:.1e}) | cuBLAS tf32 {t_tf32:.3f} ms (err {e_tf32:.1e}) | "
          f"tcgen05 tf32x3 persistent {t3:.3f} ms (err {e3:.1e}, {gb/t3*1e3:.0f} GB/s, {3*fl/t3*1e3:.0f} TF/s tf32) | "
          f"one-tile {t13:.3f} ms | stg-epilogue {t23:.3f} ms (err {e23:.1e}) | tf32x1 persistent {t1:.3f} ms (err {e1:.1e}, {gb/t1*1e3:.0f} GB/s)")
# ---- weight gradient
for n in (256, 192, 96):
    g = torch.randn(rows, n, device=dev)
    ref = g.double().t() @ x.double()
    t_fp32 = t(lambda: (g.t() @ x, g.sum(0)))
    t_dw = t(lambda: MSDA.linear_tf32x3_backward_weight(g, x, 3, True))
    e = ((MSDA.linear_tf32x3_backward_weight(g, x, 3, True)[0].double() - ref).abs().max() / ref.abs().max()).item()
    e32 = (((g.t() @ x).double() - ref).abs().max() / ref.abs().max()).item()
    print(f"dW N={n:3d} rows={rows}: cuBLAS fp32 (+sum) {t_fp32:.3f} ms (err {e32:.1e}) | tcgen05 tf32x3 {t_dw:.3f} ms (err {e:.1e}, {rows*(256+n)*4/1e9/t_dw*1e3:.0f} GB/s)")
# ---- weight gradient: accumulation-chain length A/B (rows reduced per CTA capped at 256 * c)
g = torch.randn(rows, 256, device=dev)
ref = g.double().t() @ x.double()
for c in (0, 16, 8, 4, 2):
    MSDA.linear_set_tuning(0, c)          # bm2f_linear_tuning_t.dw_row_cap
    t_dw = t(lambda: MSDA.linear_tf32x3_backward_weight(g, x, 3, True))
    e = ((MSDA.linear_tf32x3_backward_weight(g, x, 3, True)[0].double() - ref).abs().max() / ref.abs().max()).item()
    MSDA.linear_set_tuning(0, 0)
    print(f"dW N=256 rows={rows} row cap {256 * c if c else 'none'}: {t_dw:.3f} ms (err {e:.1e})")
