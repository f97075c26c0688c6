"""A/B of the persistent tcgen05 GEMM: single CTAs (split 3) vs clusters of two CTAs with TMA-multicast weights (split 33).
Checks results against float64 on ragged row counts first, then times the encoder's GEMM shapes at cfg-2 size."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bm2f_b200
MSDA = bm2f_b200.load_extension()
dev = torch.device("cuda:0"); torch.manual_seed(0)

def err(a, ref): return ((a.double() - ref).abs().max() / ref.abs().max()).item()

def V(variant, fn):
    """run fn with the GEMM A/B variant `variant` (bm2f_linear_tuning_t.variant), then restore the default"""
    MSDA.linear_set_tuning(variant, 0)
    try:
        return fn()
    finally:
        MSDA.linear_set_tuning(0, 0)


for rows in (1, 127, 129, 128 * 5 + 17, 128 * 296, 128 * 297 + 5):
    for k, n in ((256, 256), (256, 192), (256, 96), (256, 1024), (1024, 256)):
        x = torch.randn(rows, k, device=dev); w = torch.randn(n, k, device=dev) / k ** 0.5; b = torch.randn(n, device=dev)
        ref = x.double() @ w.double().t() + b.double()
        e1, e2 = err(MSDA.linear_tf32x3(x, w, b, 3), ref), err(V(3, lambda: MSDA.linear_tf32x3(x, w, b, 3)), ref)
        assert e2 < 1e-5, (rows, k, n, e1, e2)
print("cluster variant matches float64 on all ragged shapes", flush=True)

def t(fn, reps=10):
    fn(); fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best

rows = 16 * 21504
for k, n in ((256, 256), (256, 192), (256, 96), (256, 1024), (1024, 256)):
    x = torch.randn(rows, k, device=dev); w = torch.randn(n, k, device=dev) / k ** 0.5; b = torch.randn(n, device=dev)
    t1 = t(lambda: MSDA.linear_tf32x3(x, w, b, 3)); t2 = t(lambda: V(3, lambda: MSDA.linear_tf32x3(x, w, b, 3)))
    line = f"{k:4d} -> {n:4d}: single {t1:.3f} ms | cluster-2 multicast {t2:.3f} ms"
    if n == 1024:
        t3 = t(lambda: MSDA.linear_relu_tf32x3(x, w, b, 3)); t4 = t(lambda: V(3, lambda: MSDA.linear_relu_tf32x3(x, w, b, 3)))
        line += f" | +relu single {t3:.3f} | +relu cluster {t4:.3f}"
    print(line, flush=True)

print("activation bytes in flight: producer warps x k-blocks in registers", flush=True)
for k, n in ((256, 256), (256, 1024), (1024, 256)):
    x = torch.randn(rows, k, device=dev); w = torch.randn(n, k, device=dev) / k ** 0.5; b = torch.randn(n, device=dev)
    ref = x[:4096].double() @ w.double().t() + b.double()
    line = f"{k:4d} -> {n:4d}:"
    for name, sp in (("4w x 3", 3), ("8w x 5", 43), ("4w x 4", 53), ("8w x 4", 63)):
        y = V(sp // 10, lambda: MSDA.linear_tf32x3(x, w, b, sp % 10))
        assert err(y[:4096], ref) < 1e-5, (k, n, sp)
        line += f"  {name} {t(lambda: V(sp // 10, lambda: MSDA.linear_tf32x3(x, w, b, sp % 10))):.3f} ms"
    print(line, flush=True)

print("CTA pairs (tcgen05 cta_group::2), split 73", flush=True)
for rws in (1, 129, 128 * 5 + 17, 128 * 297 + 5):
    for k, n in ((256, 256), (256, 192), (256, 96), (256, 1024), (1024, 256)):
        x = torch.randn(rws, k, device=dev); w = torch.randn(n, k, device=dev) / k ** 0.5; b = torch.randn(n, device=dev)
        ref = x.double() @ w.double().t() + b.double()
        e = err(V(7, lambda: MSDA.linear_tf32x3(x, w, b, 3)), ref)
        torch.cuda.synchronize()
        print(f"  rows {rws:6d} {k:4d}->{n:4d} err {e:.2e}", flush=True)
        assert e < 1e-5
for k, n in ((256, 256), (256, 192), (256, 96), (256, 1024), (1024, 256)):
    x = torch.randn(rows, k, device=dev); w = torch.randn(n, k, device=dev) / k ** 0.5; b = torch.randn(n, device=dev)
    t1 = t(lambda: MSDA.linear_tf32x3(x, w, b, 3)); t2 = t(lambda: V(7, lambda: MSDA.linear_tf32x3(x, w, b, 3)))
    print(f"{k:4d} -> {n:4d}: single {t1:.3f} ms | CTA pair {t2:.3f} ms", flush=True)

print("single TF32 pass: register-staged activations (split 51) vs TMA-loaded activations (split 1, default)", flush=True)
for rws in (1, 129, 128 * 5 + 17, 128 * 297 + 5):
    for k, n in ((256, 256), (256, 192), (256, 96), (256, 1024), (1024, 256)):
        x = torch.randn(rws, k, device=dev); w = torch.randn(n, k, device=dev) / k ** 0.5; b = torch.randn(n, device=dev)
        ref = x.double() @ w.double().t() + b.double()
        e51 = err(V(5, lambda: MSDA.linear_tf32x3(x, w, b, 1)), ref) if n % 256 == 0 else float("nan")
        e1 = err(MSDA.linear_tf32x3(x, w, b, 1), ref)
        torch.cuda.synchronize()
        assert e1 < 3e-3, (rws, k, n, e51, e1)
print("   errors vs float64 (last shape): register-staged %.2e, TMA %.2e" % (e51, e1), flush=True)
for k, n in ((256, 256), (256, 192), (256, 96), (256, 1024), (1024, 256)):
    x = torch.randn(rows, k, device=dev); w = torch.randn(n, k, device=dev) / k ** 0.5; b = torch.randn(n, device=dev)
    t3 = t(lambda: MSDA.linear_tf32x3(x, w, b, 3)); t1 = t(lambda: MSDA.linear_tf32x3(x, w, b, 1))
    t51 = t(lambda: V(5, lambda: MSDA.linear_tf32x3(x, w, b, 1))) if n % 256 == 0 else float("nan")
    torch.backends.cuda.matmul.allow_tf32 = True
    tc = t(lambda: torch.nn.functional.linear(x, w, b))
    torch.backends.cuda.matmul.allow_tf32 = False
    print(f"{k:4d} -> {n:4d}: tf32x3 {t3:.3f} ms | tf32x1 register-staged {t51:.3f} ms | tf32x1 TMA activations {t1:.3f} ms | cuBLAS TF32 {tc:.3f} ms", flush=True)

