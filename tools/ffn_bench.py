"""FFN GEMM shapes (256 -> 1024 with ReLU, 1024 -> 256) forward / grad_x / grad_W: tcgen05 kernels vs cuBLAS fp32."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
import bm2f_b200
MSDA = bm2f_b200.load_extension()
dev = torch.device("cuda:0")
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 16 * 21504
def t(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best
x = torch.randn(rows, 256, device=dev); w1 = torch.randn(1024, 256, device=dev) / 16; b1 = torch.randn(1024, device=dev)
w2 = torch.randn(256, 1024, device=dev) / 32; b2 = torch.randn(256, device=dev)
h = F.relu(F.linear(x, w1, b1)); g2 = torch.randn(rows, 256, device=dev); g1 = torch.randn(rows, 1024, device=dev)
print(f"rows={rows}")
print(f"fwd  256->1024 relu : tcgen05 {t(lambda: MSDA.linear_relu_tf32x3(x, w1, b1, 3)):.3f} ms | cuBLAS fp32+relu {t(lambda: F.relu(F.linear(x, w1, b1))):.3f} ms")
print(f"fwd 1024->256       : tcgen05 {t(lambda: MSDA.linear_tf32x3(h, w2, b2, 3)):.3f} ms | cuBLAS fp32 {t(lambda: F.linear(h, w2, b2)):.3f} ms")
print(f"grad_x of 1024->256 (out 1024): tcgen05 {t(lambda: MSDA.linear_tf32x3_backward_input(g2, w2, 3)):.3f} ms | cuBLAS fp32 {t(lambda: g2 @ w2):.3f} ms")
print(f"grad_x of 256->1024 (K 1024) : tcgen05 {t(lambda: MSDA.linear_tf32x3_backward_input(g1, w1, 3)):.3f} ms | cuBLAS fp32 {t(lambda: g1 @ w1):.3f} ms")
print(f"grad_W of 256->1024 : tcgen05 {t(lambda: MSDA.linear_tf32x3_backward_weight(g1, x, 3, True)):.3f} ms | cuBLAS fp32 {t(lambda: (g1.t() @ x, g1.sum(0))):.3f} ms")
print(f"grad_W of 1024->256 : tcgen05 {t(lambda: MSDA.linear_tf32x3_backward_weight(g2, h, 3, True)):.3f} ms | cuBLAS fp32 {t(lambda: (g2.t() @ h, g2.sum(0))):.3f} ms")
print("")
