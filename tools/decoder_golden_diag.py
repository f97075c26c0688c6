"""Per-tensor error report of the pixel-decoder mirror against the reference golden vectors (max-relative, L2-relative,
number of entries off by more than 1e-4 of the tensor's scale) for the fused path and the reference op sequence."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import tests.test_gpu_decoder as T
torch.backends.cudnn.allow_tf32 = False
z = np.load(T.GOLD)
for fused in (True, False):
    dec, G = T._decoder(fused)
    got = T._run_decoder(dec, G)
    print(f"--- fused={fused}")
    for k in got:
        d = np.abs(got[k].astype(np.float64) - z[k]); s = np.abs(z[k]).max()
        l2 = np.sqrt((d ** 2).sum()) / np.sqrt((z[k].astype(np.float64) ** 2).sum())
        print(f"{k:70s} max-rel {d.max() / s:9.2e}  l2-rel {l2:9.2e}  n(>1e-4 scale) {(d > 1e-4 * s).sum():6d} / {d.size}")
