import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bm2f_b200
MSDA = bm2f_b200.load_extension()
dev = torch.device("cuda:0")
x = torch.randn(344064, 256, device=dev); w = torch.randn(256, 256, device=dev) / 16; b = torch.randn(256, device=dev)
for _ in range(3):
    y = MSDA.linear_tf32x3(x, w, b, 3)
torch.cuda.synchronize(); print("ok", float(y.abs().mean()))
