"""torch.profiler breakdown of one fused encoder layer fwd+bwd at cfg 2 (kernel-time table)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from bm2f_b200 import workloads as W
from bm2f_b200.encoder import MSDeformAttnTransformerEncoderLayer
wl = W.WORKLOADS[2]; dev = torch.device("cuda:0"); B = 16
layer = MSDeformAttnTransformerEncoderLayer(256, 1024, 0.0, "relu", 3, 8, 4).to(dev)
shapes, start = W.level_tensors(wl.levels, dev)
src = torch.randn(B, wl.S, 256, device=dev, requires_grad=True); pos = torch.randn(B, wl.S, 256, device=dev)
ref = W.reference_points(wl.levels, B).to(dev); mask = torch.zeros(B, wl.S, dtype=torch.bool, device=dev)
go = torch.randn(B, wl.S, 256, device=dev)
def run():
    out = layer(src, pos, ref, shapes, start, mask); out.backward(go)
for _ in range(3): run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    run(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
