"""Small fixed program for compute-sanitizer: every kernel family once on tiny shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bm2f_b200 import workloads as W
from bm2f_b200.encoder import MSDeformAttnTransformerEncoderLayer
from bm2f_b200.ops.functions import MSDeformAttnFunction
dev = torch.device("cuda:0")
levels = ((3, 5), (6, 10), (12, 20))
inp = {k: v.to(dev) for k, v in W.make_inputs(levels, 2, seed=3).items()}
for dt in (torch.float32, torch.float64, torch.bfloat16):
    v = inp["value"].to(dt).requires_grad_(True)
    lt = torch.float64 if dt == torch.float64 else torch.float32
    loc = inp["loc"].to(lt).requires_grad_(True); at = inp["attn"].to(lt).requires_grad_(True)
    out = MSDeformAttnFunction.apply(v, inp["shapes"], inp["start"], loc, at, 128)
    out.backward(inp["grad_out"].to(dt))
layer = MSDeformAttnTransformerEncoderLayer(256, 1024, 0.0, "relu", 3, 8, 4).to(dev)
S = sum(h * w for h, w in levels)
src = torch.randn(2, S, 256, device=dev, requires_grad=True); pos = torch.randn(2, S, 256, device=dev)
ref = W.reference_points(levels, 2).to(dev)
y = layer(src, pos, ref, inp["shapes"], inp["start"], torch.zeros(2, S, dtype=torch.bool, device=dev))
y.sum().backward()
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
