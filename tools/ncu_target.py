"""Small fixed program for ncu: cfg-2 shaped inputs, 3 x (forward, backward) through the C ABI.
   python tools/ncu_target.py [batch] [tuning k=v,...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from bm2f_b200 import cabi
from bm2f_b200 import workloads as W

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
tun = None
if len(sys.argv) > 2 and sys.argv[2]:
    tun = cabi.make_tuning(**{k: int(v) for k, v in (x.split("=") for x in sys.argv[2].split(","))})
dev = torch.device("cuda:0")
inp = {k: v.to(dev) for k, v in W.workload_inputs(2, batch=batch).items()}
N, S, M, D = inp["value"].shape
dims = (N, S, M, D, 3, S, 4)
out = torch.empty(N, S, M * D, device=dev)
gv, gl, ga = torch.empty_like(inp["value"]), torch.empty_like(inp["loc"]), torch.empty_like(inp["attn"])
p = {k: v.data_ptr() for k, v in inp.items()}
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    cabi.forward(p["value"], p["shapes"], p["start"], p["loc"], p["attn"], out.data_ptr(), dims, 0, tun, st)
    cabi.backward(p["value"], p["shapes"], p["start"], p["loc"], p["attn"], p["grad_out"], gv.data_ptr(),
                  gl.data_ptr(), ga.data_ptr(), dims, 0, tun, st)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()), float(gv.abs().mean()))
