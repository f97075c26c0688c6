"""Host<->device copy rates of this box with pinned memory: H2D alone, D2H alone, both directions at once.
The `e2e` number of bench.py moves 1.1 GB each way per layer call; this is the ceiling it can reach."""
import torch
dev = torch.device("cuda:0")
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device=dev); d_out = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def run(up, down, reps=5):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s1.wait_event(a); s2.wait_event(a)
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if down:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    b.record(); torch.cuda.synchronize()
    return reps * n / (a.elapsed_time(b) * 1e-3) / 1e9

run(True, True, 1)
print(f"H2D alone           {run(True, False):6.1f} GB/s")
print(f"D2H alone           {run(False, True):6.1f} GB/s")
print(f"both, each direction {run(True, True):6.1f} GB/s")
