"""Scratch probe: raw pinned H2D bandwidth of the box (the ceiling of bench.py's e2e) and per-call latency of small scans."""
import sys, time
import torch
sys.path.insert(0, ".")
n = 4 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    d.copy_(h, non_blocking=True)
e1.record(); torch.cuda.synchronize()
print(f"H2D pinned 4 GiB x5: {5 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9:.2f} GB/s")
e0.record()
for _ in range(5):
    h.copy_(d, non_blocking=True)
e1.record(); torch.cuda.synchronize()
print(f"D2H pinned 4 GiB x5: {5 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9:.2f} GB/s")
