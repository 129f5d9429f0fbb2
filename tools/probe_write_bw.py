"""What does a pure WRITE stream reach on this GPU?  (The per-site kernel and the S = 1 sliding tile write 36-76 bytes of
rows per 1-41 bytes read; the copy peak of MEASURED_PEAKS.json is half reads, half writes.)
Times torch fill_ (one stream) and eight interleaved fills of separate arrays (the shape of a row table: 8 output columns)."""
import torch

n = 1 << 29  # 4 GiB of float64
x = torch.empty(n, dtype=torch.float64, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for _ in range(3):
    x.fill_(1.0)
torch.cuda.synchronize()
ev[0].record()
for _ in range(10):
    x.fill_(2.0)
ev[1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / 10
print(f"fill_ 4 GiB: {ms:.3f} ms = {n * 8 / ms / 1e6:.0f} GB/s")
y = torch.empty(n, dtype=torch.float64, device="cuda")
for _ in range(3):
    y.copy_(x)
torch.cuda.synchronize()
ev[0].record()
for _ in range(10):
    y.copy_(x)
ev[1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / 10
print(f"copy 4 GiB -> 4 GiB: {ms:.3f} ms = {2 * n * 8 / ms / 1e6:.0f} GB/s (read + write)")
