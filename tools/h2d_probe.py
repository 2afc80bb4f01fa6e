"""Measure pinned host->device copy bandwidth on this box (the ceiling of bench.py's e2e arm)."""
import time
import torch
for mb in (64, 256, 1024, 3072):
    h = torch.empty(mb << 20, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(mb << 20, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    print(f"H2D pinned {mb} MiB: {5 * (mb << 20) / 1e9 / (e0.elapsed_time(e1) / 1e3):.1f} GB/s")
    del h, d
