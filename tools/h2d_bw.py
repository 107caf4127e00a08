"""Pinned host -> device copy bandwidth of this box (what bounds bench.py's e2e number: 896 MB of inputs per C4 step)."""
import json

import torch

n = 796 * 1024 * 1024 // 4
h = torch.empty(n, dtype=torch.float32, pin_memory=True)
d = torch.empty(n, dtype=torch.float32, device="cuda")
for _ in range(2):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    d.copy_(h, non_blocking=True)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
print(json.dumps({"h2d_bytes": n * 4, "ms": ms, "GB_per_s": n * 4 / ms / 1e6}))
