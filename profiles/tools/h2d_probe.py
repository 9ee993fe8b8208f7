#!/usr/bin/env python3
"""Raw pinned host<->device copy bandwidth of this box (the ceiling of the e2e numbers)."""
import json, torch
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
res = {}
for name, (dst, src) in {"h2d": (d, h), "d2h": (h, d)}.items():
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        dst.copy_(src, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    res[name + "_GBps"] = 5 * n / (e0.elapsed_time(e1) / 1000) / 1e9
print(json.dumps(res))
