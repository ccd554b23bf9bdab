"""Pinned host -> device copy bandwidth of the box (what bounds bench.py's e2e leg): python tools/h2d_bw.py"""
import json
import torch

n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
res = {}
for name, size in (("1GiB", n), ("150MB", 150_528_000)):
    for _ in range(2):
        d[:size].copy_(h[:size], non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8):
        d[:size].copy_(h[:size], non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    res[name + "_GBps"] = round(8 * size / (e0.elapsed_time(e1) * 1e-3) / 1e9, 2)
print(json.dumps(res))
