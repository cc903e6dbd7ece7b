#!/usr/bin/env python
"""Does the H2D rate depend on WHICH pinned buffer is copied?  Allocates several pinned 12 MB buffers (all alive),
times each one's steady-state H2D, three rounds.  Diagnostic for the run-to-run spread of bench.py's e2e leg."""
import json

import torch

n = 12_000_000
d = torch.empty(n, dtype=torch.uint8, device="cuda")


def rate(h, reps=12):
    for _ in range(6):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        d.copy_(h, non_blocking=True)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return round(n / (ts[len(ts) // 2] * 1e-3) / 1e9, 1)


bufs = []
for i in range(10):
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    h.fill_(i)
    bufs.append(h)
res = {"torch_pin_memory": [[rate(h) for h in bufs] for _ in range(3)]}
# one large pinned allocation cut into 12 MB pieces
big = torch.empty(16 * n, dtype=torch.uint8).pin_memory()
big.fill_(3)
res["slices_of_one_192MB_buffer"] = [rate(big[k * n:(k + 1) * n]) for k in range(16)]
# three 4 MB copies issued back to back (what igd_b200_count does) from the first buffer
h = bufs[0]
parts = [h[k * 4_000_000:(k + 1) * 4_000_000] for k in range(3)]
dd = [torch.empty(4_000_000, dtype=torch.uint8, device="cuda") for _ in range(3)]
ts = []
for _ in range(12):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for x, y in zip(dd, parts):
        x.copy_(y, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ts.sort()
res["three_4MB_copies_gbs"] = round(n / (ts[len(ts) // 2] * 1e-3) / 1e9, 1)
print(json.dumps(res))
