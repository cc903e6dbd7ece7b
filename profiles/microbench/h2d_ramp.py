#!/usr/bin/env python
"""H2D bandwidth of consecutive 12 MB pinned copies from a cold start, and again after idle gaps: shows how the
PCIe link (and the copy path) warm up.  Diagnostic for bench.py's e2e warm-up."""
import json
import time

import torch

n = 12_000_000
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h.fill_(1)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()


def burst(k):
    out = []
    for _ in range(k):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        d.copy_(h, non_blocking=True)
        b.record()
        torch.cuda.synchronize()
        out.append(round(n / (a.elapsed_time(b) * 1e-3) / 1e9, 1))
    return out


res = {"cold": burst(40)}
for gap in (0.001, 0.01, 0.1, 1.0):
    time.sleep(gap)
    res[f"after_{gap}s_idle"] = burst(8)
# with kernels in between (the e2e step: copy, 0.25 ms of kernels, copy ...)
x = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
seq = []
for _ in range(12):
    x.fill_(1)
    torch.cuda.synchronize()
    seq += burst(1)
res["interleaved_with_256MB_fill"] = seq
print(json.dumps(res))
