#!/usr/bin/env python
"""Host<->device copy rates from pinned memory on this box (what bounds bench.py's e2e leg)."""
import torch
dev = torch.device('cuda:0')
for mb in (1, 2, 4, 16, 64):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    for name, src, dst in (('h2d', h, d), ('d2h', d, h)):
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 20
        print('%s %3d MB  %8.1f us  %6.1f GB/s' % (name, mb, us, n / us / 1e3))
# both directions at once on two streams
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
n = 16 << 20
h1, h2 = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
d1, d2 = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    with torch.cuda.stream(s1):
        d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / 10
print('h2d + d2h 16 MB each, concurrent: %.1f us -> %.1f GB/s per direction' % (us, n / us / 1e3))
