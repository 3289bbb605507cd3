#!/usr/bin/env python
"""What does the memory system sustain for in-place updates vs a copy? (GB/s, read + write bytes)"""
import torch
def t(fn, n=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(n):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
for mb in (64, 350, 1024, 4096):
    n = mb * (1 << 20) // 4
    x = torch.rand(n, device='cuda'); y = torch.empty_like(x)
    r = {}
    r['copy y<-x'] = 2 * n * 4 / t(lambda: y.copy_(x)) / 1e6
    r['in-place x*=a'] = 2 * n * 4 / t(lambda: x.mul_(1.0000001)) / 1e6
    r['out-of-place y=x*a'] = 2 * n * 4 / t(lambda: torch.mul(x, 1.0000001, out=y)) / 1e6
    r['read-only sum'] = n * 4 / t(lambda: x.sum()) / 1e6
    r['write-only fill'] = n * 4 / t(lambda: y.fill_(1.0)) / 1e6
    print('%5d MB  ' % mb + '  '.join('%s %.0f GB/s' % kv for kv in r.items()))
