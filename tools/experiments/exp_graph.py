#!/usr/bin/env python
"""Experiment: back-to-back astro_tick launches vs the same ticks replayed from a CUDA graph."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from astro_b200 import core
from astro_b200 import _native as nat
from astro_b200.batched import BatchedGames
from astro_b200.pool import make_pool
n = 1 << 20
games = BatchedGames(core.DEFAULT_CONFIG, n, bullet_cap=32, precision=32, device=0)
pool = make_pool(core.DEFAULT_CONFIG, 4096)
games.set_reset_pool_arrays(pool['ships'], pool['planets'], pool['np']); games.reset_all()
flags = nat.TICK_AUTO_RESET
for _ in range(900): games.step_raw(0, flags)
R = 8
ring = torch.randint(0, 6, (R, n, 2), dtype=torch.uint8, device='cuda')
ptrs = [ring[i].data_ptr() for i in range(R)]
def timed(fn, reps):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); fn(reps); e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)
def plain(reps):
    for k in range(reps): games.step_raw(ptrs[k % R], flags)
plain(50)
t_plain = timed(plain, 400) / 400
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    plain(16)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for k in range(R): games.step_raw(ptrs[k], flags)
    def replay(reps):
        for _ in range(reps // R): g.replay()
    replay(48)
    t_graph = timed(replay, 400) / 400
print(json.dumps(dict(us_per_tick_plain=1e3 * t_plain, us_per_tick_graph=1e3 * t_graph)))
