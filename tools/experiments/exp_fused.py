#!/usr/bin/env python
"""astro_tick_many against one launch per tick: us per tick for T ticks per launch (1M games, stationary population)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from astro_b200 import core
from astro_b200 import _native as nat
from astro_b200.batched import BatchedGames
from astro_b200.pool import make_pool
ap = argparse.ArgumentParser()
ap.add_argument('--games', type=int, default=1 << 20)
ap.add_argument('--preroll', type=int, default=600)
ap.add_argument('--ticks', type=int, default=512)
a = ap.parse_args()
games = BatchedGames(core.DEFAULT_CONFIG, a.games, bullet_cap=32, precision=32, device=0)
pool = make_pool(core.DEFAULT_CONFIG, 4096)
games.set_reset_pool_arrays(pool['ships'], pool['planets'], pool['np']); games.reset_all()
flags = nat.TICK_AUTO_RESET
R = 64
ring = torch.randint(0, 6, (R, games.n_pad, 2), dtype=torch.uint8).cuda()
ev = torch.zeros((R, games.n_pad), dtype=torch.uint8, device='cuda')
games.step_many_raw(0, 0, a.preroll, flags)
torch.cuda.synchronize()
for T in (1, 2, 4, 8, 16, 32, 64):
    for name, ap_, ep_ in (('ring controls + events', ring.data_ptr(), ev.data_ptr()),):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        games.step_many_raw(ap_, ep_, T, flags)
        torch.cuda.synchronize()
        e0.record()
        for k in range(a.ticks // T):
            games.step_many_raw(ap_, ep_, T, flags)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (a.ticks // T * T)
        print('T=%2d  %-24s %.2f us per tick  %.3e env-steps/s' % (T, name, us, a.games / us * 1e6), flush=True)
