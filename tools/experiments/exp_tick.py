#!/usr/bin/env python
"""Kernel experiments: time the tick alone for a chosen config / batch shape and report algorithmic GB/s.
usage: tools/exp_tick.py [--games N] [--cap K] [--reload-time T] [--solo] [--steps S] [--flags F]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from astro_b200 import core
from astro_b200 import _native as nat
from astro_b200.batched import BatchedGames
from astro_b200.pool import make_pool
import bench
ap = argparse.ArgumentParser()
ap.add_argument('--games', type=int, default=1 << 20)
ap.add_argument('--cap', type=int, default=32)
ap.add_argument('--reload-time', type=float, default=None)
ap.add_argument('--solo', action='store_true')
ap.add_argument('--steps', type=int, default=300)
ap.add_argument('--preroll', type=int, default=600)
ap.add_argument('--flags', type=int, default=0)
ap.add_argument('--device-controls', action='store_true')
a = ap.parse_args()
cfg = core.SOLO_CONFIG if a.solo else core.DEFAULT_CONFIG
if a.reload_time is not None:
    cfg = cfg._replace(reload_time=a.reload_time)
S = 1 if cfg.solo else 2
games = BatchedGames(cfg, a.games, bullet_cap=a.cap, precision=32, device=0)
pool = make_pool(cfg, 4096)
games.set_reset_pool_arrays(pool['ships'], pool['planets'], pool['np']); games.reset_all()
flags = nat.TICK_AUTO_RESET | a.flags
for _ in range(a.preroll): games.step_raw(0, flags)
ring = torch.randint(0, 6, (8, games.n_pad, S), dtype=torch.uint8).cuda()
ptrs = [0] * 8 if a.device_controls else [ring[i].data_ptr() for i in range(8)]
for k in range(20): games.step_raw(ptrs[k % 8], flags)
games.stats(clear=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for k in range(a.steps): games.step_raw(ptrs[k % 8], flags)
e1.record(); torch.cuda.synchronize()
st = games.stats(clear=True)
us = 1e3 * e0.elapsed_time(e1) / a.steps
alg = bench.algorithmic_bytes(st, S, actions_in_hbm=not a.device_controls)
print(json.dumps(dict(games=a.games, cap=a.cap, S=S, reload_time=cfg.reload_time, us_per_tick=round(us, 2),
                      bytes_per_step=round(alg / st['env_steps'], 1), GBps=round(alg / a.steps / us / 1e3, 1),
                      frac=round(alg / a.steps / us / 1e3 / 6548.8, 3), mean_bullets=round(st['bullets_in'] / st['env_steps'], 2),
                      env_steps_per_s='%.4g' % (st['env_steps'] / a.steps / us * 1e6))))
