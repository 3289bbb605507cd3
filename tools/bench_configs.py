#!/usr/bin/env python
"""Throughput of the BASELINE.json configs that bench.py does not time itself (one B200):
  #2  4,096 games, random controls           — many ticks per launch / one launch per tick
  #3  65,536 games, bullet pool K = 400 and K = 32: the stationary population, and the first ticks after every pool was filled
      to capacity (the stress state of tests/test_gpu_parity.py::_stress_fill: spawn / despawn compaction at full load)
Prints one JSON line per case with env-steps/s and algorithmic GB/s (device counters)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from astro_b200 import core
from astro_b200.batched import BatchedGames
from astro_b200.pool import make_pool

cfg = core.DEFAULT_CONFIG
pool = make_pool(cfg, 1024)


def alg_bytes(st):
    return st['env_steps'] * (2 * (4 + 32 + 8) + 2 + 1) + 32 * st['planets_live'] + 16 * (st['bullets_in'] + st['bullets_out'])


def timed(g, fuse, ticks):
    g.stats(clear=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(ticks // fuse):
        g.step_many(fuse, None, auto_reset=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    st = g.stats(clear=True)
    return dict(env_steps_per_s=st['env_steps'] / (ms * 1e-3), us_per_tick=1e3 * ms / (ticks // fuse * fuse), algorithmic_GBps=alg_bytes(st) / (ms * 1e-3) / 1e9,
                mean_bullets=st['bullets_in'] / max(1, st['env_steps']), overflow=st['overflow'])


def fresh(n, K):
    g = BatchedGames(cfg, n, bullet_cap=K, precision=32, seed=0)
    g.set_reset_pool_arrays(pool['ships'], pool['planets'], pool['np'])
    g.reset_all()
    return g


for n, K, label in ((4096, 32, '#2 4,096 games'), (65536, 400, '#3 65,536 games, K=400'), (65536, 32, '#3 65,536 games, K=32')):
    g = fresh(n, K)
    g.step_many(600, None, auto_reset=True)
    for fuse in (64, 1):
        print(json.dumps(dict(config=label, state='stationary population', ticks_per_launch=fuse, **timed(g, fuse, 640 if fuse > 1 else 300))), flush=True)
    if n == 65536:
        # every pool filled to capacity with bullets spread over the arena (most fly on for dozens of ticks)
        r = np.random.RandomState(3)
        arr = g.get_arrays()
        bl = np.concatenate([r.uniform(-1.0, 1.0, (n, K, 2)), r.uniform(-0.3, 0.3, (n, K, 2))], axis=2)
        g.set_arrays(arr['ships'], arr['planets'], arr['n_planets'], bl, np.full(n, K), arr['tick'])
        print(json.dumps(dict(config=label, state='every pool filled to capacity, first 16 ticks', ticks_per_launch=1, **timed(g, 1, 16))), flush=True)
    del g
