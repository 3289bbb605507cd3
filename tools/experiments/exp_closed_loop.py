"""Closed-loop host stepping (astro_tick_host) under different slice counts / io forms: us per tick at 1M games."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from astro_b200 import core
from astro_b200.batched import BatchedGames
from astro_b200.pool import make_pool

cfg, N = core.DEFAULT_CONFIG, 1 << 20
pool = make_pool(cfg, 1024)
g = BatchedGames(cfg, N, bullet_cap=32, precision=32, seed=0)
g.set_reset_pool_arrays(pool['ships'], pool['planets'], pool['np'])
g.reset_all()
for _ in range(300):
    g.step(None, auto_reset=True)
gen = torch.Generator().manual_seed(0)
acts = torch.randint(0, 6, (16, g.n_pad, 2), dtype=torch.uint8, generator=gen)
packed = g.pack_controls(acts).contiguous().pin_memory()
acts = acts.pin_memory()
planes = torch.zeros(g.planes_shape(), dtype=torch.int32).pin_memory()
evb = torch.zeros(g.n_pad, dtype=torch.uint8).pin_memory()
for slices in (1, 2, 3, 4, 6, 8):
    os.environ['ASTRO_HOST_SLICES'] = str(slices)
    for form in ('packed+planes', 'bytes'):
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for k in range(64):
                if form == 'bytes':
                    g.step_host(acts[k % 16], evb, auto_reset=True)
                else:
                    g.step_host(packed[k % 16], planes, auto_reset=True, packed=True, planes=True)
            dt = time.perf_counter() - t0
        print('slices %d  %-14s %.1f us/tick  %.3g env-steps/s' % (slices, form, 1e6 * dt / 64, N * 64 / dt), flush=True)
