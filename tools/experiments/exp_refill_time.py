import os, sys, time
sys.path.insert(0, '/root/repo')
import torch
from astro_b200 import core, _native as nat
from astro_b200.batched import BatchedGames
cfg, N = core.DEFAULT_CONFIG, 1 << 20
g = BatchedGames(cfg, N, bullet_cap=32, precision=32, seed=0)
g.enable_fresh_games(quota=48)
g.reset_all()
for _ in range(10):
    g.step_many(20, None, auto_reset=True)
L = nat.lib()
e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e[0].record()
    g.step_many(20, None, auto_reset=True)      # (includes a refill before the launch)
    e[1].record()
    t1 = time.perf_counter()
    L.astro_fresh_games_refill(g._h, g._stream())
    e[2].record()
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    print('launch+refill %.1f us (host %.1f)   refill alone %.1f us (host %.1f)' % (1e3 * e[0].elapsed_time(e[1]), 1e6 * (t1 - t0), 1e3 * e[1].elapsed_time(e[2]), 1e6 * (t2 - t1)))
