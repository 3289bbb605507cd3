"""Cost of fresh-game mode (pool-free re-creation) against the reset pool: us per tick at 1M games."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from astro_b200 import core
from astro_b200.batched import BatchedGames
from astro_b200.pool import make_pool
cfg, N = core.DEFAULT_CONFIG, 1 << 20
pool = make_pool(cfg, 4096)
for mode in ("pool", "fresh"):
    g = BatchedGames(cfg, N, bullet_cap=32, precision=32, seed=0)
    if mode == "pool":
        g.set_reset_pool_arrays(pool["ships"], pool["planets"], pool["np"])
    else:
        g.enable_fresh_games(quota=48)
    g.reset_all()
    for _ in range(30):
        g.step_many(20, None, auto_reset=True)
    for fuse in (20, 64, 1):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 640 if fuse > 1 else 240
        e0.record()
        for _ in range(n // fuse):
            g.step_many(fuse, None, auto_reset=True)
        e1.record()
        torch.cuda.synchronize()
        print(mode, "ticks/launch", fuse, "us/tick %.2f" % (1e3 * e0.elapsed_time(e1) / (n // fuse * fuse)), flush=True)
    print(mode, 'awaiting', g.stats()["awaiting"])
