import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from astro_b200 import core
from astro_b200.batched import BatchedGames
from astro_b200.pool import make_pool
mode = sys.argv[1]
cfg, N = core.DEFAULT_CONFIG, 1 << 20
g = BatchedGames(cfg, N, bullet_cap=32, precision=32, seed=0)
if mode == 'pool':
    pool = make_pool(cfg, 4096)
    g.set_reset_pool_arrays(pool["ships"], pool["planets"], pool["np"])
else:
    g.enable_fresh_games(quota=48)
g.reset_all()
for _ in range(15):
    g.step_many(20, None, auto_reset=True)
for _ in range(60):
    g.step_many(1, None, auto_reset=True)
torch.cuda.synchronize()
