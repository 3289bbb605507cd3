#!/usr/bin/env python
"""Small workload for compute-sanitizer: every kernel of the library on a few tiles."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from astro_b200 import core
from astro_b200 import _native as nat
from astro_b200.batched import BatchedGames
from astro_b200.pool import make_pool
pool = make_pool(core.DEFAULT_CONFIG, 64)
for prec, flags in ((32, 0), (32, nat.TICK_GENERIC_KERNEL), (64, 0)):
    g = BatchedGames(core.DEFAULT_CONFIG, 256, bullet_cap=32, precision=prec, device=0, seed=1)
    g.set_reset_pool_arrays(pool['ships'], pool['planets'], pool['np'])
    g.reset_all()
    g.tick_flags = flags
    for k in range(60):
        g.step(None, auto_reset=(k % 7 != 0))
        if k % 20 == 19:
            g.reset_done()
            g.observe()
    print(prec, flags, g.stats())
torch.cuda.synchronize()
print('sanitize case done')
