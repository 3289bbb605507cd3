#!/usr/bin/env python
"""Phase timeline of the tick kernel (needs a library built with -DASTRO_TIMELINE, passed via ASTRO_B200_LIB):
every warp stamps clock64() at its phase boundaries; prints mean / percentiles of each phase in SM cycles."""
import argparse, ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from astro_b200 import core
from astro_b200 import _native as nat
from astro_b200.batched import BatchedGames
from astro_b200.pool import make_pool
ap = argparse.ArgumentParser()
ap.add_argument('--games', type=int, default=1 << 20)
ap.add_argument('--preroll', type=int, default=600)
ap.add_argument('--flags', type=int, default=0)
a = ap.parse_args()
games = BatchedGames(core.DEFAULT_CONFIG, a.games, bullet_cap=32, precision=32, device=0)
pool = make_pool(core.DEFAULT_CONFIG, 4096)
games.set_reset_pool_arrays(pool['ships'], pool['planets'], pool['np']); games.reset_all()
flags = nat.TICK_AUTO_RESET | a.flags
ring = torch.randint(0, 6, (8, games.n_pad, 2), dtype=torch.uint8).cuda()
for k in range(a.preroll): games.step_raw(ring[k % 8].data_ptr(), flags)
buf = torch.zeros((a.games // 32, 8), dtype=torch.int64, device='cuda')
L = nat.lib()
L.astro_debug_set_timeline.argtypes = [ctypes.c_void_p]
assert L.astro_debug_set_timeline(buf.data_ptr()) == 0
torch.cuda.synchronize()
games.step_raw(ring[0].data_ptr(), flags)
torch.cuda.synchronize()
t = buf.cpu().numpy().astype(np.float64)
names = ['start->meta', 'meta->staged(prefix+cp.async issue)', 'staged->ships/planets arrived', 'physics+stores', 'wait bullets', 'bullet loop', 'terminal/spawn/stats']
d = np.diff(t, axis=1)
life = t[:, 7] - t[:, 0]
print('tiles %d   lifetime mean %.0f  p10 %.0f  p50 %.0f  p90 %.0f cycles' % (len(t), life.mean(), *np.percentile(life, [10, 50, 90])))
for i, n in enumerate(names):
    print('%-40s mean %7.0f  p10 %7.0f  p50 %7.0f  p90 %7.0f   %4.1f%%' % (n, d[:, i].mean(), *np.percentile(d[:, i], [10, 50, 90]), 100 * d[:, i].mean() / life.mean()))
