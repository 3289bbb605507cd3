#!/usr/bin/env python
"""observe_kernel alone: get_features_batch + roll_ships + to_batch for every game, both perspectives.
Output [N, 2, 36, 15] f32 (4,320 B/game) is written to HBM; N is chosen so it exceeds L2."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from astro_b200 import core
from astro_b200 import _native as nat
from astro_b200.batched import BatchedGames
from astro_b200.pool import make_pool
ap = argparse.ArgumentParser()
ap.add_argument('--games', type=int, default=1 << 18)
ap.add_argument('--reps', type=int, default=50)
args = ap.parse_args()
n = args.games
games = BatchedGames(core.DEFAULT_CONFIG, n, bullet_cap=32, precision=32, device=0)
pool = make_pool(core.DEFAULT_CONFIG, 4096)
games.set_reset_pool_arrays(pool['ships'], pool['planets'], pool['np']); games.reset_all()
for _ in range(700): games.step_raw(0, nat.TICK_AUTO_RESET)
st = games.stats(clear=True)
games.step_raw(0, nat.TICK_AUTO_RESET)
st = games.stats(clear=True)
obs = torch.empty((games.n_pad, 2, 36, 15), dtype=torch.float32, device='cuda')
for _ in range(5): games.observe(out=obs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.reps): games.observe(out=obs)
e1.record(); torch.cuda.synchronize()
us = 1e3 * e0.elapsed_time(e1) / args.reps
wr = n * 2 * 36 * 15 * 4
rd = n * (4 + 40) + 16 * (st['planets_live'] + st['bullets_in'])
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists('MEASURED_PEAKS.json') else 6548.8
print(json.dumps(dict(kernel='observe_kernel<float,2>', games=n, us_per_launch=us, bytes_written=wr, bytes_read=rd,
                      achieved_GBps=(wr + rd) / (us * 1e-6) / 1e9, frac_of_measured_peak=(wr + rd) / (us * 1e-6) / 1e9 / peak,
                      observations_per_s=2 * n / (us * 1e-6))))
