#!/usr/bin/env python
"""Bot-driven game loops on the device (BASELINE.json configs[0] and [4] shapes): N games, every tick
bot kernels -> tick kernel, no host between ticks (astro_rollout_device).  One JSON line per mode, with
the CPU reference-port rate of the same loop (oracle ScriptBot + step, one core) for the scripted mode."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from astro_b200 import core, rl
from astro_b200.batched import BatchedGames
ap = argparse.ArgumentParser()
ap.add_argument('--games', type=int, default=16384)
ap.add_argument('--ticks', type=int, default=1000)
ap.add_argument('--cpu-games', type=int, default=256)
ap.add_argument('--cpu-ticks', type=int, default=300)
args = ap.parse_args()
cfg = core.DEFAULT_CONFIG
torch.manual_seed(0)
net = rl.ValueNetwork(solo=False, nout=6).cuda().eval()
for bots in (('script', 'script'), ('policy', 'script'), ('policy', 'policy'), ('stream', 'stream')):
    games = BatchedGames(cfg, args.games, bullet_cap=32, precision=32, device=0)
    games.set_reset_pool_on_device(4096)
    games.reset_all()
    games.set_policy(net)
    games.rollout_device(200, bots=bots)
    games.stats(clear=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    games.rollout_device(args.ticks, bots=bots)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    st = games.stats()
    line = dict(bots='%s vs %s' % bots, games=args.games, ticks=args.ticks, env_steps_per_s=st['env_steps'] / (ms * 1e-3),
                us_per_tick=1e3 * ms / args.ticks, episodes=st['episodes'], wins0=st['wins0'], wins1=st['wins1'],
                both_lost=st['both_lost'], timeouts=st['timeouts'])
    if bots == ('script', 'script'):
        # CPU: the oracle's ScriptBot + step for the same loop, one core (the unmodified Python reference
        # does 2.4 k ticks/s/core on this loop)
        from oracle import astro_oracle as ao
        from astro_b200.pool import make_pool
        pool = make_pool(cfg, 256)
        n = args.cpu_games
        b = ao.Batch(n, 2, 64)
        b.ships[:], b.planets[:], b.np_[:] = pool['ships'][:n], pool['planets'][:n], pool['np'][:n]
        t0 = time.perf_counter()
        steps = 0
        for k in range(args.cpu_ticks):
            ctl = ao.script_batch(cfg, b)
            b, rew, done, ev = ao.step_batch(cfg, b, ctl)
            steps += n
            idx = np.nonzero(done)[0]
            if len(idx):
                pick = (idx * 7 + k) % 256
                b.ships[idx], b.planets[idx], b.np_[idx], b.nb[idx] = pool['ships'][pick], pool['planets'][pick], pool['np'][pick], 0
                b.reload[idx], b.t[idx] = 0.0, 0.0
        line['cpu_port_env_steps_per_s_1core'] = steps / (time.perf_counter() - t0)
    print(json.dumps(line), flush=True)
