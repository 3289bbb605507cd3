#!/usr/bin/env python
"""BASELINE.json configs[4]: self-play rollout — 16,384 games x 1,000 ticks, every tick
observe() -> ValueNetwork.forward (both perspectives) -> greedy controls -> step(auto_reset).
Prints one JSON line (secondary measurement; bench.py is the headline)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from astro_b200 import core, rl
from astro_b200.batched import BatchedGames
from astro_b200.pool import make_pool

ap = argparse.ArgumentParser()
ap.add_argument('--games', type=int, default=16384)
ap.add_argument('--ticks', type=int, default=1000)
ap.add_argument('--bullet-cap', type=int, default=32)
ap.add_argument('--fused', action='store_true', help='features + ValueNetwork + argmax in one CUDA kernel (astro_policy_controls)')
ap.add_argument('--shared', action='store_true', help='one observation tensor for both ships (observe(shared=True) + forward_both)')
args = ap.parse_args()
torch.manual_seed(0)
dev = torch.device('cuda', 0)
cfg = core.DEFAULT_CONFIG
games = BatchedGames(cfg, args.games, bullet_cap=args.bullet_cap, precision=32, device=0, seed=0)
pool = make_pool(cfg, 4096)
games.set_reset_pool_arrays(pool['ships'], pool['planets'], pool['np'])
games.reset_all()
net = rl.ValueNetwork(solo=False, nout=6).to(dev).eval()
obs = torch.empty((games.n_pad, 36, 15) if args.shared else (games.n_pad, 2, 36, 15), dtype=torch.float32, device=dev)


games.set_policy(net)
act = torch.full((games.n_pad, 2), 2, dtype=torch.uint8, device=dev)


def tick():
    if args.fused:
        games.policy_controls(out=act)
        games.step(act, auto_reset=True, want_reward=False)
        return
    with torch.no_grad():
        o = games.observe(out=obs, shared=args.shared)   # [N, 2, 36, 15] / [N, 36, 15]
        q = net.forward_both(o) if args.shared else net(o)   # [N, 2, 6]
        a = q.argmax(-1).to(torch.uint8)                 # greedy controls for both ships
    games.step(a, auto_reset=True, want_reward=False)


for _ in range(200):
    tick()
games.stats(clear=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
# observe / tick kernels alone (same states), for the split
t_obs0, t_obs1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t_obs0.record()
for _ in range(50):
    games.observe(out=obs, shared=args.shared)
t_obs1.record()
t_pol0, t_pol1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t_pol0.record()
for _ in range(50):
    games.policy_controls(out=act)
t_pol1.record()
t_tk0, t_tk1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
t_tk0.record()
for _ in range(50):
    games.step_raw(act.data_ptr(), 1)
t_tk1.record()
e0.record()
for _ in range(args.ticks):
    tick()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
st = games.stats()
obs_us = 1e3 * t_obs0.elapsed_time(t_obs1) / 50
obs_bytes = obs.numel() * 4
print(json.dumps(dict(fused=args.fused, shared=args.shared, workload='configs[4]: %d games x %d ticks, observe -> ValueNetwork(6) -> greedy -> step' % (args.games, args.ticks),
                      env_steps_per_s=st['env_steps'] / (ms * 1e-3), ms_per_tick=ms / args.ticks,
                      observe_us=obs_us, policy_kernel_us=1e3 * t_pol0.elapsed_time(t_pol1) / 50, tick_kernel_us=1e3 * t_tk0.elapsed_time(t_tk1) / 50, observe_write_GBps=obs_bytes / (obs_us * 1e-6) / 1e9,
                      episodes=st['episodes'], wins0=st['wins0'], wins1=st['wins1'], both_lost=st['both_lost'],
                      timeouts=st['timeouts'], overflow=st['overflow'])))
