"""One rank of the world_size-2 gloo test (tests/test_host_cpu.py): bench.py's sharding and
reduction helpers with the oracle standing in for the device tick (no GPU here)."""
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from astro_b200 import core, rng  # noqa: E402
from astro_b200.pool import make_pool  # noqa: E402
from oracle import astro_oracle as ao  # noqa: E402


def rollout(first_game, n, ticks, pool, seed=3):
    b = ao.Batch(n, 2, 32)
    pick = rng.pool_pick(seed, first_game + np.arange(n), np.zeros(n, dtype=np.uint32), pool['ships'].shape[0])
    b.ships[:], b.planets[:], b.np_[:] = pool['ships'][pick], pool['planets'][pick], pool['np'][pick]
    st = ao.rollout(core.DEFAULT_CONFIG, b, pool, seed, first_game, 0, ticks)
    return b, st


def digest(parts):
    h = hashlib.sha256()
    for p in parts:
        h.update(np.ascontiguousarray(p).tobytes())
    return h.hexdigest()


def main():
    dist.init_process_group('gloo')
    rank, world = dist.get_rank(), dist.get_world_size()
    plan = bench.shard_plan(world, rank, 256)
    pool = make_pool(core.DEFAULT_CONFIG, 64)
    t0 = time.perf_counter()
    b, st = rollout(plan['first_game'], plan['n_games'], 50, pool)
    my_ms = 1e3 * (time.perf_counter() - t0)
    total = bench.reduce_stats(torch.from_numpy(st.copy()), dist)
    max_ms = bench.reduce_max(my_ms, 'cpu', dist)
    # shard-independence: the two shards together equal one process running all 512 games
    gathered = [torch.zeros(plan['n_games'], 2, 5, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(b.ships.copy()))
    nbs = [torch.zeros(plan['n_games'], dtype=torch.int32) for _ in range(world)]
    dist.all_gather(nbs, torch.from_numpy(b.nb.copy()))
    sharded = digest([torch.cat(gathered).numpy(), torch.cat(nbs).numpy()])
    full, _ = rollout(0, plan['total'], 50, pool)
    single = digest([full.ships, full.nb])
    names = ('episodes', 'wins0', 'wins1', 'both_lost', 'timeouts', 'env_steps', 'bullets_spawned', 'overflow')
    print(json.dumps(dict(rank=rank, first_game=plan['first_game'], total=dict(zip(names, total.tolist())),
                          my_ms=my_ms, max_ms=max_ms, sharded_digest=sharded, single_process_digest=single)))
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
