"""One rank of the multi-GPU test of astro_stats_allreduce (tests/test_gpu_parity.py): every rank runs its own shard for a
few launches, then the counters are summed over the ranks through peer memory and compared with the NCCL all-reduce."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from astro_b200 import core  # noqa: E402
from astro_b200 import _native as nat  # noqa: E402
from astro_b200.batched import BatchedGames  # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    n = 4096
    games = BatchedGames(core.DEFAULT_CONFIG, n, bullet_cap=32, precision=32, device=local, seed=9, first_game=rank * n)
    games.set_reset_pool_on_device(256)
    games.reset_all()
    ok = games.stats_peer_init(dist)
    rounds = []
    if ok:
        for k in range(6):
            games.step_many(20 + 3 * rank + k, None, auto_reset=True)     # ranks arrive at different times
            want = games.stats_tensor(clear=False).clone()
            dist.all_reduce(want, op=dist.ReduceOp.SUM)
            got = games.stats_allreduce(clear=(k % 2 == 1)).clone()
            rounds.append(bool((want == got).all()) and int(got[nat.STAT_NAMES.index('env_steps')]) > 0)
        after_clear = games.stats()
        rounds.append(after_clear['env_steps'] == 0)
    with open(os.path.join(os.environ['ASTRO_PEER_OUT'], 'rank%d.json' % rank), 'w') as f:      # (the ranks' stdout interleaves)
        json.dump(dict(rank=rank, ok=ok, error=games.peer_error, rounds=rounds), f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
