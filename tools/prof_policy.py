"""One policy_controls launch on 16,384 games in the stationary population (ncu target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from astro_b200 import core, rl
from astro_b200.batched import BatchedGames
g = BatchedGames(core.DEFAULT_CONFIG, 16384, bullet_cap=32, precision=32, device=0)
g.set_reset_pool_on_device(4096)
g.reset_all()
g.step_many(300, None, auto_reset=True)
torch.manual_seed(0)
g.set_policy(rl.ValueNetwork(solo=False, nout=6).cuda())
out = torch.full((g.n_pad, 2), 2, dtype=torch.uint8, device='cuda')
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    g.policy_controls(out=out)
torch.cuda.synchronize()
e0.record()
for _ in range(20):
    g.policy_controls(out=out)
e1.record()
torch.cuda.synchronize()
print('policy_controls %.2f us per launch (16,384 games)' % (1e3 * e0.elapsed_time(e1) / 20))
