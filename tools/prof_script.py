import sys, os
sys.path.insert(0, '/root/repo')
import torch
from astro_b200 import core
from astro_b200.batched import BatchedGames
g = BatchedGames(core.DEFAULT_CONFIG, 1 << 18, bullet_cap=32, precision=32, device=0)
g.set_reset_pool_on_device(4096); g.reset_all()
g.rollout_device(300, bots=('script', 'script'))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a = g.script_controls()
torch.cuda.synchronize()
e0.record()
for _ in range(200): g.script_controls(out=a)
e1.record(); torch.cuda.synchronize()
print('script_kernel us per launch (262144 games):', 1e3 * e0.elapsed_time(e1) / 200)
