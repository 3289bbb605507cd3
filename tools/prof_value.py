"""astro_value_forward on the observation of 16,384 games (both perspectives) in the stationary population: time per launch
against the PyTorch forward of the same tensor (ncu target)."""
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
net = rl.ValueNetwork(solo=False, nout=6).cuda()
obs = g.observe()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    for name, f in (('astro_value_forward', net), ('torch forward', net.forward_torch)):
        for _ in range(3):
            q = f(obs)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            q = f(obs)
        e1.record()
        torch.cuda.synchronize()
        print('%s %.2f us per call (16,384 games x 2 views x 36 rows)' % (name, 1e3 * e0.elapsed_time(e1) / 20))
    print('max |fused - torch| = %.3g' % float((net(obs) - net.forward_torch(obs)).abs().max()))
