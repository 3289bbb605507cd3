#!/usr/bin/env python
"""Is the tick's throughput limit per SM or device-wide?  Time the tick with n SMs taken away by a hog kernel
(tools/ubench/sm_hog.cu, bounded to 60 ms): per-SM limit -> time scales with 148 / (148 - n); device-wide -> it does not."""
import ctypes, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from astro_b200 import core
from astro_b200 import _native as nat
from astro_b200.batched import BatchedGames
from astro_b200.pool import make_pool
H = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ubench', 'libsm_hog.so'))
H.hog_start.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
games = BatchedGames(core.DEFAULT_CONFIG, 1 << 20, bullet_cap=32, precision=32, device=0)
pool = make_pool(core.DEFAULT_CONFIG, 4096)
games.set_reset_pool_arrays(pool['ships'], pool['planets'], pool['np']); games.reset_all()
flags = nat.TICK_AUTO_RESET
ring = torch.randint(0, 6, (8, games.n_pad, 2), dtype=torch.uint8).cuda()
for k in range(600): games.step_raw(ring[k % 8].data_ptr(), flags)
side = torch.cuda.Stream()
for n_hog in (0, 37, 74, 111):
    torch.cuda.synchronize()
    if n_hog:
        assert H.hog_start(ctypes.c_void_p(side.cuda_stream), n_hog, 60) == 0
        time.sleep(0.003)     # let the hog CTAs land
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(100): games.step_raw(ring[k % 8].data_ptr(), flags)
    e1.record(); e1.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / 100
    torch.cuda.synchronize()
    print(json.dumps(dict(sms_hogged=n_hog, sms_left=148 - n_hog, us_per_tick=round(us, 1), us_times_fraction_left=round(us * (148 - n_hog) / 148, 1))))
