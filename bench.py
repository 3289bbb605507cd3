#!/usr/bin/env python
"""bench.py — env-steps/s of the batched Astro tick on N B200s, as a fraction of the HBM roofline,
next to the CPU step loop timed on the same box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one game tick (astro.core.step) of every game of the batch.  Workload (BASELINE.json
configs[3] shape, weak scaling): 1,048,576 duel games PER GPU, default planets (max_planets=4),
uniform random controls, bullet pool K=32, finished games re-created from a 4,096-state pool built
by core.create; the population is pre-rolled to its stationary bullet count before timing.  The
state (0.67 GB per GPU) and the controls exceed the 126 MB L2, so no flush is needed between steps.

Numbers on the JSON line:
  value     whole-job env-steps/s, controls already resident in HBM (a ring of pre-generated
            [games, 2] u8 arrays, a different one every tick), every tick's events written to HBM;
            CUDA events, max over ranks.  The ticks go through astro_tick_many, --fuse ticks per
            launch: games do not interact, so the kernel runs the ticks of a tile back to back and
            the state travels from one tick to the next through L2.  `per_tick_launch` is the same
            loop as one launch per tick (what a policy in the loop needs).
  e2e       the same through BatchedGames.step_host(): pinned HOST controls in, events out,
            copies inside the timed region.
  roofline  algorithmic bytes of the tick kernel per launch (from the device counters of the timed
            region) / average launch duration, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the oracle port of the reference loop on all host cores (rank 0, N=1 only).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'env-steps/sec'
UNIT = 'env-steps/s'
FALLBACK_HBM_GBS = 6650.0   # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


# --------------------------------------------------------------------------- multi-rank helpers
def shard_plan(world, rank, games_per_gpu):
    """Env-parallel sharding: rank r owns the contiguous global games [r*n, (r+1)*n)."""
    return dict(first_game=rank * games_per_gpu, n_games=games_per_gpu, total=world * games_per_gpu)


def reduce_stats(stats, dist=None):
    """Sum the int64 counter vector over ranks (the only collective of the path)."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def reduce_max(value, device, dist=None):
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def algorithmic_bytes(st, S, actions_in_hbm=True, with_reward=False):
    """Bytes the tick must move per the layout in include/astro_b200.h: every live byte of state
    read once and written once, controls read, events written.  st = device counters."""
    per_step = 2 * (4 + 16 * S + 4 * S) + (S if actions_in_hbm else 0) + 1 + ((4 * S + 1) if with_reward else 0)
    return (st['env_steps'] * per_step + 32 * st['planets_live'] + 16 * (st['bullets_in'] + st['bullets_out']))


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the GPU is under load."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {}
        for n in dir(nv):
            if n.startswith('nvmlClocksThrottleReason') or n.startswith('nvmlClocksEventReason'):
                v = getattr(nv, n)
                if isinstance(v, int) and v:
                    names[v] = n.replace('nvmlClocksThrottleReason', '').replace('nvmlClocksEventReason', '')
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit and bit & (bit - 1) == 0:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join()
        idle = {'GpuIdle', 'None', 'ApplicationsClocksSetting'}
        return dict(sm_mhz=(float(np.median(self.samples)) if self.samples else None), sm_max_mhz=self.max_mhz,
                    reasons=sorted(r for r in self.reasons if r not in idle), samples=len(self.samples))


# --------------------------------------------------------------------------- CPU baseline (oracle)
def cpu_rollout_rate(n_games, warm_ticks, ticks, threads, seed=0, pool_size=1024):
    """Times the oracle port of the reference step loop (same workload: duel, default planets,
    counter-stream random controls, auto-reset from a create() pool)."""
    from astro_b200 import core, rng
    from oracle import astro_oracle as ao
    from astro_b200.pool import make_pool
    cfg = core.DEFAULT_CONFIG
    pool = make_pool(cfg, pool_size)
    b = ao.Batch(n_games, 2, 32)
    pick = rng.pool_pick(seed, np.arange(n_games), np.zeros(n_games, dtype=np.uint32), pool_size)
    b.ships[:], b.planets[:], b.np_[:] = pool['ships'][pick], pool['planets'][pick], pool['np'][pick]
    ao.rollout(cfg, b, pool, seed, 0, 0, warm_ticks, threads=threads)
    t0 = time.perf_counter()
    st = ao.rollout(cfg, b, pool, seed, 0, warm_ticks, ticks, threads=threads)
    dt = time.perf_counter() - t0
    return float(st[5]) / dt, dt, int(st[5])


def run_reference(args):
    """--impl reference: the reference's CPU step loop (oracle port; the Python reference cannot
    travel to the GPU box) on all host cores, same metric/config as our arm."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_games = args.ref_games
    rate, dt, steps = cpu_rollout_rate(n_games, max(args.warmup, 3), args.steps, threads)
    sample = '%d games x %d ticks (of the %d-games-per-GPU workload), oracle C port of astro/core.py step' % (
        n_games, args.steps, args.games_per_gpu)
    line = dict(metric=METRIC, value=rate, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=max(args.warmup, 3),
                ms_per_step=1e3 * dt / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='f64', data='synthetic', impl='reference', config=workload_config(args),
                cpu_baseline=dict(value=rate, unit=UNIT, cores=threads, kind='port', sample=sample),
                e2e=dict(value=rate, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


def workload_config(args):
    return dict(workload='BASELINE configs[3] shape, weak-scaled: %d duel games per GPU, default planets '
                         '(max_planets=4), uniform random controls, auto-reset from a %d-state create() pool'
                         % (args.games_per_gpu, args.pool),
                games_per_gpu=args.games_per_gpu, bullet_cap=args.bullet_cap, reset_pool=args.pool,
                preroll_ticks=args.preroll, state_precision='fp32 state, fp64-exact predicates',
                ticks_per_launch=args.fuse,
                l2_policy='state (0.67 GB/GPU) and controls exceed the 126 MB L2; no flush between steps; inside a launch the '
                          'ticks of a tile run back to back, so a tile\'s state deliberately stays in L2 from one tick to the next',
                parallelism='env-parallel shards, one process per GPU, NCCL only for the stats reduce')


def bind_to_gpu_numa_node(index):
    """Pins this rank to the CPUs NVML lists as local to its GPU, BEFORE the pinned host buffers are allocated
    (first touch puts them on that NUMA node): the end-to-end leg moves 3 MB per tick and GPU through host memory,
    and with several ranks per box a buffer on the far socket costs every copy a trip over the socket link."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1 and 64 * w + b < n_cpu}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return sorted(cpus)
    except Exception:
        pass
    return None


# --------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from astro_b200 import core
    from astro_b200 import _native as nat
    from astro_b200.batched import BatchedGames
    from astro_b200.pool import make_pool

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        raise SystemExit('--gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run for N>1)' % (args.gpus, world))
    torch.cuda.set_device(local)
    bind_to_gpu_numa_node(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    plan = shard_plan(world, rank, args.games_per_gpu)
    n, S, K = plan['n_games'], 2, args.bullet_cap
    cfg = core.DEFAULT_CONFIG

    pool = make_pool(cfg, args.pool)     # host: core.create over generate_configs (seed 42)
    games = BatchedGames(cfg, n, bullet_cap=K, precision=32, device=local, seed=args.seed, first_game=plan['first_game'])
    games.set_reset_pool_arrays(pool['ships'], pool['planets'], pool['np'])
    games.reset_all()
    flags = nat.TICK_AUTO_RESET | args.tick_flags
    for _ in range(args.preroll):        # reach the stationary population (device counter-stream controls)
        games.step_raw(0, flags)
    torch.cuda.synchronize()
    flags |= args.timed_flags            # (experiment builds: bits that only apply after the pre-roll)

    # ring of control arrays resident in HBM / in pinned host memory: R different arrays, one per tick
    R = max(8, args.fuse)
    gen = torch.Generator(device='cpu').manual_seed(1234 + rank)
    host_ring = torch.randint(0, 6, (R, games.n_pad, S), dtype=torch.uint8, generator=gen).pin_memory()
    dev_ring = host_ring.to(dev)
    ptrs = [dev_ring[i].data_ptr() for i in range(R)]
    events_host = torch.empty(games.n_pad, dtype=torch.uint8).pin_memory()
    events_dev = torch.empty((R, games.n_pad), dtype=torch.uint8, device=dev)

    def run_steps(k_steps):
        """k_steps ticks, args.fuse per launch, tick k reading control array k % R; returns the launches."""
        done, n_launch = 0, 0
        while done < k_steps:
            r0 = done % R
            t = min(args.fuse, k_steps - done, R - r0)
            games.step_many_raw(ptrs[r0], events_dev[r0].data_ptr(), t, flags)
            done += t
            n_launch += 1
        return n_launch

    sampler = ClockSampler(local)
    sampler.start()
    run_steps(args.warmup)
    games.stats_tensor(clear=True)
    launches0 = games.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    n_launches = run_steps(args.steps)
    st_t = games.stats_tensor(clear=True).clone()
    reduce_stats(st_t, dist)             # NCCL: the episode-statistics reduction, once per rollout
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    my_ms = e0.elapsed_time(e1)
    ms = reduce_max(my_ms, dev, dist)
    launches = games.launches - launches0
    total = dict(zip(nat.STAT_NAMES, (int(x) for x in st_t.cpu().numpy())))
    value = total['env_steps'] / (ms * 1e-3)

    # kernel-only roofline of the tick kernel on this rank: one launch per step, back to back
    games_stats_local = total if world == 1 else None
    if world > 1:
        # per-rank counters for the local roofline: re-measure a short local window
        games.stats_tensor(clear=True)
        torch.cuda.synchronize()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        run_steps(args.steps)
        r1.record()
        torch.cuda.synchronize()
        kern_ms = r0.elapsed_time(r1)
        games_stats_local = games.stats(clear=True)
    else:
        kern_ms = my_ms
    alg = algorithmic_bytes(games_stats_local, S)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', FALLBACK_HBM_GBS))
    achieved = alg / (kern_ms * 1e-3) / 1e9
    traffic = None
    try:
        prof = json.load(open(os.path.join(ROOT, 'profiles', 'tick_kernel_ncu_summary.json')))
        traffic = prof.get('dram_bytes_per_launch')
    except Exception:
        pass
    roofline = dict(bound='hbm', achieved=achieved, peak=peak, unit='GB/s', frac=achieved / peak, traffic=traffic,
                    kernel='tick_f32_kernel<2,true,%s>' % ('true' if args.fuse > 1 else 'false'),
                    peak_source='MEASURED_PEAKS.json hbm_gbs' if peaks else 'fallback',
                    ticks_per_launch=args.steps / n_launches,
                    algorithmic_bytes_per_launch=alg / n_launches,
                    bytes_per_env_step=alg / max(1, games_stats_local['env_steps']),
                    mean_planets=games_stats_local['planets_live'] / max(1, games_stats_local['env_steps']),
                    mean_bullets=games_stats_local['bullets_in'] / max(1, games_stats_local['env_steps']),
                    avg_launch_us=1e3 * kern_ms / n_launches,
                    note='algorithmic bytes = every live byte of state read and written once PER TICK; with several ticks '
                         'of a tile per launch most of that traffic stays in L2 (traffic = DRAM bytes per launch, ncu)')

    # the same loop as one launch per tick (a policy between the ticks needs this form)
    pt_steps = max(R, min(args.steps, 400))
    games.stats_tensor(clear=True)
    torch.cuda.synchronize()
    e0.record()
    for k in range(pt_steps):
        games.step_raw(ptrs[k % R], flags)
    e1.record()
    torch.cuda.synchronize()
    pt_ms = reduce_max(e0.elapsed_time(e1), dev, dist)
    pt_stats = games.stats(clear=True)
    per_tick = dict(value=world * n * pt_steps / (pt_ms * 1e-3), unit=UNIT, ms_per_step=pt_ms / pt_steps, steps=pt_steps,
                    frac_of_hbm_peak=algorithmic_bytes(pt_stats, S) / (pt_ms * 1e-3) / 1e9 / peak,
                    kernel='tick_f32_kernel<2,true,false>')

    # e2e: the public API with HOST buffers; every tick's controls are copied in from pinned host
    # memory and its events copied out, all inside the timed region (copies of neighbouring ticks
    # overlap the kernel: BatchedGames.rollout_host -> astro_rollout_host)
    # (a call covers E ticks: E different control arrays in pinned host memory, E event arrays back)
    E = max(R, (args.e2e_call // R) * R) if args.e2e_steps >= args.e2e_call else R
    e2e_steps = max(E, (max(3, min(args.steps, args.e2e_steps)) // E) * E)
    e2e_ring = host_ring if E == R else host_ring.repeat(E // R, 1, 1).pin_memory()
    events_ring = torch.empty((E, games.n_pad), dtype=torch.uint8).pin_memory()
    games.rollout_host(e2e_ring, events_ring, auto_reset=True)      # warm-up: E ticks
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for k in range(e2e_steps // E):
        games.rollout_host(e2e_ring, events_ring, auto_reset=True)
    e1.record()
    torch.cuda.synchronize()
    e2e_ms = reduce_max(e0.elapsed_time(e1), dev, dist)
    e2e_value = world * n * e2e_steps / (e2e_ms * 1e-3)
    # the unpipelined form (copy in, tick, copy out, synchronise, every tick) for reference
    t0 = time.perf_counter()
    for k in range(R):
        games.step_host(host_ring[k % R], events_host, auto_reset=True)
    sync_ms = 1e3 * (time.perf_counter() - t0)
    e2e_sync_value = world * n * R / (reduce_max(sync_ms, dev, dist) * 1e-3)
    clocks = sampler.stop()

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            n_cpu = 16384
            rate, _, _ = cpu_rollout_rate(n_cpu, 20, 40, threads)           # calibrate
            ticks = int(max(50, min(100000, args.cpu_seconds * rate / n_cpu)))
            rate, dt, steps = cpu_rollout_rate(n_cpu, 20, ticks, threads)
            cpu = dict(value=rate, unit=UNIT, cores=threads, kind='port',
                       sample='%d games x %d ticks of the same workload (%.1f s), oracle C port of astro/core.py '
                              'step with auto-reset, one thread per host core' % (n_cpu, ticks, dt))
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
                    dtype='f32', data='synthetic', config=workload_config(args), roofline=roofline,
                    cpu_baseline=cpu, clocks=clocks,
                    e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=games.n_pad * S,
                             d2h_bytes_per_step=games.n_pad, steps=e2e_steps,
                             api='BatchedGames.rollout_host -> astro_rollout_host (copies overlap the kernel)', ticks_per_call=E,
                             unpipelined_value=e2e_sync_value, unpipelined_api='BatchedGames.step_host -> astro_tick_host'),
                    gpu_launches=launches, per_tick_launch=per_tick,
                    episode_stats={k: total[k] for k in ('episodes', 'wins0', 'wins1', 'both_lost', 'timeouts', 'overflow')})
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=2000)
    ap.add_argument('--warmup', type=int, default=100)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--games-per-gpu', type=int, default=1 << 20)
    ap.add_argument('--bullet-cap', type=int, default=32)
    ap.add_argument('--pool', type=int, default=4096)
    ap.add_argument('--preroll', type=int, default=600)
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--e2e-steps', type=int, default=256)
    ap.add_argument('--e2e-call', type=int, default=128, help='ticks per rollout_host call of the e2e leg')
    ap.add_argument('--cpu-seconds', type=float, default=20.0)
    ap.add_argument('--ref-games', type=int, default=65536)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--fuse', type=int, default=64, help='ticks per launch of the timed loop (astro_tick_many); 1 = one launch per tick')
    ap.add_argument('--tick-flags', type=int, default=0, help='extra ASTRO_TICK_* bits (kernel A/B)')
    ap.add_argument('--timed-flags', type=int, default=0, help='extra tick bits after the pre-roll (experiment builds)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
