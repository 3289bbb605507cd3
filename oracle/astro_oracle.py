"""ctypes loader for the CPU oracle (oracle/astro_oracle.c) — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product package (astro_b200/) never does.
"""
import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libastro_oracle.so')

MAXP = 4
EV_HIT0, EV_HIT1, EV_TIMEOUT, EV_FIRED, EV_OVERFLOW = 1, 2, 4, 8, 16


class Config(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        'gravity', 'dt', 'max_time', 'reload_time', 'bullet_speed', 'ship_thrust', 'ship_rspeed',
        'ship_radius', 'planet_mass', 'planet_radius')] + [('solo', C.c_int32), ('reserved', C.c_int32)]

    @classmethod
    def from_any(cls, cfg):
        """cfg: mapping or namedtuple with the reference Config's world fields (core.py:20-41)."""
        get = (lambda k: cfg[k]) if isinstance(cfg, dict) else (lambda k: getattr(cfg, k))
        c = cls()
        for n, _ in cls._fields_[:10]:
            setattr(c, n, float(get(n)))
        c.solo = int(bool(get('solo')))
        return c


def build(force=False):
    src = os.path.join(HERE, 'astro_oracle.c')
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', HERE, '-s'])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        vp, i32, i64, u32, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_double
        L.ao_sincos_f32.argtypes = [vp, vp, vp, i64]
        L.ao_wrap_unit_square.argtypes = [vp, vp, i64]
        L.ao_norm_angle.argtypes = [vp, vp, i64]
        L.ao_collisions.argtypes = [vp, vp, i32, vp]
        L.ao_step_one.argtypes = [C.POINTER(Config), i32, i32, i32, i32, vp, vp, vp, f64, f64, vp,
                                  vp, vp, vp, vp, vp, vp, vp, vp]
        L.ao_step_one.restype = i32
        L.ao_step_one_raw.argtypes = L.ao_step_one.argtypes
        L.ao_step_one_raw.restype = i32
        L.ao_step_batch.argtypes = [C.POINTER(Config), i64, i32, i32] + [vp] * 18 + [i32]
        L.ao_features.argtypes = [i32, i32, i32, vp, vp, vp, i32, i32, vp]
        L.ao_features.restype = i32
        L.ao_script_control.argtypes = [C.POINTER(Config), i32, i32, vp, vp, i32, f64, f64]
        L.ao_script_control.restype = i32
        L.ao_script_batch.argtypes = [C.POINTER(Config), i64, i32, vp, vp, vp, f64, f64, vp]
        L.ao_rollout.argtypes = [C.POINTER(Config), i64, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp,
                                 i64, vp, vp, vp, u32, i64, u32, i32, i32, vp]
        L.ao_rollout.restype = i64
        L.ao_actions.argtypes = [u32, i64, i64, u32, i32, vp]
        L.ao_pool_pick.argtypes = [u32, i64, i64, vp, u32, vp]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def sincos_f32(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    s, c = np.empty_like(x), np.empty_like(x)
    lib().ao_sincos_f32(_p(x), _p(s), _p(c), x.size)
    return s, c


def wrap_unit_square(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    o = np.empty_like(x)
    lib().ao_wrap_unit_square(_p(x), _p(o), x.size)
    return o


def norm_angle(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    o = np.empty_like(x)
    lib().ao_norm_angle(_p(x), _p(o), x.size)
    return o


def collisions(x, r):
    x = np.ascontiguousarray(x, dtype=np.float64)
    r = np.ascontiguousarray(r, dtype=np.float64)
    hit = np.zeros(r.shape[0], dtype=np.uint8)
    lib().ao_collisions(_p(x), _p(r), r.shape[0], _p(hit))
    return hit.astype(bool)


def step_one(cfg, ships, planets, bullets, reload, t, control, bullet_cap=-1, raw=False):
    """One game, one tick.  Returns dict(done, reward, events[, ships, planets, bullets, reload, t]).
    raw=True: the state is what core.create returned (float32 arrays): the reference's first-tick arithmetic."""
    c = cfg if isinstance(cfg, Config) else Config.from_any(cfg)
    ships = np.ascontiguousarray(ships, dtype=np.float64).reshape(-1, 5)
    planets = np.ascontiguousarray(planets, dtype=np.float64).reshape(-1, 4)
    bullets = np.ascontiguousarray(bullets, dtype=np.float64).reshape(-1, 4)
    control = np.ascontiguousarray(control, dtype=np.int64)
    S, P, B = ships.shape[0], planets.shape[0], bullets.shape[0]
    so, po, bo = np.empty_like(ships), np.empty_like(planets), np.empty((B + S, 4))
    nb, ro, to = C.c_int32(0), C.c_double(0), C.c_double(0)
    rew, ev = np.zeros(S), C.c_int32(0)
    term = (lib().ao_step_one_raw if raw else lib().ao_step_one)(C.byref(c), S, P, B, bullet_cap, _p(ships), _p(planets), _p(bullets),
                             float(reload), float(t), _p(control), _p(so), _p(po), _p(bo),
                             C.byref(nb), C.byref(ro), C.byref(to), _p(rew), C.byref(ev))
    out = dict(done=bool(term), reward=rew, events=ev.value)
    if not term:
        out.update(ships=so, planets=po, bullets=bo[:nb.value].copy(), reload=ro.value, t=to.value)
    return out


class Batch:
    """Host image of a batch of games in the oracle's fixed-stride float64 layout."""

    def __init__(self, n, S, K):
        self.n, self.S, self.K = n, S, K
        self.ships = np.zeros((n, S, 5))
        self.planets = np.zeros((n, MAXP, 4))
        self.np_ = np.zeros(n, dtype=np.int32)
        self.bullets = np.zeros((n, K, 4))
        self.nb = np.zeros(n, dtype=np.int32)
        self.reload = np.zeros(n)
        self.t = np.zeros(n)
        self.episode = np.zeros(n, dtype=np.uint32)

    def copy(self):
        b = Batch.__new__(Batch)
        b.n, b.S, b.K = self.n, self.S, self.K
        for k in ('ships', 'planets', 'np_', 'bullets', 'nb', 'reload', 't', 'episode'):
            setattr(b, k, getattr(self, k).copy())
        return b


def step_batch(cfg, b, control, alive=None, nthreads=1):
    """All games one tick (no reset).  Returns (next Batch, reward [n,S], done [n], events [n])."""
    c = cfg if isinstance(cfg, Config) else Config.from_any(cfg)
    o = b.copy()
    control = np.ascontiguousarray(control, dtype=np.int64)
    alive = None if alive is None else np.ascontiguousarray(alive, dtype=np.uint8)
    reward = np.zeros((b.n, b.S))
    done = np.zeros(b.n, dtype=np.uint8)
    events = np.zeros(b.n, dtype=np.int32)
    lib().ao_step_batch(C.byref(c), b.n, b.S, b.K, _p(b.ships), _p(b.planets), _p(b.np_), _p(b.bullets),
                        _p(b.nb), _p(b.reload), _p(b.t), _p(control), _p(alive), _p(o.ships), _p(o.planets),
                        _p(o.bullets), _p(o.nb), _p(o.reload), _p(o.t), _p(reward), _p(done), _p(events),
                        nthreads)
    return o, reward, done, events


def features(S, ships, planets, bullets, me, n_rows):
    ships = np.ascontiguousarray(ships, dtype=np.float64).reshape(-1, 5)
    planets = np.ascontiguousarray(planets, dtype=np.float64).reshape(-1, 4)
    bullets = np.ascontiguousarray(bullets, dtype=np.float64).reshape(-1, 4)
    out = np.empty((n_rows, 1 + 5 * S + 4), dtype=np.float32)
    r = lib().ao_features(S, planets.shape[0], bullets.shape[0], _p(ships), _p(planets), _p(bullets),
                          me, n_rows, _p(out))
    if r < 0:
        raise ValueError('n_rows too small')
    return out


SCRIPT_ARGS = dict(avoid_distance=0.1, avoid_threshold=0.45)   # script.py:16-20


def script_control(cfg, ships, planets, me, avoid_distance=0.1, avoid_threshold=0.45):
    c = cfg if isinstance(cfg, Config) else Config.from_any(cfg)
    ships = np.ascontiguousarray(ships, dtype=np.float64).reshape(-1, 5)
    planets = np.ascontiguousarray(planets, dtype=np.float64).reshape(-1, 4)
    return lib().ao_script_control(C.byref(c), ships.shape[0], planets.shape[0], _p(ships), _p(planets), me,
                                   avoid_distance, avoid_threshold)


def script_batch(cfg, b, avoid_distance=0.1, avoid_threshold=0.45):
    """ScriptBot controls for every ship of every game of a Batch -> int64 [n, S]."""
    c = cfg if isinstance(cfg, Config) else Config.from_any(cfg)
    out = np.zeros((b.n, b.S), dtype=np.int64)
    lib().ao_script_batch(C.byref(c), b.n, b.S, _p(b.ships), _p(b.planets), _p(b.np_), avoid_distance,
                          avoid_threshold, _p(out))
    return out


def rollout(cfg, b, pool, seed, first_game, step0, n_ticks, threads=1):
    """In-place rollout with counter-stream controls and auto-reset from `pool`
    (dict ships [M,S,5], planets [M,4,4], np [M]) — the CPU baseline loop.  Games are split
    into `threads` contiguous slices, one Python thread each (ctypes drops the GIL).
    Returns stats int64[8]: episodes, wins0, wins1, both_lost, timeouts, env_steps,
    bullets_spawned, overflow."""
    c = cfg if isinstance(cfg, Config) else Config.from_any(cfg)
    L = lib()
    M = 0 if pool is None else pool['ships'].shape[0]
    ps = None if pool is None else np.ascontiguousarray(pool['ships'], dtype=np.float64)
    pp = None if pool is None else np.ascontiguousarray(pool['planets'], dtype=np.float64)
    pn = None if pool is None else np.ascontiguousarray(pool['np'], dtype=np.int32)

    def run(lo, hi):
        st = np.zeros(8, dtype=np.int64)
        L.ao_rollout(C.byref(c), hi - lo, b.S, b.K, _p(b.ships[lo:hi]), _p(b.planets[lo:hi]),
                     _p(b.np_[lo:hi]), _p(b.bullets[lo:hi]), _p(b.nb[lo:hi]), _p(b.reload[lo:hi]),
                     _p(b.t[lo:hi]), _p(b.episode[lo:hi]), M, _p(ps), _p(pp), _p(pn), seed,
                     first_game + lo, step0, n_ticks, 1, _p(st))
        return st

    threads = max(1, min(threads, b.n))
    cuts = np.linspace(0, b.n, threads + 1).astype(np.int64)
    if threads == 1:
        return run(0, b.n)
    with ThreadPoolExecutor(threads) as ex:
        parts = list(ex.map(lambda k: run(int(cuts[k]), int(cuts[k + 1])), range(threads)))
    return np.sum(parts, axis=0)


def actions(seed, first_game, n, step, S):
    out = np.empty((n, S), dtype=np.int64)
    lib().ao_actions(seed, first_game, n, step, S, _p(out))
    return out


def pool_pick(seed, first_game, episode, M):
    episode = np.ascontiguousarray(episode, dtype=np.uint32)
    out = np.empty(episode.shape[0], dtype=np.int64)
    lib().ao_pool_pick(seed, first_game, episode.shape[0], _p(episode), M, _p(out))
    return out
