"""Shared helpers of the parity tests: golden loaders, pools, oracle <-> BatchedGames glue."""
import itertools as it
import json
import os

import numpy as np

from astro_b200 import core, rng
from oracle import astro_oracle as ao

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def same_bits(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and bool((bits(a) == bits(b)).all())


def config_from(d):
    return core.Config(**{k: d[k] for k in core.Config._fields})


def load_traj():
    z = np.load(os.path.join(G, 'traj.npz'))
    meta = json.load(open(os.path.join(G, 'traj.json')))
    return z, meta


def state_from_arrays(ships, planets, bullets, reload_, t):
    ships = np.asarray(ships, dtype=np.float64).reshape(-1, 5)
    planets = np.asarray(planets, dtype=np.float64).reshape(-1, 4)
    bullets = np.asarray(bullets, dtype=np.float64).reshape(-1, 4)
    return core.State(
        ships=core.Bodies(ships[:, 0:2].copy(), ships[:, 2:4].copy(), ships[:, 4].copy()),
        planets=core.Bodies(planets[:, 0:2].copy(), planets[:, 2:4].copy(), None),
        bullets=core.Bodies(bullets[:, 0:2].copy(), bullets[:, 2:4].copy(), None),
        reload=float(reload_), t=float(t))


from astro_b200.pool import make_pool  # noqa: E402,F401


def oracle_batch_from(arr, S, K):
    """get_arrays() dict -> oracle Batch (float64 image) + alive mask."""
    n = arr['ships'].shape[0]
    b = ao.Batch(n, S, K)
    b.ships[:] = arr['ships']
    b.planets[:] = arr['planets']
    b.np_[:] = arr['n_planets']
    b.bullets[:] = arr['bullets']
    b.nb[:] = arr['n_bullets']
    b.episode[:] = arr['episode']
    return b, (~arr['finished']).astype(np.uint8)


def start_from_pool(games, pool, seed, first_game):
    """Host twin of BatchedGames.reset_all(): game g <- pool[pick(seed, first_game+g, episode 0)]."""
    n = games.n
    pick = rng.pool_pick(seed, first_game + np.arange(n), np.zeros(n, dtype=np.uint32), pool['ships'].shape[0])
    return pick
