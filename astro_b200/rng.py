"""Counter-based streams shared by the device kernels and the host harness.

The synthetic workload needs uniform controls 0..5 per ship per
tick and a reset-pool pick per finished episode that can be regenerated on the
host, bit for bit, for the oracle comparison.  Both are pure functions of small
integer keys — no state, no ordering constraints between games.

The same arithmetic lives in `csrc/astro_device.cuh` (`mix32`, `action_for`,
`pool_pick`); this file is the host-side (numpy, vectorised) statement of it.
"""
import numpy as np

_U32 = np.uint32
_M1 = _U32(0x7FEB352D)
_M2 = _U32(0x846CA68B)
_GOLD = _U32(0x9E3779B1)
_POOL_SALT = _U32(0xA5A5A5A5)


def mix32(x):
    """32-bit avalanche mixer (xorshift-multiply, "lowbias32" constants)."""
    with np.errstate(over='ignore'):
        x = np.asarray(x, dtype=_U32).copy()
        x ^= x >> _U32(16)
        x *= _M1
        x ^= x >> _U32(15)
        x *= _M2
        x ^= x >> _U32(16)
    return x


def _mulhi(h, m):
    return ((h.astype(np.uint64) * np.uint64(m)) >> np.uint64(32)).astype(np.int64)


def actions(seed, game_ids, step, nships):
    """Controls for `game_ids` at rollout step `step` -> int64 [len(game_ids), nships].

    key = (seed, global game id, rollout step index, ship); value uniform in 0..5.
    """
    with np.errstate(over='ignore'):
        g = np.asarray(game_ids, dtype=_U32)
        h0 = mix32(_U32(seed) ^ (g * _GOLD))
        out = np.empty((g.shape[0], nships), dtype=np.int64)
        for s in range(nships):
            k = _U32((int(step) * 2 + s) & 0xFFFFFFFF)
            out[:, s] = _mulhi(mix32(h0 ^ k), 6)
    return out


def pool_pick(seed, game_ids, episode, pool_size):
    """Reset-pool entry for global game `g`; `episode` is the pick key: 0 for the initial fill,
    1 + the stream step of the tick that ended the previous game afterwards."""
    with np.errstate(over='ignore'):
        g = np.asarray(game_ids, dtype=_U32)
        e = np.asarray(episode, dtype=_U32)
        h0 = mix32(_U32(seed) ^ _POOL_SALT ^ (g * _GOLD))
        return _mulhi(mix32(h0 ^ e), pool_size)


def config_seeds(seed, count, skip=0):
    """Seeds of the first `count` configs of `core.generate_configs(config)` after skipping `skip`
    (core.py:77-83: RandomState(config.seed).randint(2**30) per config) -> uint32 [count].
    The bulk draw consumes the MT19937 stream exactly like `count` single draws."""
    stream = np.random.RandomState(int(seed))
    if skip:
        stream.randint(1 << 30, size=int(skip))
    return stream.randint(1 << 30, size=int(count)).astype(np.uint32)


_EXPLORE_SALT = _U32(0x3C6EF372)
_EXPLORE_PICK = _U32(0x85EBCA6B)


def explore_step(seed, game_ids, step, ticks, state, dt, t_in, t_out, live=None):
    """Host twin of csrc explore_kernel (rl.EpsilonGreedy.__call__, rl.py:10-30) for every ship.

    state  int32 [n, S]: (tick of the previous call) << 8 | (control + 1), 0 = idle; updated in place
    ticks  int   [n]: the games' tick counters; live bool [n] (finished games are skipped)
    returns int64 [n, S]: the random control where a ship's random policy is active, else -1."""
    g = np.asarray(game_ids, dtype=_U32)
    n, S = state.shape
    live = np.ones(n, dtype=bool) if live is None else np.asarray(live, dtype=bool)
    out = np.full((n, S), -1, dtype=np.int64)
    with np.errstate(over='ignore'):
        base = mix32(_U32(seed) ^ _EXPLORE_SALT ^ (g * _GOLD))
        for s in range(S):
            h = mix32(base ^ _U32((int(step) * 2 + s) & 0xFFFFFFFF))
            u = (h >> _U32(8)).astype(np.float64) * (1.0 / 16777216.0)
            policy = (state[:, s] & 0xff) - 1
            gap = dt * (np.asarray(ticks, dtype=np.int64) - (state[:, s] >> 8)).astype(np.float64)
            enter = (policy < 0) & (np.exp(-gap / t_in) < u)
            leave = (policy >= 0) & (np.exp(-gap / t_out) < u)
            pick = _mulhi(mix32(h ^ _EXPLORE_PICK), 5)
            policy = np.where(enter, pick, np.where(leave, -1, policy))
            new_state = ((np.asarray(ticks, dtype=np.int64) << 8) | (policy + 1)).astype(np.int32)
            state[:, s] = np.where(live, new_state, state[:, s])
            out[:, s] = np.where(live, policy, -1)
    return out
