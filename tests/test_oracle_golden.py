"""Pins the CPU oracle (oracle/astro_oracle.c) bit for bit against vectors produced by the
UNMODIFIED reference (tests/golden/make_golden.py ran /root/reference/astro/core.py,
util.py, rl.py in the build container).  CPU-only; nothing here touches the product path.
"""
import json
import os

import numpy as np
import pytest

from oracle import astro_oracle as ao

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def _same(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and bool((_bits(a) == _bits(b)).all())


# ---------------------------------------------------------------- reference's own KATs

def test_kat_collisions():
    """astro/test/test_core.py:6-17."""
    k = json.load(open(os.path.join(G, 'kat.json')))['collisions']
    assert k['hit'] == [False, True, True, True]
    assert ao.collisions(k['x'], k['r']).tolist() == k['hit']


def test_kat_geometry():
    """astro/test/test_util.py:59-93: direction, wrap_unit_square, norm_angle."""
    k = json.load(open(os.path.join(G, 'kat.json')))
    d = k['direction']
    s, c = ao.sincos_f32(np.array(d['bearing'], dtype=np.float32))
    got = np.stack([s, c], axis=1).astype(np.float64)
    np.testing.assert_allclose(got, d['expected'], atol=d['atol'])
    assert _same(got, d['value'])
    w = k['wrap_unit_square']
    got = ao.wrap_unit_square(np.array(w['x']))
    assert _same(got, w['value'])
    np.testing.assert_allclose(got[:2], w['expected_first2'], atol=w['atol'])
    n = k['norm_angle']
    got = ao.norm_angle(np.array(n['b']))
    assert _same(got, n['value'])
    np.testing.assert_allclose(got[:2], n['expected_first2'], atol=n['atol'])


def test_sincos_bits():
    """numpy's float32 sin/cos (util.direction, util.py:87-92): 42k inputs, bit-exact."""
    z = np.load(os.path.join(G, 'sincos.npz'))
    s, c = ao.sincos_f32(z['x'])
    assert (s.view(np.uint32) == z['sin_bits']).all()
    assert (c.view(np.uint32) == z['cos_bits']).all()


# ---------------------------------------------------------------- trajectories

def _traj():
    z = np.load(os.path.join(G, 'traj.npz'))
    meta = json.load(open(os.path.join(G, 'traj.json')))
    return z, meta


def test_trajectories_teacher_forced_and_free_running():
    """Every tick of 39 reference games (core.step, core.py:215-303): the oracle, fed the
    reference's pre-step state, reproduces the reference's post-step state, bullet order and
    count, reload, t, reward and done — and does so again when run free from tick 0."""
    z, meta = _traj()
    total = 0
    for m in meta:
        g = m['game']
        ships, planets = z['g%d_ships' % g], z['g%d_planets' % g]
        nb, bullets = z['g%d_nb' % g], z['g%d_bullets' % g]
        reload_, t_, ctrl, rew = z['g%d_reload' % g], z['g%d_t' % g], z['g%d_control' % g], z['g%d_reward' % g]
        off = np.concatenate([[0], np.cumsum(nb)])
        cfg = ao.Config.from_any(m['config'])
        n = m['nticks']
        free = None
        for k in range(n):
            cur = (ships[k], planets[k], bullets[off[k]:off[k + 1]], reload_[k], t_[k])
            if free is not None:
                assert all(_same(a, b) for a, b in zip(free, cur)), (g, k, 'free-running drifted')
            out = ao.step_one(cfg, *cur, ctrl[k])
            assert _same(out['reward'], rew[k]), (g, k)
            last = k == n - 1
            if last and not m['truncated']:
                assert out['done'], (g, k)
                free = None
                continue
            assert not out['done'], (g, k)
            if last:
                nxt = (z['g%d_final_ships' % g], z['g%d_final_planets' % g], z['g%d_final_bullets' % g],
                       z['g%d_final_reload_t' % g][0], z['g%d_final_reload_t' % g][1])
            else:
                nxt = (ships[k + 1], planets[k + 1], bullets[off[k + 1]:off[k + 2]], reload_[k + 1], t_[k + 1])
            got = (out['ships'], out['planets'], out['bullets'], out['reload'], out['t'])
            for name, a, b in zip(('ships', 'planets', 'bullets', 'reload', 't'), got, nxt):
                assert _same(a, b), (g, k, name)
            free = got
            total += 1
    assert total > 6000


def test_raw_create_trajectories_first_tick_float32():
    """Games the reference played from core.create's RAW arrays (float32 ships and planet positions, not
    canonicalised): on the first tick numpy runs _gravity, the squared distances of _collisions and the planets'
    a * dt in float32 (core.py:138-153, :200-212, :189), and bullets born on that tick are float32 throughout.  The
    oracle's raw mode reproduces the first tick bit for bit, the plain mode every later tick — teacher-forced and
    free-running; and the plain mode must NOT match the raw first tick of a multi-planet game (the pin is real)."""
    z = np.load(os.path.join(G, 'traj_raw.npz'))
    meta = json.load(open(os.path.join(G, 'traj_raw.json')))
    total, plain_differs = 0, 0
    for m in meta:
        g = m['game']
        ships, planets = z['g%d_ships' % g], z['g%d_planets' % g]
        nb, bullets = z['g%d_nb' % g], z['g%d_bullets' % g]
        reload_, t_, ctrl, rew = z['g%d_reload' % g], z['g%d_t' % g], z['g%d_control' % g], z['g%d_reward' % g]
        off = np.concatenate([[0], np.cumsum(nb)])
        cfg = ao.Config.from_any(m['config'])
        n = m['nticks']
        assert nb[0] == 0 and (ships[0].astype(np.float32) == ships[0]).all() and (planets[0][:, :2].astype(np.float32) == planets[0][:, :2]).all()
        free = None
        for k in range(n):
            cur = (ships[k], planets[k], bullets[off[k]:off[k + 1]], reload_[k], t_[k])
            if free is not None:
                assert all(_same(a, b) for a, b in zip(free, cur)), (g, k, 'free-running drifted')
            out = ao.step_one(cfg, *cur, ctrl[k], raw=(k == 0))
            assert _same(out['reward'], rew[k]), (g, k)
            last = k == n - 1
            if last and not m['truncated']:
                assert out['done'], (g, k)
                continue
            assert not out['done'], (g, k)
            if last:
                nxt = (z['g%d_final_ships' % g], z['g%d_final_planets' % g], z['g%d_final_bullets' % g],
                       z['g%d_final_reload_t' % g][0], z['g%d_final_reload_t' % g][1])
            else:
                nxt = (ships[k + 1], planets[k + 1], bullets[off[k + 1]:off[k + 2]], reload_[k + 1], t_[k + 1])
            got = (out['ships'], out['planets'], out['bullets'], out['reload'], out['t'])
            for name, a, b in zip(('ships', 'planets', 'bullets', 'reload', 't'), got, nxt):
                assert _same(a, b), (g, k, name)
            if k == 0:
                plain = ao.step_one(cfg, *cur, ctrl[k])
                plain_differs += not (_same(plain['ships'], nxt[0]) and _same(plain['planets'], nxt[1]) and _same(plain['bullets'], nxt[2]))
            free = got
            total += 1
    assert total > 1000 and plain_differs >= 8


def test_edge_cases():
    """70 hand-built single-step cases through the reference: predicate knife edges, the
    bullet-cull .any quirk, terminal precedence, firing order, wrap, all 36 control pairs."""
    z = np.load(os.path.join(G, 'edges.npz'))
    meta = json.load(open(os.path.join(G, 'edges.json')))
    assert len(meta) >= 70
    for c in meta:
        i = c['case']
        out = ao.step_one(c['config'], z['c%d_ships' % i], z['c%d_planets' % i], z['c%d_bullets' % i],
                          c['reload'], c['t'], c['control'])
        assert out['done'] == c['done'], c['name']
        assert _same(out['reward'], c['reward']), c['name']
        if not c['done']:
            assert _same(out['ships'], z['c%d_o_ships' % i]), c['name']
            assert _same(out['planets'], z['c%d_o_planets' % i]), c['name']
            assert _same(out['bullets'], z['c%d_o_bullets' % i].reshape(-1, 4)), c['name']
            assert _same(out['reload'], c['o_reload']) and _same(out['t'], c['o_t']), c['name']


def test_schedule():
    """reload / t are Python-float accumulators (core.py:257-280,302): spawn ticks and the
    timeout tick observed on the reference equal the oracle's."""
    sched = json.load(open(os.path.join(G, 'schedule.json')))
    assert sched['default']['spawn_ticks'][:3] == [14, 29, 44] and sched['default']['timeout_tick'] == 2999
    for name, s in sched.items():
        cfg = dict(s['config'], gravity=0.0)
        S = 1 if cfg['solo'] else 2
        ships = np.array([[-0.9, -0.9, 0, 0, -3 * np.pi / 4], [0.9, 0.9, 0, 0, np.pi / 4]])[:S]
        planets, bullets = np.zeros((1, 4)), np.zeros((0, 4))
        reload_, t, k, spawn = 0.0, 0.0, 0, []
        ctl = np.full(S, 2)
        while True:
            assert float(reload_).hex() == s['reload_hex'][k] and float(t).hex() == s['t_hex'][k]
            out = ao.step_one(cfg, ships, planets, bullets, reload_, t, ctl)
            if out['done']:
                break
            if out['events'] & ao.EV_FIRED:
                spawn.append(k)
            ships, planets, bullets, reload_, t = (out[x] for x in ('ships', 'planets', 'bullets', 'reload', 't'))
            k += 1
        assert spawn == s['spawn_ticks'] and k == s['timeout_tick'], name


def test_script_bot_controls():
    """script.ScriptBot (script.py:13-91) on every tick of the reference's scripted games: the
    oracle picks the control the reference's bot picked (both perspectives, solo too)."""
    z, meta = _traj()
    checked = 0
    for m in meta:
        if 'script' not in m['kind']:
            continue
        g, S = m['game'], m['nships']
        ships, planets, ctrl = z['g%d_ships' % g], z['g%d_planets' % g], z['g%d_control' % g]
        scripted = range(S) if m['kind'] != 'duel_nothing_vs_script' else [1]
        for k in range(m['nticks']):
            for me in scripted:
                got = ao.script_control(m['config'], ships[k], planets[k], me)
                assert got == ctrl[k][me], (g, k, me)
                checked += 1
    assert checked > 2000


def test_features():
    """rl.ValueNetwork.get_features / roll_ships / get_features_batch (rl.py:43-112,
    core.py:306-327) on states sampled from the reference trajectories."""
    z, meta = _traj()
    f = np.load(os.path.join(G, 'features.npz'))
    fm = json.load(open(os.path.join(G, 'features.json')))
    n = 0
    for e in fm:
        g, S = e['game'], e['nships']
        nb = z['g%d_nb' % g]
        off = np.concatenate([[0], np.cumsum(nb)])
        rows = []
        for k in e['ticks']:
            sh, pl, bl = z['g%d_ships' % g][k], z['g%d_planets' % g][k], z['g%d_bullets' % g][off[k]:off[k + 1]]
            nrow = pl.shape[0] + bl.shape[0]
            for me in range(S):
                ref = f['g%d_t%d_f%d' % (g, k, me)]
                got = ao.features(S, sh, pl, bl, me, nrow)
                assert ref.dtype == np.float32 and got.shape == ref.shape
                assert (got.view(np.uint32) == ref.view(np.uint32)).all(), (g, k, me)
                n += 1
            rows.append((sh, pl, bl))
        ref = f['g%d_batch' % g]
        nmax = ref.shape[1]
        got = np.stack([ao.features(S, sh, pl, bl, 0, nmax) for sh, pl, bl in rows])
        assert (got.view(np.uint32) == ref.view(np.uint32)).all(), g
    assert n > 100


def test_batch_matches_single():
    """ao_step_batch (fixed-stride host image) == ao_step_one game by game, incl. alive mask."""
    z, meta = _traj()
    duel = [m for m in meta if m['nships'] == 2 and m['kind'] == 'duel_random']
    K = 32
    b = ao.Batch(len(duel), 2, K)
    ctl = np.zeros((b.n, 2), dtype=np.int64)
    for i, m in enumerate(duel):
        g = m['game']
        k = m['nticks'] - 1 if i % 3 == 0 else m['nticks'] // 2
        nb = z['g%d_nb' % g]
        off = np.concatenate([[0], np.cumsum(nb)])
        P = m['nplanets']
        b.ships[i] = z['g%d_ships' % g][k]
        b.planets[i, :P] = z['g%d_planets' % g][k]
        b.np_[i] = P
        b.nb[i] = nb[k]
        b.bullets[i, :nb[k]] = z['g%d_bullets' % g][off[k]:off[k + 1]]
        b.reload[i], b.t[i] = z['g%d_reload' % g][k], z['g%d_t' % g][k]
        ctl[i] = z['g%d_control' % g][k]
    alive = np.ones(b.n, dtype=np.uint8)
    alive[1] = 0
    o, reward, done, events = ao.step_batch(duel[0]['config'], b, ctl, alive)
    assert done[0] == 1 and done[1] == 1 and reward[1].tolist() == [0, 0]
    for i in range(b.n):
        if not alive[i]:
            continue
        P, B = b.np_[i], b.nb[i]
        one = ao.step_one(duel[0]['config'], b.ships[i], b.planets[i, :P], b.bullets[i, :B], b.reload[i], b.t[i],
                          ctl[i], bullet_cap=K)
        assert one['done'] == bool(done[i]) and _same(one['reward'], reward[i]) and one['events'] == events[i]
        if not one['done']:
            assert _same(one['ships'], o.ships[i]) and _same(one['planets'], o.planets[i, :P])
            assert o.nb[i] == one['bullets'].shape[0] and _same(one['bullets'], o.bullets[i, :o.nb[i]])
            assert _same(one['reload'], o.reload[i]) and _same(one['t'], o.t[i])


def test_counter_streams_match_host_rng():
    from astro_b200 import rng
    for seed in (0, 7, 123456789):
        for step in (0, 1, 999, 70000):
            a = ao.actions(seed, 5, 1000, step, 2)
            assert (a == rng.actions(seed, np.arange(5, 1005), step, 2)).all()
            assert a.min() >= 0 and a.max() <= 5
        ep = np.arange(1000, dtype=np.uint32) % 17
        assert (ao.pool_pick(seed, 5, ep, 4096) == rng.pool_pick(seed, np.arange(5, 1005), ep, 4096)).all()
