"""CPU-side tests: host logic of the drop-in API against the reference's golden vectors, the
schedule, the C-ABI library's exports, and the multi-rank sharding logic over gloo.  No compute
call into the CUDA library happens here (there is no GPU in the build container)."""
import ctypes
import itertools as it
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from astro_b200 import core, rng
from astro_b200 import _native as nat
from astro_b200.schedule import Schedule
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_create_matches_reference_bits_and_dtypes():
    """core.create (core.py:86-135): 48 seeded configs, values bit-identical, dtypes identical."""
    z = np.load(os.path.join(H.G, 'create.npz'))
    meta = json.load(open(os.path.join(H.G, 'create.json')))
    assert len(meta) == 48
    for m in meta:
        s = core.create(H.config_from(m['config']))
        for name, arr in (('ships_x', s.ships.x), ('ships_dx', s.ships.dx), ('ships_b', s.ships.b),
                          ('planets_x', s.planets.x), ('planets_dx', s.planets.dx)):
            ref = z['%s_%s' % (m['key'], name)]
            assert str(arr.dtype) == m['dtypes'][name], (m['key'], name)
            assert arr.shape == ref.shape and arr.tobytes() == ref.tobytes(), (m['key'], name)
        assert s.bullets.x.shape == (0, 2) and s.reload == 0.0 and s.t == 0.0 and s.planets.b is None


def test_create_is_deterministic_and_default_known_answer():
    a, b = core.create(core.DEFAULT_CONFIG), core.create(core.DEFAULT_CONFIG)
    assert np.array_equal(a.ships.x, b.ships.x) and np.array_equal(a.planets.dx, b.planets.dx)
    np.testing.assert_allclose(a.ships.x, [[-0.19652985, 0.03709475], [0.9, -0.9]], atol=1e-7)
    assert a.planets.x.shape[0] == 3


def test_generate_configs_seeds():
    kat = json.load(open(os.path.join(H.G, 'kat.json')))
    got = [int(c.seed) for c in it.islice(core.generate_configs(core.DEFAULT_CONFIG), 8)]
    assert got == kat['generate_configs_seeds']
    assert got[:4] == [534895718, 199900595, 862061404, 787846414]


def test_direction_known_answers():
    kat = json.load(open(os.path.join(H.G, 'kat.json')))['direction']
    d = core.direction(np.array(kat['bearing']))
    assert d.dtype == np.float32
    np.testing.assert_allclose(d, kat['expected'], atol=kat['atol'])


def test_schedule_matches_reference_observation():
    """Spawn ticks / timeout tick / the reload and t sequences observed on the reference."""
    sched = json.load(open(os.path.join(H.G, 'schedule.json')))
    for name, s in sched.items():
        sc = Schedule(H.config_from(s['config']))
        assert sc.fire_ticks == s['spawn_ticks'], name
        assert sc.timeout_tick == s['timeout_tick'] and sc.n_ticks == s['timeout_tick'] + 1, name
        assert [float(v).hex() for v in sc.reload] == s['reload_hex'], name
        assert [float(v).hex() for v in sc.t] == s['t_hex'], name
        for k in s['spawn_ticks'][:5]:
            assert (sc.fire_bits[k >> 5] >> (k & 31)) & 1
        assert int(sum(bin(int(w)).count('1') for w in sc.fire_bits)) == len(s['spawn_ticks'])
    d = Schedule(core.DEFAULT_CONFIG)
    assert d.fire_ticks[:3] == [14, 29, 44] and d.timeout_tick == 2999
    assert d.tick_of(d.reload[77], d.t[77]) == 77 and d.tick_of(0.123, 0.5) is None


def test_schedule_from_arbitrary_origin():
    s = Schedule(core.DEFAULT_CONFIG, reload0=0.28, t0=59.95)
    assert s.fire_ticks == [0] and s.timeout_tick == 2
    with pytest.raises(ValueError):
        Schedule(core.DEFAULT_CONFIG._replace(max_time=1e9))


def test_roll_ships_and_bots():
    s = core.create(core.DEFAULT_CONFIG)
    r = core.roll_ships(s, 1)
    assert np.array_equal(r.ships.x[0], s.ships.x[1]) and np.array_equal(r.ships.b[1], s.ships.b[0])
    assert r.planets is s.planets and core.roll_ships(None, 1) is None

    class Echo(core.Bot):
        def __init__(self):
            self.seen = []

        def __call__(self, state):
            return 3 if state.ships.x[0, 0] == s.ships.x[0, 0] else 4

        def reward(self, state, reward):
            self.seen.append(reward)
    bots = [Echo(), Echo()]
    assert core.Bots.control(bots, s).tolist() == [3, 4]
    core.Bots.reward(bots, None, np.array([1, -1]))
    assert bots[0].seen == [1] and bots[1].seen == [-1]
    assert core.Bots.data(bots) == [None, None]


def test_log_roundtrip_and_reference_format(tmp_path):
    """save_log/load_log (core.py:413-443): JSONL, `_type` tags and {_values,_shape} arrays."""
    s = core.create(core.DEFAULT_CONFIG)
    game = core.Game(config=core.DEFAULT_CONFIG, winner=1, ticks=[
        core.Tick(state=s, control=np.array([2, 3]), reward=np.array([0.0, 0.0], dtype=np.float32), bot_data=[None, {'a': 1}])])
    path = str(tmp_path / 'sub' / 'g.jsonl')
    core.save_log(path, game)
    lines = open(path).read().strip().split('\n')
    head, tick = json.loads(lines[0]), json.loads(lines[1])
    assert head['config']['_type'] == 'astro.core:Config' and head['winner'] == 1
    assert tick['_type'] == 'astro.core:Tick' and tick['state']['ships']['x']['_shape'] == [2, 2]
    # arrays are NESTED lists, as the reference writes them (util.py:24-26): astro.js reads ships.x[i][0]
    assert tick['state']['ships']['x']['_values'] == s.ships.x.tolist() and len(tick['state']['ships']['x']['_values'][0]) == 2
    assert tick['control'] == {'_values': [2, 3], '_shape': [2]}
    back = core.load_log(path)
    assert back.config == game.config and back.winner == 1
    assert np.allclose(back.ticks[0].state.ships.x, s.ships.x) and back.ticks[0].state.planets.b is None


def test_rng_streams():
    a = rng.actions(0, np.arange(100000), 3, 2)
    assert a.min() == 0 and a.max() == 5
    assert abs(np.bincount(a.ravel(), minlength=6) / a.size - 1 / 6).max() < 0.01
    p = rng.pool_pick(0, np.arange(100000), np.zeros(100000, dtype=np.uint32), 4096)
    assert p.min() >= 0 and p.max() < 4096 and len(np.unique(p)) > 4000


def test_cabi_library_loads_and_exports_every_declared_symbol():
    """The shared library exists (built by __graft_entry__.build), loads without a GPU, and
    exports exactly the entry points include/astro_b200.h declares."""
    header = open(os.path.join(ROOT, 'include', 'astro_b200.h')).read()
    declared = set(re.findall(r'^(?:int|int64_t|const char\*)\s+(astro_\w+)\s*\(', header, flags=re.M))
    assert declared == set(nat.EXPORTS)
    from astro_b200 import build
    build.build_native()
    L = ctypes.CDLL(nat.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert nat.lib().astro_abi_version() == nat.ABI_VERSION
    out = subprocess.run(['nm', '-D', '--defined-only', nat.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if ' T ' in l and l.split()[-1].startswith('astro_')}
    assert exported == declared


def test_no_product_import_of_the_oracle():
    """The product package never references oracle/ (a CPU fallback would void parity)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'astro_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in text and 'from oracle' not in text and 'astro_oracle' not in text, f


def test_product_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from astro_b200.batched import BatchedGames
    with pytest.raises(RuntimeError, match='CUDA'):
        BatchedGames(core.DEFAULT_CONFIG, 32)
    with pytest.raises(RuntimeError, match='CUDA'):
        core.step(core.create(core.DEFAULT_CONFIG), np.array([2, 2]), core.DEFAULT_CONFIG)


def test_two_rank_gloo_sharding_and_stats_reduce():
    """world_size 2 over gloo: disjoint env slices, per-rank first_game offsets, the stats
    all-reduce and the max-over-ranks timing reduce that bench.py uses (bench.shard_plan /
    bench.reduce_stats), exercised with the oracle standing in for the device tick."""
    script = os.path.join(ROOT, 'tests', 'gloo_worker.py')
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT='29611', PYTHONPATH=ROOT)
    procs = [subprocess.Popen([sys.executable, script], env=dict(env, RANK=str(r), WORLD_SIZE='2', LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    res = [json.loads(o.strip().splitlines()[-1]) for o in outs]
    assert res[0]['total'] == res[1]['total'] and res[0]['total']['env_steps'] == 2 * 256 * 50
    assert res[0]['first_game'] == 0 and res[1]['first_game'] == 256
    assert res[0]['single_process_digest'] == res[0]['sharded_digest'] == res[1]['sharded_digest']
    assert res[0]['max_ms'] == res[1]['max_ms'] >= max(res[0]['my_ms'], res[1]['my_ms']) - 1e-9


def test_forward_both_equals_forward_on_rolled_features():
    """rl.ValueNetwork.forward_both: ship 1's view = ship 0's features with the ship column groups
    exchanged (core.roll_ships + rl.py:62-70), folded into the first layer's weights."""
    import torch
    from astro_b200 import rl
    torch.manual_seed(1)
    net = rl.ValueNetwork(solo=False, nout=6)
    x = torch.randn(7, 12, 15)
    x[:, :, 0] = (torch.rand(7, 12) < 0.5).float()
    x[2, 5:] = -1.0
    x[4, 1:] = -1.0
    rolled = x.clone()
    rolled[..., 1:6], rolled[..., 6:11] = x[..., 6:11], x[..., 1:6]
    rolled[x[..., 0] < 0] = -1.0
    with torch.no_grad():
        both = net.forward_both(x)
        assert both.shape == (7, 2, 6)
        assert torch.allclose(both[:, 0], net(x), atol=1e-6)
        assert torch.allclose(both[:, 1], net(rolled), atol=1e-6)


def test_tick_kernel_requests_its_rows_before_it_uses_meta():
    """The tick's first batch of loads (meta, ship rows, bearings, controls) is one round trip to HBM only
    if all of them are issued before the first meta-dependent load (planet rows are predicated on the
    planet count; the bullet list is requested with LDGSTS).  ptxas used to reorder this at random,
    costing 10 % of the tick — pinned with a warp barrier in load_tile_in (csrc/tick_f32.cuh); checked
    here on the SASS of the built library (no GPU needed)."""
    import shutil
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump):
        pytest.skip('cuobjdump not available')
    sass = subprocess.run([cuobjdump, '-sass', nat.LIB_PATH], capture_output=True, text=True).stdout
    for name in ('tick_f32_kernelILi2ELb1ELb0ELb0ELb0', 'tick_f32_kernelILi2ELb1ELb1ELb0ELb0', 'tick_f32_kernelILi2ELb1ELb1ELb0ELb1', 'tick_f32_kernelILi2ELb0ELb0ELb0ELb0', 'tick_f32_kernelILi1ELb1ELb0ELb0ELb0'):
        body = sass.split('Function : ')
        body = [b for b in body if name in b.split('\n', 1)[0]]
        assert len(body) == 1, name
        ops = [l for l in body[0].split('\n') if re.search(r'/\*[0-9a-f]{4}\*/', l) and re.search(r'\b(LDG|LDGSTS)\b', l.replace('.', ' '))]
        first_dependent = next(i for i, l in enumerate(ops) if 'LDGSTS' in l or re.search(r'@!?P\d+\s+LDG\.E\.128', l))
        head = ops[:first_dependent]
        n_ships = 2 if 'Li2E' in name else 1
        assert sum('LDG.E.128' in l for l in head) == n_ships, (name, head)            # ship rows
        assert sum(bool(re.search(r'LDG\.E\s', l)) for l in head) >= 1 + n_ships, (name, head)   # meta + bearings


def _golden_net(solo):
    import torch
    from astro_b200 import rl
    z = np.load(os.path.join(H.G, 'network.npz'))
    net = rl.ValueNetwork(solo=solo, nout=6)
    pre = 'solo_' if solo else 'duel_'
    sd = {k: torch.from_numpy(z[pre + k.replace('.', '_')]) for k in net.state_dict()}
    net.load_state_dict(sd)
    return net, z


def test_value_network_matches_reference_outputs():
    """astro_b200.rl.ValueNetwork.forward / forward_both with the reference's seeded weights (tests/golden/network.npz,
    produced by the unmodified rl.ValueNetwork.evaluate_batch / evaluate, rl.py:115-165) on the reference's own feature
    batches: outputs within 2e-6 (|q| <= 1), both perspectives, -1 padding neutral."""
    import torch
    f = np.load(os.path.join(H.G, 'features.npz'))
    fm = json.load(open(os.path.join(H.G, 'features.json')))
    n_duel = n_solo = 0
    for e in fm:
        g, solo = e['game'], e['nships'] == 1
        net, z = _golden_net(solo)
        batch = torch.from_numpy(f['g%d_batch' % g])
        with torch.no_grad():
            got = net(batch).numpy()
            assert np.abs(got - z['g%d_q0' % g]).max() <= 2e-6, g
            assert np.abs(got - z['g%d_q0_single' % g]).max() <= 2e-6, g
            if not solo:
                both = net.forward_both(batch).numpy()
                assert np.abs(both[:, 0] - z['g%d_q0' % g]).max() <= 2e-6 and np.abs(both[:, 1] - z['g%d_q1' % g]).max() <= 2e-6, g
                padded = torch.cat([batch, torch.full((batch.shape[0], 5, 15), -1.0)], dim=1)
                assert np.abs(net(padded).numpy() - z['g%d_q0' % g]).max() <= 2e-6
                n_duel += 1
            else:
                n_solo += 1
    assert n_duel >= 10 and n_solo >= 1
    assert np.abs(np.load(os.path.join(H.G, 'network.npz'))['g0_q0']).max() > 0.05      # (the outputs are not all ~0)


def test_library_seed_stream_equals_generate_configs():
    """astro_config_seeds (the host MT19937 that feeds fresh-game mode) against numpy's RandomState stream that
    core.generate_configs draws from (core.py:77-83), and against the reference's first seeds."""
    L = nat.lib()
    for seed, skip, n in ((42, 0, 3000), (7, 12345, 5000), (2 ** 31 + 5, 3, 10), (0, 623, 1300)):
        out = np.zeros(n, dtype=np.uint32)
        assert L.astro_config_seeds(seed & 0xFFFFFFFF, skip, n, out.ctypes.data_as(ctypes.c_void_p)) == 0
        assert (out == rng.config_seeds(seed, n, skip)).all(), (seed, skip)
    out = np.zeros(4, dtype=np.uint32)
    L.astro_config_seeds(42, 0, 4, out.ctypes.data_as(ctypes.c_void_p))
    assert out.tolist() == [534895718, 199900595, 862061404, 787846414]
    assert [c.seed for c in it.islice(core.generate_configs(core.DEFAULT_CONFIG), 4)] == out.tolist()
