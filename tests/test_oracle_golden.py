"""CPU tests: the oracle (oracle/atc_oracle.c) against golden vectors recorded from the live reference."""
import os

import numpy as np
import pytest

from oracle.oracle import Oracle
from tests import golden_util as G


class OracleImpl(object):
    def __init__(self, tr):
        m = tr['meta']
        E = tr['action'].shape[1]
        self.o = Oracle(sector=m['scenario'], random_entrypoints=m['random_entrypoints'], n_env=E, n_ac=1,
                        dt=m['dt'], reward_shaping=m['reward_shaping'], normalize_state=m['normalize_state'],
                        discrete=m['discrete'])

    def reset(self, mask, spawn):
        return self.o.reset(mask, spawn)

    def set_state(self, st, ts):
        self.o.set_state(st, ts)

    def step(self, a):
        return self.o.step(a, autoreset=False)

    def get_state(self):
        return self.o.get_state()

    def metrics(self):
        return self.o.metrics()


@pytest.mark.parametrize('name', G.trace_names())
def test_oracle_replays_reference_trace(name):
    tr = G.load_trace(name)
    worst = G.replay(tr, OracleImpl(tr))
    print(name, worst)


@pytest.mark.parametrize('scn', ['LOWW', 'SimpleScenario'])
def test_oracle_constants_match_reference(scn):
    k = G.kat()[scn]
    c = Oracle(sector=scn).constants()
    for key in ('faf', 'iaf', 'corner1', 'corner2'):
        np.testing.assert_allclose(c[key], k[key], rtol=0, atol=1e-12)
    np.testing.assert_allclose(c['normal'], k['faf_iaf_normal'], atol=1e-15)
    np.testing.assert_array_equal(c['bbox'], k['bbox'])
    np.testing.assert_allclose(c['dmax'], k['world_max_distance'], rtol=1e-15)
    assert c['faf_mva'] == k['faf_mva'] and c['phi_to'] == k['phi_to_runway']
    np.testing.assert_array_equal(c['nmin'].astype(np.float32), np.asarray(k['norm_min'], np.float32))
    np.testing.assert_array_equal(c['nmax'].astype(np.float32), np.asarray(k['norm_max'], np.float32))
    np.testing.assert_allclose(c['tri_h'], k['corridor_horizontal'], atol=1e-12)
    np.testing.assert_allclose(c['tri_1'], k['corridor1'], atol=1e-12)
    np.testing.assert_allclose(c['tri_2'], k['corridor2'], atol=1e-12)


@pytest.mark.parametrize('scn', ['LOWW', 'SimpleScenario'])
def test_oracle_geometry_matches_reference(scn):
    z = np.load(os.path.join(G.GOLDEN, 'geometry_%s.npz' % scn))
    o = Oracle(sector=scn)
    np.testing.assert_array_equal(o.mva(z['pts']), z['mva'])
    np.testing.assert_array_equal(o.inside_corridor(z['corr']), z['inside'])


def test_reference_unit_tests_model_test_py():
    """The reference's own 8 cases (envs/atc/model_test.py:10-92): SimpleScenario MVAs, runway (20, 20, 0, 180)."""
    k = G.kat()['model_test']
    import json, tempfile
    with open(os.path.join(os.path.dirname(G.GOLDEN), '..', 'atc_reinforcement_learning_b200', 'sectors',
                           'SimpleScenario.json')) as f:
        doc = json.load(f)
    doc['runway'] = {'x': 20, 'y': 20, 'h': 0, 'phi_from_runway': 180}
    with tempfile.NamedTemporaryFile('w', suffix='.json', delete=False) as f:
        json.dump(doc, f)
    o = Oracle(sector=f.name)
    os.unlink(f.name)
    assert k['get_mva_height_34_1'] == 3500 and o.mva([[34, 1]])[0] == 3500
    faf = o.constants()['faf']
    assert o.mva([faf])[0] == k['faf_mva']
    for x, y, h, phi, exp in k['inside_corridor']:
        assert bool(o.inside_corridor([[x, y, h, phi]])[0]) == exp
    assert [e for *_, e in k['inside_corridor']] == [True, False]
    for x, y, phi, exp in k['inside_corridor_angle']:
        assert bool(o.inside_corridor_angle([[x, y, phi]])[0]) == exp
    assert [e for *_, e in k['inside_corridor_angle']] == [False, False, False, True]
    np.testing.assert_array_equal(o.constants()['bbox'], [0.0, 0.0, 35.0, 40.0])
    assert k['bbox'] == [0.0, 0.0, 35.0, 40.0]


def test_kat_corridor_gates_and_k0():
    k = G.kat()
    o = Oracle(sector='LOWW')
    rows = np.asarray([r[:4] for r in k['K7_corridor']], np.float64)
    exp = np.asarray([r[4] for r in k['K7_corridor']], np.uint8)
    np.testing.assert_array_equal(o.inside_corridor(rows), exp)
    obs = Oracle(sector='LOWW').reset()
    np.testing.assert_allclose(obs.reshape(10), k['K0_reset_obs'], rtol=1e-7)
