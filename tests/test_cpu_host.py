"""CPU tests (no GPU): the host side of the product — C-ABI library exports and struct layout, the sector compiler and its
exact MVA grid (checked against the oracle's brute-force scan), scenario files, argument validation — plus the oracle's
own extension semantics (several aircraft, separation, wind, spawn RNG) and their degeneration to the pinned path."""
import ctypes as C
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle.oracle import Oracle
from tests import golden_util as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------------------------------------ C ABI
def test_library_loads_and_exports_every_declared_symbol():
    from atc_reinforcement_learning_b200 import _native as nat
    nat.build_library()
    with open(os.path.join(ROOT, 'include', 'atc_b200.h')) as f:
        hdr = f.read()
    declared = sorted(set(re.findall(r'\b(atc_[a-z_0-9]+)\s*\(', hdr)))
    assert len(declared) >= 12
    lib = C.CDLL(nat.LIB_PATH)           # loads without a GPU / driver (static cudart)
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(nat.EXPORTS) == declared
    assert nat.lib().atc_abi_version() == nat.ABI_VERSION == int(re.search(r'#define ATC_ABI_VERSION (\d+)', hdr).group(1))


def test_ctypes_structs_match_the_c_header(tmp_path):
    from atc_reinforcement_learning_b200 import _native as nat
    src = tmp_path / 'sz.c'
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "atc_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(AtcSectorDesc),sizeof(AtcSimParams),sizeof(AtcBuffers),sizeof(AtcStepIO),'
                   'offsetof(AtcSectorDesc,grid_cell),offsetof(AtcSectorDesc,wind),offsetof(AtcSimParams,seed));return 0;}\n')
    exe = tmp_path / 'sz'
    subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), '-o', str(exe), str(src)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    exp = [C.sizeof(nat.AtcSectorDesc), C.sizeof(nat.AtcSimParams), C.sizeof(nat.AtcBuffers), C.sizeof(nat.AtcStepIO),
           nat.AtcSectorDesc.grid_cell.offset, nat.AtcSectorDesc.wind.offset, nat.AtcSimParams.seed.offset]
    assert got == exp


def test_create_rejects_bad_arguments_without_a_gpu():
    """Argument validation happens before any CUDA call, so it can be exercised here; errors are status codes +
    atc_last_error, never exceptions across the ABI."""
    from atc_reinforcement_learning_b200 import _native as nat, LOWW, CompiledSector
    lib = nat.lib()
    cs = CompiledSector(LOWW())
    desc = nat.sector_desc(cs)
    p = nat.AtcSimParams(timestep=1.0, reward_shaping=1, normalize_state=1, n_env=4, n_aircraft=9)
    h = C.c_void_p()
    assert lib.atc_create(C.byref(desc), C.byref(p), 0, C.byref(h)) == -1 and not h.value
    assert b'n_aircraft' in lib.atc_last_error(None)
    p.n_aircraft = 1
    p.timestep = 0.0
    assert lib.atc_create(C.byref(desc), C.byref(p), 0, C.byref(h)) == -1
    assert b'timestep' in lib.atc_last_error(None)
    assert lib.atc_create(None, C.byref(p), 0, C.byref(h)) == -1
    assert lib.atc_step(None, None, None, 0, None) == -1


def test_env_refuses_to_run_without_cuda():
    import torch
    from atc_reinforcement_learning_b200 import BatchedAtcEnv
    with pytest.raises(ValueError):
        BatchedAtcEnv(4, 1, device='cpu')
    with pytest.raises(ValueError):
        BatchedAtcEnv(4, 9)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match='no CPU path'):
            BatchedAtcEnv(4, 1)


# ------------------------------------------------------------------------------------------------ sector compiler
@pytest.mark.parametrize('scn', ['LOWW', 'SimpleScenario'])
def test_sector_compiler_constants_match_reference(scn):
    import atc_reinforcement_learning_b200 as P
    k = G.kat()[scn]
    cs = P.CompiledSector(getattr(P, scn)())
    for key, val in (('faf', cs.faf), ('iaf', cs.iaf), ('corner1', cs.corner1), ('corner2', cs.corner2)):
        np.testing.assert_allclose(val, k[key], rtol=0, atol=1e-12)
    np.testing.assert_allclose(cs.normal, k['faf_iaf_normal'], atol=1e-15)
    assert cs.bbox.tolist() == k['bbox'] and cs.faf_mva == k['faf_mva'] and cs.phi_to_runway == k['phi_to_runway']
    np.testing.assert_allclose(cs.world_max_distance, k['world_max_distance'], rtol=1e-15)
    np.testing.assert_array_equal(cs.norm_min, np.asarray(k['norm_min'], np.float32))
    np.testing.assert_array_equal(cs.norm_max, np.asarray(k['norm_max'], np.float32))
    np.testing.assert_allclose(cs.tri_h, k['corridor_horizontal'], atol=1e-12)
    assert [int(h) for h in cs.mva_height] == k['mva_heights']
    assert np.diff(cs.ring_off).tolist() == k['mva_ring_sizes']
    np.testing.assert_array_equal(cs.mva_bounds, np.asarray(k['mva_bounds']))
    # the oracle derives the same constants independently (C, libm)
    oc = Oracle(scn).constants()
    np.testing.assert_allclose(cs.faf, oc['faf'], atol=1e-13)
    np.testing.assert_allclose(cs.tri_1, oc['tri_1'], atol=1e-13)


@pytest.mark.parametrize('scn,cell', [('LOWW', 0.25), ('LOWW', 0.5), ('LOWW', 0.1), ('SimpleScenario', 0.25)])
def test_mva_grid_is_exact(scn, cell):
    """The grid + per-cell programs reproduce the reference's first-match scan on the golden points, on random
    points, and on points placed on / next to every vertex, edge, cell border and cell corner."""
    import atc_reinforcement_learning_b200 as P
    cs = P.CompiledSector(getattr(P, scn)(), cell=cell)
    ora = Oracle(scn)
    z = np.load(os.path.join(G.GOLDEN, 'geometry_%s.npz' % scn))
    idx = cs.lookup_np(z['pts'][:, 0], z['pts'][:, 1])
    h = np.where(idx < 0, -1, cs.mva_height[np.maximum(idx, 0)]).astype(np.int32)
    np.testing.assert_array_equal(h, z['mva'])
    rng = np.random.RandomState(5)
    n = 300000
    pts = [np.stack([rng.uniform(cs.bbox[0] - 1, cs.bbox[2] + 1, n), rng.uniform(cs.bbox[1] - 1, cs.bbox[3] + 1, n)], 1)]
    # cell borders / corners, +- a few ulps and +- 1e-9
    gx = cs.grid_x0 + np.arange(cs.grid_nx + 1) * cell
    gy = cs.grid_y0 + np.arange(cs.grid_ny + 1) * cell
    bx = rng.choice(gx, 40000)
    by = rng.choice(gy, 40000)
    for d in (0.0, 1e-9, -1e-9):
        pts.append(np.stack([bx + d, rng.uniform(cs.bbox[1], cs.bbox[3], 40000)], 1))
        pts.append(np.stack([rng.uniform(cs.bbox[0], cs.bbox[2], 40000), by + d], 1))
        pts.append(np.stack([bx + d, by - d], 1))
    pts.append(np.stack([np.nextafter(bx, np.inf), np.nextafter(by, -np.inf)], 1))
    # along every edge, and just off it
    for ring in cs.rings:
        for i in range(1, len(ring)):
            t = rng.uniform(0, 1, 400)[:, None]
            on = ring[i - 1] * (1 - t) + ring[i] * t
            nrm = np.array([ring[i][1] - ring[i - 1][1], ring[i - 1][0] - ring[i][0]])
            nrm = nrm / (np.linalg.norm(nrm) + 1e-300)
            for d in (0.0, 1e-12, -1e-12, 1e-7, -1e-7, 1e-3, -1e-3):
                pts.append(on + d * nrm)
    pts = np.concatenate(pts, 0)
    want = ora.mva_index(pts)
    np.testing.assert_array_equal(cs.lookup_np(pts[:, 0], pts[:, 1]), want)
    # The kernel picks the cell from float32 coordinates, so near a border it may take the neighbour: every cell whose
    # rectangle comes within `margin` of the point must give the same exact answer, and the float32 index must be one
    # of those cells.
    mg = cs.margin
    ix32, iy32 = cs.cell_index_np(pts[:, 0], pts[:, 1])
    inside = ((pts[:, 0] > cs.grid_x0 + cell) & (pts[:, 0] < cs.grid_x0 + (cs.grid_nx - 1) * cell) &
              (pts[:, 1] > cs.grid_y0 + cell) & (pts[:, 1] < cs.grid_y0 + (cs.grid_ny - 1) * cell))
    lo_x = np.floor((pts[:, 0] - mg - cs.grid_x0) / cell).astype(np.int64)
    hi_x = np.floor((pts[:, 0] + mg - cs.grid_x0) / cell).astype(np.int64)
    lo_y = np.floor((pts[:, 1] - mg - cs.grid_y0) / cell).astype(np.int64)
    hi_y = np.floor((pts[:, 1] + mg - cs.grid_y0) / cell).astype(np.int64)
    assert ((ix32 >= lo_x) & (ix32 <= hi_x) & (iy32 >= lo_y) & (iy32 <= hi_y))[inside].all()
    amb = np.nonzero(inside & ((lo_x != hi_x) | (lo_y != hi_y)))[0]
    assert len(amb) > 1000
    for cx, cy in ((lo_x, lo_y), (hi_x, lo_y), (lo_x, hi_y), (hi_x, hi_y)):
        got = cs.lookup_np(pts[amb, 0], pts[amb, 1], cells=(cx[amb], cy[amb]))
        np.testing.assert_array_equal(got, want[amb])
    # NaN / inf are "outside"
    bad = np.array([[np.nan, 10.0], [10.0, np.nan], [np.inf, 10.0], [-np.inf, -np.inf]])
    assert (cs.lookup_np(bad[:, 0], bad[:, 1]) == -1).all() and (ora.mva_index(bad) == -1).all()


def test_scenario_files_and_validation(tmp_path):
    import atc_reinforcement_learning_b200 as P
    from atc_reinforcement_learning_b200.scenarios import SECTOR_DIR
    loww = P.LOWW(random_entrypoints=True)
    assert len(loww.mvas) == 12 and len(loww.entrypoints) == 9 and len(P.LOWW().entrypoints) == 1
    assert loww.runway.phi_to_runway == 340 and P.SimParameters(1).timestep == 1
    assert G.kat()['LOWW_random_entrypoints'] == [[e.x, e.y, e.phi, e.levels] for e in loww.entrypoints]
    with open(os.path.join(SECTOR_DIR, 'LOWW.json')) as f:
        doc = json.load(f)
    bad = dict(doc, format='something-else')
    p = tmp_path / 'bad.json'
    p.write_text(json.dumps(bad))
    with pytest.raises(ValueError):
        P.load_scenario(str(p))
    doc['runway'] = {'x': 500.0, 'y': 500.0, 'h': 0, 'phi_from_runway': 90}     # FAF outside the airspace
    p.write_text(json.dumps(doc))
    with pytest.raises(ValueError):
        P.CompiledSector(P.load_scenario(str(p)))
    with pytest.raises(ValueError):
        P.SimParameters(0)
    with pytest.raises(ValueError):
        P.CompiledSector(P.LOWW(), wind=np.zeros((1, 4, 2), np.float32))


# ------------------------------------------------------------------------------------------------ oracle extensions
def _acts(rng, T, N, A, repeat=20):
    return np.repeat(rng.uniform(-1, 1, ((T + repeat - 1) // repeat, N, A, 3)).astype(np.float32), repeat, 0)[:T]


def test_oracle_multi_aircraft_degenerates_to_single_aircraft():
    """A = 4, aircraft 2000 ft apart (never in conflict): every aircraft's observation is bit-identical to an
    independent single-aircraft (reference-pinned) env, reward is the tree sum, done is the OR."""
    N, A, T = 64, 4, 150
    rng = np.random.RandomState(3)
    spawn = np.zeros((N, A, 5))
    spawn[..., 0] = rng.uniform(25, 45, (N, A)); spawn[..., 1] = rng.uniform(35, 55, (N, A))
    spawn[..., 2] = 8000 + 2000 * np.arange(A)[None, :]
    spawn[..., 3] = rng.uniform(0, 360, (N, A)); spawn[..., 4] = 250
    acts = rng.uniform(-1, 1, (T, N, A, 3)).astype(np.float32)
    acts[..., 1] = (spawn[..., 2] / 19000 - 1)[None]
    multi, single = Oracle('LOWW', n_env=N, n_ac=A), Oracle('LOWW', n_env=N * A, n_ac=1)
    multi.reset(spawn=spawn); single.reset(spawn=spawn.reshape(N * A, 1, 5))
    alive = np.ones(N, bool)
    for t in range(T):
        om, _, rm, dm, tm = multi.step(acts[t])
        os_, _, rs, ds, ts = single.step(acts[t].reshape(N * A, 1, 3))
        sel = np.repeat(alive, A)
        np.testing.assert_array_equal(om.reshape(N * A, 10)[sel], os_.reshape(N * A, 10)[sel])
        r4 = rs.reshape(N, A)
        np.testing.assert_array_equal(rm[alive], ((r4[:, 0] + r4[:, 1]) + (r4[:, 2] + r4[:, 3]))[alive])
        np.testing.assert_array_equal(dm[alive], ds.reshape(N, A).any(1)[alive])
        np.testing.assert_array_equal(((tm[:, None] >> (8 + 3 * np.arange(A))) & 7)[alive], (ts & 0xFF).reshape(N, A)[alive])
        alive &= ~dm.astype(bool)
    assert alive.sum() > 5


def test_oracle_zero_wind_equals_no_wind_and_wind_moves_aircraft():
    N, A, T = 128, 2, 80
    acts = _acts(np.random.RandomState(1), T, N, A)
    o0 = Oracle('LOWW', True, n_env=N, n_ac=A, seed=4)
    o1 = Oracle('LOWW', True, n_env=N, n_ac=A, seed=4, wind=np.zeros((3, 5, 2), np.float32))
    w = np.zeros((2, 2, 2), np.float32); w[..., 0] = 36.0            # 36 kt from the west = +0.01 nm/s in x
    o2 = Oracle('LOWW', True, n_env=N, n_ac=A, seed=4, wind=w)
    for o in (o0, o1, o2):
        o.reset()
    a, b = o0.rollout(acts), o1.rollout(acts)
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x, y)
    np.testing.assert_array_equal(o0.get_state()[0], o1.get_state()[0])
    s0 = o0.get_state()[0]
    o0b = Oracle('LOWW', True, n_env=N, n_ac=A, seed=4); o0b.reset()
    o0b.step(acts[0]); o2.step(acts[0])
    d = o2.get_state()[0] - o0b.get_state()[0]
    np.testing.assert_allclose(d[..., 0], 0.01, atol=1e-12)
    np.testing.assert_allclose(d[..., 1], 0.0, atol=1e-12)
    assert s0.shape == (N, A, 5)


def test_oracle_spawn_rng_properties():
    """Counter-based spawn: deterministic in (seed, global env index, episode), entry points without replacement,
    levels from the chosen entry point's list, sharding-invariant."""
    N, A = 512, 8
    o = Oracle('LOWW', True, n_env=N, n_ac=A, seed=123)
    o.reset()
    st = o.get_state()[0]
    eps = {(e[0], e[1]): (e[2], e[3]) for e in G.kat()['LOWW_random_entrypoints']}
    for e in range(N):
        xy = [tuple(v) for v in st[e, :, :2]]
        assert len(set(xy)) == A                                   # without replacement
        for a in range(A):
            phi, levels = eps[xy[a]]
            assert st[e, a, 3] == phi and st[e, a, 2] / 100 in levels and st[e, a, 4] == 250
    o2 = Oracle('LOWW', True, n_env=N, n_ac=A, seed=123); o2.reset()
    np.testing.assert_array_equal(o2.get_state()[0], st)
    o3 = Oracle('LOWW', True, n_env=N, n_ac=A, seed=124); o3.reset()
    assert (o3.get_state()[0] != st).any()
    part = Oracle('LOWW', True, n_env=100, n_ac=A, seed=123, env_index_base=200); part.reset()
    np.testing.assert_array_equal(part.get_state()[0], st[200:300])
    o.reset()                                                       # next episode index -> new draw
    assert (o.get_state()[0] != st).any()
    # every entry point and a spread of levels are used
    assert len({tuple(v) for v in st[:, :, :2].reshape(-1, 2)}) == 9
    # fewer entry points than aircraft: with replacement (documented), still valid states
    o1 = Oracle('LOWW', False, n_env=4, n_ac=3, seed=1); o1.reset()
    assert (o1.get_state()[0][..., 0] == 10).all()


def test_oracle_separation_rule():
    cases = [(2.9, 500.0, True), (3.5, 500.0, False), (2.9, 1001.0, False), (2.9, 998.0, True), (0.5, 1500.0, False)]
    N = len(cases)
    spawn = np.zeros((N, 2, 5))
    for i, (dx, dh, _) in enumerate(cases):
        spawn[i, 0] = [35.0, 45.0, 9000.0, 0.0, 200.0]
        spawn[i, 1] = [35.0 + dx, 45.0, 9000.0 + dh, 0.0, 200.0]
    o = Oracle('LOWW', n_env=N, n_ac=2)
    o.reset(spawn=spawn)
    acts = np.zeros((N, 2, 3), np.float32)
    acts[:, 0, 1] = 9000 / 19000 - 1
    acts[:, 1, 1] = [(9000 + c[1]) / 19000 - 1 for c in cases]
    acts[..., 2] = -1.0
    obs, raw, rew, done, term = o.step(acts)
    np.testing.assert_array_equal((term & 0xFF) == 5, [c[2] for c in cases])
    np.testing.assert_array_equal(done.astype(bool), [c[2] for c in cases])
    assert (rew[(term & 0xFF) == 5] < -190).all() and (rew[(term & 0xFF) == 0] > -1).all()


def test_oracle_rollout_raw_equals_stepwise():
    """rollout(raw=True) is T x step(autoreset=True): same obs / reward / flags, and info["original_state"] is the raw
    observation of the moved aircraft on every row — on the terminal rows of auto-reset envs too, where `obs` already
    holds the reset observation (atc_gym.py:192 returns the pre-reset state)."""
    N, A, T = 64, 2, 400
    acts = _acts(np.random.RandomState(5), T, N, A)
    o1 = Oracle('LOWW', True, n_env=N, n_ac=A, seed=9); o1.reset()
    o2 = Oracle('LOWW', True, n_env=N, n_ac=A, seed=9); o2.reset()
    obs, raw, rew, done, term = o1.rollout(acts, raw=True)
    assert done.sum() > 10
    for t in range(T):
        s_obs, s_raw, s_rew, s_done, s_term = o2.step(acts[t], autoreset=True)
        np.testing.assert_array_equal(obs[t], s_obs)
        np.testing.assert_array_equal(raw[t], s_raw)
        np.testing.assert_array_equal(rew[t], s_rew)
        np.testing.assert_array_equal(done[t], s_done)
    # on a terminal row the two observations differ (reset obs vs the terminal state); elsewhere obs = normalise(raw)
    t, e = np.argwhere(done)[0]
    assert not np.allclose(obs[t, e], raw[t, e])
    c = o1.constants()
    live = ~done.astype(bool)
    norm = (raw[live] - c['nmin'].astype(np.float32) - 0.5 * c['nmax'].astype(np.float32)) / (0.5 * c['nmax'].astype(np.float32))
    np.testing.assert_allclose(obs[live], norm, rtol=1e-6, atol=1e-6)


def test_oracle_autoreset_and_metrics():
    N, A, T = 256, 1, 700
    o = Oracle('LOWW', n_env=N, n_ac=A)
    o.reset()
    obs, rew, done, term = o.rollout(_acts(np.random.RandomState(2), T, N, A))
    m = o.metrics()
    assert done.sum() > 50 and (m['episodes'] == 1 + done.sum(0)).all()
    # the observation of a finished env is the RAW reset observation (reference quirk, atc_gym.py:351,365)
    t, e = np.argwhere(done)[0]
    np.testing.assert_allclose(obs[t, e, 0], G.kat()['K0_reset_obs'], rtol=1e-7)
    # episode return bookkeeping: last_ep_return = sum of the rewards of the last finished episode
    for e in np.nonzero(done.sum(0) >= 2)[0][:20]:
        ts = np.nonzero(done[:, e])[0]
        np.testing.assert_allclose(m['last_ep_return'][e], rew[ts[-2] + 1:ts[-1] + 1, e].sum(), rtol=1e-12)
        assert m['last_ep_len'][e] == ts[-1] - ts[-2]


@pytest.mark.parametrize('scn', ['LOWW', 'SimpleScenario'])
def test_compact_grid_is_exact(scn):
    """sector.CompactGrid (the coarse copy of the MVA accelerator the one-CTA-per-SM rollout kernel keeps in shared
    memory): its lookup + fine-grid fallback equals the reference scan on random, on-edge, near-vertex and
    cell-border points, for the point's own cell and for every neighbour within the float32 margin."""
    import atc_reinforcement_learning_b200 as P
    from atc_reinforcement_learning_b200 import _native as nat
    from atc_reinforcement_learning_b200.sector import CompiledSector, build_compact_grid
    s = getattr(P, scn)(random_entrypoints=(scn == 'LOWW'))
    fine = CompiledSector(s, cell=0.25)
    budget = int(nat.lib().atc_compact_grid_budget())
    assert 100 * 1024 < budget < 227 * 1024
    cg = build_compact_grid(s, budget)
    assert cg is not None and cg.nbytes <= budget and cg.n_lines <= 126 and cg.n_blocks <= 511
    assert build_compact_grid(s, 1024) is None                      # nothing fits
    rng = np.random.RandomState(4)
    b = fine.bbox
    n = 200000
    xs = [rng.uniform(b[0] - 2, b[2] + 2, n)]; ys = [rng.uniform(b[1] - 2, b[3] + 2, n)]
    for ring in fine.rings:                                          # on / next to the edges and vertices
        for i in range(1, len(ring)):
            t = rng.uniform(0, 1, 200)
            ex = ring[i - 1][0] + t * (ring[i][0] - ring[i - 1][0]); ey = ring[i - 1][1] + t * (ring[i][1] - ring[i - 1][1])
            for d in (0.0, 1e-12, -1e-12, 1e-9, -1e-9, 3e-9, -3e-9, 1e-6, -1e-6, 1e-3):
                xs.append(ex + d); ys.append(ey - d)
    k = rng.randint(0, 8 * cg.grid_nx, 40000); l = rng.randint(0, 8 * cg.grid_ny, 40000)   # (sub-)cell borders, corners
    for d in (0.0, 1e-9, -1e-9, 1e-6, -1e-6):
        xs.append(cg.grid_x0 + k * cg.cell / 8 + d); ys.append(cg.grid_y0 + l * cg.cell / 8 + rng.uniform(0, cg.cell, 40000))
        xs.append(cg.grid_x0 + k * cg.cell / 8 + d); ys.append(cg.grid_y0 + l * cg.cell / 8 - d)
    x, y = np.concatenate(xs), np.concatenate(ys)
    ref = fine.find_mva_np(x, y)
    got, slow = cg.lookup_np(x, y, fine)
    np.testing.assert_array_equal(got, ref)
    assert slow[:n].mean() < 0.01                                    # random points hardly ever need the fine grid
    # the float32 index may pick the neighbour of the true sub-cell when the point is within the float32 error of a
    # border: every (sub-)cell's entry holds on the cell grown by `margin`, so forcing the neighbour changes nothing
    sc = cg.cell / 8
    ix8, iy8 = cg.cell_index_np(x, y)
    fx, fy = (x - cg.grid_x0) / sc, (y - cg.grid_y0) / sc
    for dx, dy in ((1, 0), (-1, 0), (0, 1), (0, -1)):
        border = (np.floor(fx) + (dx > 0)) if dx else (np.floor(fy) + (dy > 0))
        dist = np.abs((fx if dx else fy) - border) * sc
        jx, jy = ix8 + dx, iy8 + dy
        sel = (dist < cg.margin) & (jx >= 0) & (jy >= 0) & (jx < 8 * cg.grid_nx) & (jy < 8 * cg.grid_ny) & \
            (ix8 == np.floor(fx)) & (iy8 == np.floor(fy))
        if sel.any():
            g2, _ = cg.lookup_np(x[sel], y[sel], fine, cells=(jx[sel], jy[sel]))
            np.testing.assert_array_equal(g2, ref[sel])


def test_row_stores_stay_adjacent_in_sass():
    """The five 8-byte stores of an observation row must be issued (nearly) back to back in the rollout kernel the bench
    times: the L1 merges the partial-sector writes of adjacent store instructions only, and a build whose scheduler
    spreads them between the arithmetic runs 12 % slower at 1.9 x instead of 1.4 x the payload in L1 -> L2 writes
    (found in round 2 when an unrelated change flipped the schedule).  No GPU needed: cuobjdump."""
    import importlib.util
    import shutil
    if shutil.which('cuobjdump') is None:
        pytest.skip('cuobjdump not available')
    spec = importlib.util.spec_from_file_location('sass_summary', os.path.join(ROOT, 'tools', 'sass_summary.py'))
    ss = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ss)
    from atc_reinforcement_learning_b200 import _native as nat
    nat.build_library()
    fns = ss.functions()
    checked = 0
    for name, rows in fns.items():
        if 'atc_rollout_pipe_kernel<4, false, false, 2, 14>' in name or 'atc_rollout_pipe_kernel<4, false, false, 1, 14>' in name:
            spans = ss.row_store_spans(rows)
            assert spans, name
            assert max(spans) <= 10, (name, spans)
            checked += 1
    assert checked == 2


def test_vecnorm_launch_plan_invariants():
    """The fused VecNormalize launch (csrc/atc_vecnorm.cu) cuts the T x rows observation rows into (step, slab) tiles;
    its geometry is a pure function of the machine and the job (atc_vecnorm_plan, no GPU): every column is covered, the
    slabs start at multiples of 5 columns (a thread keeps its features), no CTA owns more tiles than it has slots for,
    16-byte loads only when every step starts 16-byte aligned, scratch = what the kernel indexes."""
    from atc_reinforcement_learning_b200 import _native as nat
    L = nat.lib()
    out = (C.c_int64 * 6)()
    rng = np.random.RandomState(0)
    cases = [(148, 4, 1, 16384, 4), (148, 4, 1024, 16384, 4), (148, 4, 40, 1, 1), (148, 4, 130, 300, 3), (148, 1, 4736, 7, 1),
             (132, 2, 2048, 131072, 8), (8, 1, 256, 3, 5)]
    for _ in range(400):
        cases.append((int(rng.randint(1, 200)), int(rng.randint(1, 9)), int(rng.randint(1, 600)),
                      int(2 ** rng.uniform(0, 18)), int(rng.randint(1, 9))))
    n_ok = 0
    for n_sm, per_sm, T, N, A in cases:
        rc = L.atc_vecnorm_plan(n_sm, per_sm, T, N, A, out)
        if T > 32 * n_sm * min(per_sm, 4):
            assert rc != 0
            continue
        assert rc == 0, (n_sm, per_sm, T, N, A, L.atc_vecnorm_last_error())
        grid, slabs, slab_cols, ret_ctas, vec, scratch = [int(v) for v in out]
        rows = N * A
        n_cols = rows * 10 // vec
        assert vec == (4 if rows % 2 == 0 else 2) and rows * 10 % vec == 0
        assert 1 <= grid <= n_sm * min(per_sm, 4)
        assert slabs >= 1 and slab_cols % 5 == 0 and slabs * slab_cols >= n_cols
        assert T * slabs <= 32 * grid                               # kMaxOwned tiles per CTA
        assert 1 <= ret_ctas <= grid and ret_ctas <= (N + 255) // 256
        assert scratch == T * (slabs * 20 + ret_ctas * 3)
        n_ok += 1
    assert n_ok > 300
    # one step of the bench batch: about one tile per CTA, a fraction of the machine (latency-bound launch)
    assert L.atc_vecnorm_plan(148, 4, 1, 16384, 4, out) == 0 and int(out[0]) == int(out[1]) == 81
    # a 1024-step rollout of it: every co-resident CTA, the slab count that balances the tiles
    assert L.atc_vecnorm_plan(148, 4, 1024, 16384, 4, out) == 0 and int(out[0]) == 592 and (1024 * int(out[1])) % 592 < 592
