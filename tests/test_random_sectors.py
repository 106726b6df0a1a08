"""Sectors other than the reference's two: the sector compiler (grid, per-cell programs, single-line records) and the
kernels must be exact on ANY polygon map, not just LOWW.  Random jittered quad meshes with holes, shuffled list order
and mixed ring orientations; the oracle's brute-force first-match scan is the checker."""
import json

import numpy as np
import pytest

from oracle.oracle import Oracle
from tests.sector_gen import random_sector


def _usable(doc, tmp_path):
    import atc_reinforcement_learning_b200 as P
    p = tmp_path / (doc['name'] + '.json')
    p.write_text(json.dumps(doc))
    try:
        cs = P.CompiledSector(P.load_scenario(str(p), random_entrypoints=True), cell=0.125)
    except ValueError:
        return None, None          # the random runway put the FAF into a hole: like the reference, that is an error
    return str(p), cs


@pytest.mark.parametrize('seed', range(8))
def test_random_sector_grid_is_exact(seed, tmp_path):
    doc = random_sector(seed)
    path, cs = _usable(doc, tmp_path)
    if cs is None:
        pytest.skip('FAF outside the airspace')
    ora = Oracle(path, random_entrypoints=True)
    rng = np.random.RandomState(seed)
    n = 150000
    pts = [np.stack([rng.uniform(cs.bbox[0] - 1, cs.bbox[2] + 1, n), rng.uniform(cs.bbox[1] - 1, cs.bbox[3] + 1, n)], 1)]
    for ring in cs.rings:
        for i in range(1, len(ring)):
            t = rng.uniform(0, 1, 300)[:, None]
            on = ring[i - 1] * (1 - t) + ring[i] * t
            nrm = np.array([ring[i][1] - ring[i - 1][1], ring[i - 1][0] - ring[i][0]])
            nrm = nrm / (np.linalg.norm(nrm) + 1e-300)
            for d in (0.0, 1e-13, -1e-13, 1e-10, -1e-10, 5e-10, 2e-9, -2e-9, 1e-6, -1e-6):
                pts.append(on + d * nrm)
        pts.append(ring)                                       # the vertices themselves
    pts = np.concatenate(pts, 0)
    np.testing.assert_array_equal(cs.lookup_np(pts[:, 0], pts[:, 1]), ora.mva_index(pts))
    assert cs.line_fraction > 0.5


@pytest.mark.gpu
@pytest.mark.parametrize('seed', [1, 5])
def test_random_sector_cuda_vs_oracle(seed, tmp_path):
    import torch
    import atc_reinforcement_learning_b200 as P
    doc = random_sector(seed)
    path, cs = _usable(doc, tmp_path)
    if cs is None:
        pytest.skip('FAF outside the airspace')
    N, A, T = 1024, 4, 200
    scn = P.load_scenario(path, random_entrypoints=True)
    env = P.BatchedAtcEnv(N, A, P.SimParameters(1), scn, seed=seed, grid_cell=0.125)
    ora = Oracle(path, random_entrypoints=True, n_env=N, n_ac=A, seed=seed)
    ora.reset(); ora.reset()
    env.reset()
    rng = np.random.RandomState(seed)
    acts = np.repeat(rng.uniform(-1, 1, (T // 20, N, A, 3)).astype(np.float32), 20, 0)
    o_obs, o_rew, o_done, o_term = ora.rollout(acts)
    obs, rew, done, info = env.rollout(torch.from_numpy(acts).cuda())
    np.testing.assert_array_equal(done.cpu().numpy().astype(np.uint8), o_done)
    np.testing.assert_array_equal(info['term_code'].cpu().numpy(), o_term)
    np.testing.assert_allclose(obs.cpu().numpy(), o_obs, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(rew.cpu().numpy(), o_rew, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(env.get_state()[0].cpu().numpy(), ora.get_state()[0], rtol=0, atol=1e-9)
    pts = np.stack([rng.uniform(cs.bbox[0] - 1, cs.bbox[2] + 1, 200000), rng.uniform(cs.bbox[1] - 1, cs.bbox[3] + 1, 200000)], 1)
    np.testing.assert_array_equal(env.query_mva(pts).cpu().numpy(), ora.mva(pts))
    assert o_done.sum() > 50
