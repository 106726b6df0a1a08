"""GPU parity tests (run on the B200 with -m gpu): the CUDA path, called through the C ABI, against
  (1) golden vectors recorded from the live reference (tests/golden/),
  (2) the CPU oracle (oracle/atc_oracle.c) on the same seeded inputs,
  (3) size-independent properties at BASELINE.json's full sizes.
Tolerances: integer outputs (done, term codes, timesteps, counters) bit-exact; float64 aircraft state within 1e-9
(libm differences only); float32 observation / reward within 1e-5 + 1e-5*|ref| (north_star), in practice far tighter.
"""
import os

import numpy as np
import pytest
import torch

from tests import golden_util as G

pytestmark = pytest.mark.gpu

OBS_RTOL, OBS_ATOL = 1e-5, 1e-5


def make_env(*a, **kw):
    from atc_reinforcement_learning_b200 import BatchedAtcEnv
    return BatchedAtcEnv(*a, **kw)


def scenario(name, random_entrypoints=False):
    import atc_reinforcement_learning_b200 as P
    return getattr(P, name)(random_entrypoints=random_entrypoints)


class CudaImpl(object):
    """golden_util.replay adaptor over BatchedAtcEnv (device tensors in, numpy out)."""

    def __init__(self, tr, exact_math=False):
        from atc_reinforcement_learning_b200 import SimParameters
        m = tr['meta']
        E = tr['action'].shape[1]
        sp = SimParameters(m['dt'], reward_shaping=m['reward_shaping'], normalize_state=m['normalize_state'],
                           discrete_action_space=m['discrete'])
        self.env = make_env(E, 1, sp, scenario(m['scenario'], m['random_entrypoints']), autoreset=False,
                            track_actions=True, exact_math=exact_math)

    def reset(self, mask, spawn):
        return self.env.reset(mask, spawn).cpu().numpy()

    def set_state(self, st, ts):
        self.env.set_state(st, ts)

    def step(self, a):
        obs, rew, done, info = self.env.step(torch.from_numpy(np.ascontiguousarray(a)).cuda())
        return (obs.cpu().numpy(), info['original_state'].cpu().numpy(), rew.double().cpu().numpy(),
                done.cpu().numpy(), info['term_code'].cpu().numpy())

    def get_state(self):
        st, ts = self.env.get_state()
        return st.cpu().numpy(), ts.cpu().numpy()

    def metrics(self):
        e = self.env
        return {'actions_taken': e.actions_taken.cpu().numpy(), 'ep_return': e.ep_return.cpu().numpy(),
                'win_ring': e.win_ring.cpu().numpy()}


@pytest.mark.parametrize('name', G.trace_names())
def test_cuda_replays_reference_trace(name):
    """Step-for-step against the live reference's recorded traces (every terminal branch, discrete actions, dt=5...),
    default arithmetic: float64 state/decisions, float32 observation/shaping.  Flags exact, state 1e-9, observation
    and reward within north_star's 1e-5 (+1e-5 relative)."""
    tr = G.load_trace(name)
    worst = G.replay(tr, CudaImpl(tr), state_atol=1e-9, reward_rtol=1e-5, reward_atol=1e-5, obs_rtol=OBS_RTOL,
                     obs_atol=OBS_ATOL, return_rtol=1e-5, return_atol=1e-3)
    print(name, worst)
    assert worst['state'] <= 1e-9


@pytest.mark.parametrize('name', G.trace_names())
def test_cuda_exact_math_replays_reference_trace(name):
    """exact_math=True: float64 + libm throughout; the only float32 rounding left is the returned reward (6e-8)."""
    tr = G.load_trace(name)
    worst = G.replay(tr, CudaImpl(tr, exact_math=True), state_atol=1e-9, reward_rtol=2e-7, reward_atol=1e-9,
                     obs_rtol=2e-7, obs_atol=1e-7, return_rtol=1e-9, return_atol=1e-9)
    print(name, worst)
    assert worst['state'] <= 1e-9


@pytest.mark.parametrize('scn', ['LOWW', 'SimpleScenario'])
def test_cuda_geometry_matches_reference(scn):
    z = np.load(os.path.join(G.GOLDEN, 'geometry_%s.npz' % scn))
    env = make_env(1, 1, None, scenario(scn), autoreset=False)
    np.testing.assert_array_equal(env.query_mva(z['pts']).cpu().numpy(), z['mva'])
    np.testing.assert_array_equal(env.query_corridor(z['corr']).cpu().numpy().astype(np.uint8), z['inside'])


def test_cuda_kat_corridor_gates():
    k = G.kat()
    env = make_env(1, 1, None, scenario('LOWW'), autoreset=False)
    rows = np.asarray([r[:4] for r in k['K7_corridor']], np.float64)
    exp = np.asarray([r[4] for r in k['K7_corridor']], bool)
    np.testing.assert_array_equal(env.query_corridor(rows).cpu().numpy(), exp)


def test_reference_unit_tests_on_cuda(tmp_path):
    """The reference's own 8 unit tests (envs/atc/model_test.py:10-92) against the CUDA geometry."""
    import json
    from atc_reinforcement_learning_b200 import load_scenario
    from atc_reinforcement_learning_b200.scenarios import SECTOR_DIR
    k = G.kat()['model_test']
    with open(os.path.join(SECTOR_DIR, 'SimpleScenario.json')) as f:
        doc = json.load(f)
    doc['runway'] = {'x': 20, 'y': 20, 'h': 0, 'phi_from_runway': 180}
    p = tmp_path / 'unit.json'
    p.write_text(json.dumps(doc))
    env = make_env(1, 1, None, load_scenario(str(p)), autoreset=False)
    assert int(env.query_mva([[34, 1]])[0]) == 3500
    assert int(env.query_mva([list(env.sector.faf)])[0]) == k['faf_mva']
    for x, y, h, phi, exp in k['inside_corridor']:
        assert bool(env.query_corridor([[x, y, h, phi]])[0]) == exp
    # _inside_corridor_angle alone: put the aircraft low enough that only the angle gate decides
    for x, y, phi, exp in k['inside_corridor_angle']:
        assert bool(env.query_corridor([[x, y, 0.0, phi]])[0]) == exp
    assert env.sector.bbox.tolist() == [0.0, 0.0, 35.0, 40.0]


# ------------------------------------------------------------------------------------------------ CUDA vs oracle
def run_pair(N, A, T, seed, sector='LOWW', random_entrypoints=True, wind=None, dt=1.0, chunk=None, amp=1.0,
             repeat=20, normalize_reset_obs=False, h_bias=False, spawn=None, exact_math=False):
    from atc_reinforcement_learning_b200 import SimParameters
    from oracle.oracle import Oracle
    env = make_env(N, A, SimParameters(dt), scenario(sector, random_entrypoints), seed=seed, wind=wind,
                   track_actions=True, normalize_reset_obs=normalize_reset_obs, exact_math=exact_math)
    ora = Oracle(sector, random_entrypoints, n_env=N, n_ac=A, dt=dt, seed=seed, wind=wind,
                 normalize_reset_obs=normalize_reset_obs)
    ora.reset()                       # the env constructor resets once (atc_gym.py:61)
    o0 = ora.reset()
    g0 = env.reset().cpu().numpy()
    np.testing.assert_allclose(g0, o0, rtol=1e-6, atol=1e-6)
    if spawn is not None:
        np.testing.assert_allclose(env.reset(spawn=spawn).cpu().numpy(), ora.reset(spawn=spawn), rtol=1e-6, atol=1e-6)
    rng = np.random.RandomState(seed)
    acts = np.repeat(rng.uniform(-amp, amp, ((T + repeat - 1) // repeat, N, A, 3)).astype(np.float32), repeat, 0)[:T]
    if h_bias:
        acts[..., 1] = -np.abs(acts[..., 1]) * 0.3 - 0.7
    o_obs, o_raw, o_rew, o_done, o_term = ora.rollout(acts, raw=True)
    chunk = chunk or T
    outs = []
    for s in range(0, T, chunk):
        obs, rew, done, info = env.rollout(torch.from_numpy(acts[s:s + chunk]).cuda())
        outs.append((obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy(), info['term_code'].cpu().numpy(),
                     info['original_state'].cpu().numpy()))
    g_obs, g_rew, g_done, g_term, g_raw = (np.concatenate([o[k] for o in outs], 0) for k in range(5))
    np.testing.assert_array_equal(g_done.astype(np.uint8), o_done)
    np.testing.assert_array_equal(g_term, o_term)
    np.testing.assert_allclose(g_rew, o_rew, rtol=OBS_RTOL, atol=OBS_ATOL)
    np.testing.assert_allclose(g_obs, o_obs, rtol=OBS_RTOL, atol=OBS_ATOL)
    # info["original_state"] (atc_gym.py:192): the raw observation of the MOVED aircraft on every row, the terminal
    # rows of auto-reset envs included (there `obs` holds the reset observation instead)
    np.testing.assert_allclose(g_raw, o_raw, rtol=OBS_RTOL, atol=OBS_ATOL)
    st, ts = env.get_state()
    ost, ots = ora.get_state()
    np.testing.assert_array_equal(ts.cpu().numpy(), ots)
    np.testing.assert_allclose(st.cpu().numpy(), ost, rtol=0, atol=1e-9)
    m = ora.metrics()
    np.testing.assert_array_equal(env.episodes.cpu().numpy(), m['episodes'])
    np.testing.assert_array_equal(env.win_ring.cpu().numpy(), m['win_ring'])
    np.testing.assert_array_equal(env.last_ep_len.cpu().numpy(), m['last_ep_len'])
    np.testing.assert_array_equal(env.actions_taken.cpu().numpy(), m['actions_taken'])
    rt, at = (1e-9, 1e-9) if exact_math else (1e-5, 1e-3)
    np.testing.assert_allclose(env.ep_return.cpu().numpy(), m['ep_return'], rtol=rt, atol=at)
    np.testing.assert_allclose(env.last_ep_return.cpu().numpy(), m['last_ep_return'], rtol=rt, atol=at)
    codes = np.bincount((o_term & 0xFF)[o_done > 0], minlength=6)
    return {'dones': int(o_done.sum()), 'codes': codes.tolist(),
            'max_obs_err': float(np.abs(g_obs - o_obs).max()), 'max_rew_err': float(np.abs(g_rew - o_rew).max()),
            'max_raw_err': float(np.abs(g_raw - o_raw).max()), 'last_launch': env.last_launch}


def test_config1_single_env():
    r = run_pair(1, 1, 1500, seed=1, random_entrypoints=False)
    assert r['dones'] >= 1


def test_config2_4096x1():
    r = run_pair(4096, 1, 600, seed=2, chunk=200)
    print(r)
    assert r['codes'][2] > 100


def test_config3_16384x4_separation():
    r = run_pair(16384, 4, 160, seed=3, chunk=80)
    print(r)
    assert r['codes'][2] > 1000


def test_bench_shape_16384x4_T1024_vs_oracle():
    """The launch bench.py times — 16384 envs x 4 aircraft, ONE rollout launch of 1024 steps, default layout selection
    (one CTA per SM, compact MVA grid in shared memory) — against the oracle on every row of every env: obs,
    info["original_state"], reward bit-for-tolerance, done / term bit-exact, final state, counters.  The oracle
    replays the batch in blocks of 2048 envs (env_index_base) to bound host memory."""
    from atc_reinforcement_learning_b200 import SimParameters
    from oracle.oracle import Oracle
    N, A, T, seed, blk = 16384, 4, 1024, 11, 2048
    env = make_env(N, A, SimParameters(1.0), scenario('LOWW', True), seed=seed)
    g = torch.Generator(device='cuda').manual_seed(1234)
    acts = (torch.rand((T + 19) // 20, N, A, 3, device='cuda', generator=g) * 2 - 1).repeat_interleave(20, 0)[:T]
    acts = acts.contiguous()
    obs, rew, done, info = env.rollout(acts)
    torch.cuda.synchronize()
    ll = env.last_launch
    print(ll)
    assert ll['kernel'] == 3 and ll['n_steps'] == T and ll['pairs_per_cta'] == 14, ll
    st, ts = env.get_state()
    worst = {'obs': 0.0, 'raw': 0.0, 'rew': 0.0}
    n_done = 0
    for b0 in range(0, N, blk):
        ora = Oracle('LOWW', True, n_env=blk, n_ac=A, seed=seed, env_index_base=b0)
        ora.reset()                   # the env constructor resets once (atc_gym.py:61) and nothing else has run
        a = acts[:, b0:b0 + blk].cpu().numpy()
        o_obs, o_raw, o_rew, o_done, o_term = ora.rollout(a, raw=True)
        np.testing.assert_array_equal(done[:, b0:b0 + blk].cpu().numpy().astype(np.uint8), o_done)
        np.testing.assert_array_equal(info['term_code'][:, b0:b0 + blk].cpu().numpy(), o_term)
        for name, gpu, ref in (('obs', obs, o_obs), ('raw', info['original_state'], o_raw), ('rew', rew, o_rew)):
            gv = gpu[:, b0:b0 + blk].cpu().numpy()
            np.testing.assert_allclose(gv, ref, rtol=OBS_RTOL, atol=OBS_ATOL, err_msg=name)
            worst[name] = max(worst[name], float(np.abs(gv - ref).max()))
        ost, ots = ora.get_state()
        np.testing.assert_array_equal(ts[b0:b0 + blk].cpu().numpy(), ots)
        np.testing.assert_allclose(st[b0:b0 + blk].cpu().numpy(), ost, rtol=0, atol=1e-9)
        m = ora.metrics()
        np.testing.assert_array_equal(env.episodes[b0:b0 + blk].cpu().numpy(), m['episodes'])
        np.testing.assert_array_equal(env.last_ep_len[b0:b0 + blk].cpu().numpy(), m['last_ep_len'])
        n_done += int(o_done.sum())
    print(worst, 'episodes finished per env: %.2f' % (n_done / N))
    assert n_done >= 2 * N          # every env has been through auto-reset about twice or more


def test_dense_traffic_separation_vs_oracle():
    """Random spawns rarely conflict; pack the aircraft into a 20 x 20 nm box so the separation branch fires a lot."""
    N, A = 4096, 4
    rng = np.random.RandomState(8)
    spawn = np.zeros((N, A, 5))
    spawn[..., 0] = rng.uniform(30, 50, (N, A)); spawn[..., 1] = rng.uniform(38, 55, (N, A))
    spawn[..., 2] = rng.uniform(8000, 11000, (N, A)); spawn[..., 3] = rng.uniform(0, 360, (N, A)); spawn[..., 4] = 250
    r = run_pair(N, A, 150, seed=6, spawn=spawn)
    print(r)
    assert r['codes'][5] > 500
    r = run_pair(1024, 8, 100, seed=7, spawn=np.concatenate([spawn[:1024], spawn[1024:2048]], 1))
    print(r)
    assert r['codes'][5] > 200


def test_config4_16384x8_wind():
    rng = np.random.RandomState(99)
    wind = rng.uniform(-30, 30, (16, 16, 2)).astype(np.float32)
    r = run_pair(16384, 8, 64, seed=4, wind=wind)
    print(r)
    assert r['dones'] > 100


def test_exact_math_vs_oracle_full_precision():
    r = run_pair(4096, 4, 120, seed=41, exact_math=True)
    print(r)
    assert r['max_rew_err'] < 1e-5 and r['max_obs_err'] <= 4e-3      # reward: float32 rounding of the float64 sum


def test_fast_normalisation_stays_within_float32_rounding():
    """The default path normalises with one FFMA per slot (v * 1/half - (min + half)/half) instead of the reference's
    two subtractions and a division (atc_gym.py:187-189).  For the slots whose raw value is identical in both modes
    the result must stay within a few float32 roundings of the exact_math path (values are in [-1, 1])."""
    from atc_reinforcement_learning_b200 import SimParameters
    N, A, T = 16384, 4, 64
    g = torch.Generator(device='cuda').manual_seed(77)
    acts = (torch.rand(T, N, A, 3, device='cuda', generator=g) * 2 - 1)
    ef = make_env(N, A, SimParameters(1), scenario('LOWW', True), seed=9)
    ee = make_env(N, A, SimParameters(1), scenario('LOWW', True), seed=9, exact_math=True)
    of, oe = ef.rollout(acts), ee.rollout(acts)
    same_raw = [0, 1, 2, 3, 4, 5, 9]                      # casts of the float64 state: identical raw values
    assert torch.equal(of[3]['original_state'][..., same_raw], oe[3]['original_state'][..., same_raw])
    # rows of freshly reset envs hold raw (un-normalised) values, identical in both modes; the others are in [-1, 1]
    err = (of[0][..., same_raw].double() - oe[0][..., same_raw].double()).abs()
    assert float(err.max()) <= 4e-7, float(err.max())
    assert torch.equal(of[2], oe[2]) and torch.equal(of[3]['term_code'], oe[3]['term_code'])
    assert torch.equal(ef.state, ee.state)
    torch.testing.assert_close(of[0], oe[0], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(of[1], oe[1], rtol=1e-5, atol=1e-5)


def test_odd_aircraft_counts_and_ragged_batches():
    for N, A in ((1, 2), (7, 3), (33, 5), (129, 6), (5, 7), (1000, 8)):
        run_pair(N, A, 120, seed=10 + A)


def test_descending_traffic_hits_below_mva():
    r = run_pair(2048, 2, 500, seed=21, h_bias=True, chunk=125)
    print(r)
    assert r['codes'][1] > 20


def test_timeout_branch_batched():
    """phi command held at the current heading circles nothing: fly a racetrack that stays inside until t > 6000."""
    from atc_reinforcement_learning_b200 import SimParameters
    from oracle.oracle import Oracle
    N = 8
    env = make_env(N, 1, SimParameters(1), scenario('LOWW'), seed=0)
    ora = Oracle('LOWW', n_env=N, n_ac=1, seed=0)
    ora.reset(); ora.reset(); env.reset()
    st = np.tile(np.array([40.0, 45.0, 9000.0, 0.0, 150.0]), (N, 1, 1))
    ts = np.full(N, 5990, np.int32)
    env.set_state(st, ts); ora.set_state(st, ts)
    t = np.arange(40)
    acts = np.zeros((40, N, 1, 3), np.float32)
    acts[..., 0] = -0.5
    acts[..., 1] = 9000 / 19000 - 1
    acts[..., 2] = ((t * 3.0 % 360) / 180 - 1)[:, None, None]
    o_obs, o_rew, o_done, o_term = ora.rollout(acts)
    obs, rew, done, info = env.rollout(torch.from_numpy(acts).cuda())
    np.testing.assert_array_equal(done.cpu().numpy().astype(np.uint8), o_done)
    np.testing.assert_array_equal(info['term_code'].cpu().numpy(), o_term)
    assert ((o_term[10] & 0xFF) == 4).all() and o_done[9].sum() == 0       # t = 6001 is the first timeout step
    np.testing.assert_allclose(rew.cpu().numpy(), o_rew, rtol=OBS_RTOL, atol=OBS_ATOL)


def test_normalized_reset_obs_option_and_dt():
    run_pair(512, 4, 200, seed=31, normalize_reset_obs=True, dt=2.0)


# ------------------------------------------------------------------------------------------------ properties at full size
def _rollout_all(env, acts, chunk):
    outs = [env.rollout(acts[s:s + chunk]) for s in range(0, acts.shape[0], chunk)]
    return (torch.cat([o[0] for o in outs]), torch.cat([o[1] for o in outs]), torch.cat([o[2] for o in outs]),
            torch.cat([o[3]['term_code'] for o in outs]))


def test_full_size_step_equals_rollout_and_is_deterministic():
    """BASELINE config 3 size.  rollout(T) == T x step() bit for bit; two runs from the same seed are identical."""
    from atc_reinforcement_learning_b200 import SimParameters
    N, A, T = 16384, 4, 48
    g = torch.Generator(device='cuda').manual_seed(1234)
    acts = (torch.rand(T, N, A, 3, device='cuda', generator=g) * 2 - 1)
    e1 = make_env(N, A, SimParameters(1), scenario('LOWW', True), seed=5)
    e2 = make_env(N, A, SimParameters(1), scenario('LOWW', True), seed=5)
    e3 = make_env(N, A, SimParameters(1), scenario('LOWW', True), seed=5)
    rng = np.random.RandomState(8)                       # dense traffic: separation resets inside the window
    spawn = np.zeros((N, A, 5))
    spawn[..., 0] = rng.uniform(30, 50, (N, A)); spawn[..., 1] = rng.uniform(38, 55, (N, A))
    spawn[..., 2] = rng.uniform(8000, 11000, (N, A)); spawn[..., 3] = rng.uniform(0, 360, (N, A)); spawn[..., 4] = 250
    for e in (e1, e2, e3):
        e.reset(spawn=spawn)
    o1, r1, d1, t1 = _rollout_all(e1, acts, T)
    o2, r2, d2, t2 = _rollout_all(e2, acts, 7)
    steps = [e3.step(acts[t]) for t in range(T)]
    o3 = torch.stack([s[0] for s in steps]); r3 = torch.stack([s[1] for s in steps])
    d3 = torch.stack([s[2] for s in steps]); t3 = torch.stack([s[3]['term_code'] for s in steps])
    for a, b in ((o1, o2), (r1, r2), (d1, d2), (t1, t2), (o1, o3), (r1, r3), (d1, d3), (t1, t3)):
        assert torch.equal(a, b)
    assert torch.equal(e1.state, e2.state) and torch.equal(e1.state, e3.state)
    assert torch.equal(e1.ep_return, e3.ep_return) and torch.equal(e1.episodes, e3.episodes)
    assert int(d1.sum()) > 1000


def test_sharding_is_a_pure_partition():
    """config 5 logic: a rank owning envs [base, base + n) reproduces exactly that slice of the single-GPU job."""
    from atc_reinforcement_learning_b200 import SimParameters
    from atc_reinforcement_learning_b200.dist import shard_envs
    N, A, T = 4096, 4, 64
    g = torch.Generator(device='cuda').manual_seed(7)
    acts = (torch.rand(T, N, A, 3, device='cuda', generator=g) * 2 - 1)
    full = make_env(N, A, SimParameters(1), scenario('LOWW', True), seed=11)
    of, rf, df, tf = _rollout_all(full, acts, T)
    for rank in range(4):
        n, base = shard_envs(N, rank, 4)
        part = make_env(n, A, SimParameters(1), scenario('LOWW', True), seed=11, env_index_base=base)
        op, rp, dp, tp = _rollout_all(part, acts[:, base:base + n].contiguous(), T)
        assert torch.equal(op, of[:, base:base + n]) and torch.equal(rp, rf[:, base:base + n])
        assert torch.equal(dp, df[:, base:base + n]) and torch.equal(tp, tf[:, base:base + n])


def test_multi_aircraft_degenerates_to_independent_single_aircraft_envs():
    """A = 4 with aircraft kept > 3 nm / 1000 ft apart == 4 independent reference-semantics envs (reward = sum)."""
    from atc_reinforcement_learning_b200 import SimParameters
    N, A, T = 256, 4, 200
    rng = np.random.RandomState(3)
    spawn = np.zeros((N, A, 5))
    spawn[..., 0] = rng.uniform(25, 45, (N, A)); spawn[..., 1] = rng.uniform(35, 55, (N, A))
    spawn[..., 2] = 8000 + 2000 * np.arange(A)[None, :]            # 2000 ft apart -> never a separation violation
    spawn[..., 3] = rng.uniform(0, 360, (N, A)); spawn[..., 4] = 250
    acts = rng.uniform(-1, 1, (T, N, A, 3)).astype(np.float32)
    acts[..., 1] = (spawn[..., 2] / 19000 - 1)[None]                # hold altitude
    multi = make_env(N, A, SimParameters(1), scenario('LOWW'), autoreset=False)
    single = make_env(N * A, 1, SimParameters(1), scenario('LOWW'), autoreset=False)
    multi.reset(spawn=spawn); single.reset(spawn=spawn.reshape(N * A, 1, 5))
    alive = np.ones(N, bool)
    for t in range(T):
        om, rm, dm, im = multi.step(torch.from_numpy(acts[t]).cuda())
        os_, rs, ds, is_ = single.step(torch.from_numpy(acts[t].reshape(N * A, 1, 3)).cuda())
        sel = torch.from_numpy(np.repeat(alive, A)).cuda()
        assert torch.equal(om.reshape(N * A, 1, 10)[sel], os_[sel])
        # env reward = pairwise-tree sum of the aircraft rewards (float64 inside, float32 out)
        np.testing.assert_allclose(rm.cpu().numpy()[alive], rs.double().cpu().numpy().reshape(N, A).sum(1)[alive],
                                   rtol=1e-6, atol=1e-6)
        np.testing.assert_array_equal(dm.cpu().numpy()[alive], ds.cpu().numpy().reshape(N, A).any(1)[alive])
        alive &= ~dm.cpu().numpy()
    assert (~alive).sum() > 0 and alive.sum() > 10


def test_zero_wind_is_bit_identical_to_no_wind():
    from atc_reinforcement_learning_b200 import SimParameters
    N, A, T = 2048, 4, 100
    g = torch.Generator(device='cuda').manual_seed(5)
    acts = (torch.rand(T, N, A, 3, device='cuda', generator=g) * 2 - 1)
    e0 = make_env(N, A, SimParameters(1), scenario('LOWW', True), seed=2)
    e1 = make_env(N, A, SimParameters(1), scenario('LOWW', True), seed=2, wind=np.zeros((5, 7, 2), np.float32))
    a, b = e0.rollout(acts), e1.rollout(acts)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    assert torch.equal(e0.state, e1.state)


def test_separation_rule_unit_cases():
    """3 nm / 1000 ft, strict inequalities, evaluated after the move."""
    from atc_reinforcement_learning_b200 import SimParameters
    cases = [  # dx (nm), dh (ft), violation expected
        (2.9, 500.0, True), (3.5, 500.0, False), (2.9, 1001.0, False), (2.9, 998.0, True), (0.5, 1500.0, False)]
    N = len(cases)
    spawn = np.zeros((N, 2, 5))
    for i, (dx, dh, _) in enumerate(cases):
        spawn[i, 0] = [35.0, 45.0, 9000.0, 0.0, 200.0]
        spawn[i, 1] = [35.0 + dx, 45.0, 9000.0 + dh, 0.0, 200.0]       # same heading and speed: geometry is kept
    env = make_env(N, 2, SimParameters(1), scenario('LOWW'), autoreset=False)
    env.reset(spawn=spawn)
    acts = np.zeros((N, 2, 3), np.float32)
    acts[..., 0] = 0.0                                               # 200 kt
    acts[:, 0, 1] = 9000 / 19000 - 1
    acts[:, 1, 1] = [(9000 + c[1]) / 19000 - 1 for c in cases]
    acts[..., 2] = -1.0                                              # heading 0
    obs, rew, done, info = env.step(torch.from_numpy(acts).cuda())
    code = (info['term_code'] & 0xFF).cpu().numpy()
    np.testing.assert_array_equal(code == 5, [c[2] for c in cases])
    np.testing.assert_array_equal(done.cpu().numpy(), [c[2] for c in cases])
    assert (rew.cpu().numpy()[code == 5] < -190).all()


def test_host_buffer_path_matches_device_path():
    from atc_reinforcement_learning_b200 import SimParameters
    N, A, T = 512, 4, 40
    rng = np.random.RandomState(0)
    acts = rng.uniform(-1, 1, (T, N, A, 3)).astype(np.float32)
    e_dev = make_env(N, A, SimParameters(1), scenario('LOWW', True), seed=3)
    e_host = make_env(N, A, SimParameters(1), scenario('LOWW', True), seed=3)
    for t in range(T):
        od, rd, dd, idv = e_dev.step(torch.from_numpy(acts[t]).cuda())
        oh, rh, dh, ih = e_host.step(acts[t])
        assert isinstance(oh, np.ndarray)
        np.testing.assert_array_equal(od.cpu().numpy(), oh)
        np.testing.assert_array_equal(rd.cpu().numpy(), rh)
        np.testing.assert_array_equal(dd.cpu().numpy(), dh)
        np.testing.assert_array_equal(idv['term_code'].cpu().numpy(), ih['term_code'])
    o2 = e_host.rollout(acts[:8])
    o1 = e_dev.rollout(torch.from_numpy(acts[:8]).cuda())
    np.testing.assert_array_equal(o1[0].cpu().numpy(), o2[0])


def test_atcgym_adaptor_kats():
    """Appendix-A KATs through the AtcGym-compatible adaptor (numpy in / numpy out)."""
    from atc_reinforcement_learning_b200 import AtcGym, SimParameters, make
    k = G.kat()
    env = make('AtcEnv-v0')
    obs = env.reset()
    np.testing.assert_allclose(obs, k['K0_reset_obs'], rtol=1e-7)
    tr = G.load_trace('loww_scripted')                   # env 3 = zeros (K1), env 2 = invalid action (K6)
    s, r, d, info = env.step(np.array([0, 0, 0], np.float32))
    assert isinstance(r, float) and isinstance(d, bool) and s.dtype == np.float32 and s.shape == (10,)
    np.testing.assert_allclose(r, tr['reward'][0, 3], rtol=1e-6)
    np.testing.assert_allclose(s, tr['obs'][0, 3], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(info['original_state'], tr['raw_obs'][0, 3], rtol=1e-6)
    env.reset()
    s, r, d, info = env.step(np.array([1.5, -1.2, 3.0], np.float32))
    np.testing.assert_allclose(r, -2.0493659450497237, rtol=1e-6)
    assert not d and info['original_state'][3] == 93 and info['original_state'][4] == 250
    # discrete space (K8)
    envd = AtcGym(sim_parameters=SimParameters(1, discrete_action_space=True))
    assert envd.action_space.nvec.tolist() == k['K8_discrete']['nvec']
    envd.reset()
    s, r, d, info = envd.step(np.array([10, 100, 180]))
    np.testing.assert_allclose(r, k['K8_discrete']['reward'], rtol=1e-6)
    np.testing.assert_allclose(info['original_state'], k['K8_discrete']['raw_obs'], rtol=1e-6)
    np.testing.assert_allclose(s, k['K8_discrete']['obs'], rtol=1e-6, atol=1e-6)
    assert env.timesteps == 1 and env.actions_taken >= 1


def test_error_behaviour():
    from atc_reinforcement_learning_b200 import SimParameters
    from atc_reinforcement_learning_b200._native import AtcError
    with pytest.raises(ValueError):
        make_env(4, 9)
    with pytest.raises(ValueError):
        make_env(0, 1)
    env = make_env(4, 2, SimParameters(1), scenario('LOWW', True))
    with pytest.raises(ValueError):
        env.step(torch.zeros(4, 3, 3, device='cuda'))
    with pytest.raises(ValueError):
        env.reset(spawn=np.tile(np.array([10.0, 51.0, 99999.0, 90.0, 250.0]), (4, 2, 1)))     # invalid altitude
    with pytest.raises(ValueError):
        env.reset(spawn=np.tile(np.array([10.0, 51.0, 9000.0, 90.0, 50.0]), (4, 2, 1)))       # invalid velocity
    # invalid ACTIONS are not errors: -1 per offending channel, that channel untouched (atc_gym.py:312-315)
    from oracle.oracle import Oracle
    ora = Oracle('LOWW', True, n_env=4, n_ac=2, seed=0)
    ora.reset(); ora.reset(); env.reset()
    st0, _ = env.get_state()
    a = torch.zeros(4, 2, 3, device='cuda'); a[..., 0] = 1.5; a[..., 1] = -1.2
    obs, rew, done, info = env.step(a)
    o_obs, o_raw, o_rew, o_done, o_term = ora.step(a.cpu().numpy(), autoreset=True)
    st1, _ = env.get_state()
    assert torch.equal(st0[..., 2], st1[..., 2]) and torch.equal(st0[..., 4], st1[..., 4])   # h and v untouched
    assert not torch.equal(st0[..., 3], st1[..., 3])                                           # phi is never validated
    np.testing.assert_allclose(rew.cpu().numpy(), o_rew, rtol=1e-5, atol=1e-5)
    # 2 aircraft x 2 invalid channels x -1.0 on top of the time penalty; shaping is bounded by 2.4 per aircraft
    assert (o_rew < -4.1 + 4.8).all() and (o_rew > -4.1 - 1e-9).all()
    assert not done.any() and not o_done.any()


def test_one_cta_per_sm_layout_against_the_oracle_and_the_small_layout(monkeypatch):
    """The rollout kernel has two layouts: 14 CTAs of one mover + observer pair per SM with the fine MVA
    grid in global memory, and — for long launches of big batches — one CTA of up to 14 pairs per SM with the compact grid
    (sector.CompactGrid) in shared memory.  Force the second one for short / small / ragged / windy cases and check it
    against the oracle; then both layouts against each other, bit for bit, at the bench size."""
    from atc_reinforcement_learning_b200 import SimParameters
    monkeypatch.setenv('ATC_B200_BIG_MIN_PAIRS', '1')
    monkeypatch.setenv('ATC_B200_BIG_MIN_STEPS', '1')
    run_pair(4096, 4, 200, seed=51, chunk=100)
    for N, A in ((1, 2), (7, 3), (33, 5), (129, 6), (1000, 8), (2048, 1)):
        run_pair(N, A, 100, seed=60 + A)
    rng = np.random.RandomState(99)
    run_pair(2048, 8, 64, seed=4, wind=rng.uniform(-30, 30, (16, 16, 2)).astype(np.float32))
    run_pair(1024, 4, 120, seed=41, exact_math=True)
    run_pair(2048, 2, 300, seed=21, h_bias=True, chunk=150, sector='LOWW')
    run_pair(64, 1, 60, seed=5, sector='SimpleScenario', random_entrypoints=False)
    N, A, T = 16384, 4, 96
    g = torch.Generator(device='cuda').manual_seed(7)
    acts = (torch.rand(T, N, A, 3, device='cuda', generator=g) * 2 - 1)
    big = make_env(N, A, SimParameters(1), scenario('LOWW', True), seed=5)
    monkeypatch.setenv('ATC_B200_NO_SMEM_GRID', '1')
    small = make_env(N, A, SimParameters(1), scenario('LOWW', True), seed=5)
    ob, os_ = big.rollout(acts), small.rollout(acts)
    assert torch.equal(ob[0], os_[0]) and torch.equal(ob[1], os_[1]) and torch.equal(ob[2], os_[2])
    assert torch.equal(ob[3]['term_code'], os_[3]['term_code'])
    assert torch.equal(ob[3]['original_state'], os_[3]['original_state'])
    assert torch.equal(big.state, small.state) and torch.equal(big.ep_return, small.ep_return)
    assert int(ob[2].sum()) > 500                         # episodes ended (and re-spawned) inside the window
