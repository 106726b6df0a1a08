"""§8f rank 1: on-device VecNormalize / VecCheckNan against the numpy restatement of stable-baselines' algorithm."""
import numpy as np
import pytest
import torch

from oracle.vecnorm_ref import RunningMeanStd as RefRMS, VecNormalizeRef, normalize as ref_normalize

pytestmark = pytest.mark.gpu


def test_running_mean_std_matches_reference_algorithm():
    from atc_reinforcement_learning_b200.vec_normalize import RunningMeanStd
    rng = np.random.RandomState(0)
    rms, ref = RunningMeanStd(10), RefRMS(shape=(10,))
    scale = np.array([1, 10, 1000, 38000, 0.01, 5, 300, 1, 180, 2e4])
    for it, n in enumerate((1, 7, 4096, 65536, 333, 1 << 20)):
        x = (rng.randn(n, 10) * scale + scale * (it - 2)).astype(np.float32)
        rms.update(torch.from_numpy(x).cuda())
        ref.update(x.astype(np.float64))
        np.testing.assert_allclose(rms.mean.cpu().numpy(), ref.mean, rtol=1e-10, atol=1e-9)
        np.testing.assert_allclose(rms.var.cpu().numpy(), ref.var, rtol=1e-8)
        np.testing.assert_allclose(float(rms.count), ref.count, rtol=1e-12)
        got = rms.normalize(torch.from_numpy(x).cuda(), 1e-8, 5.0).cpu().numpy()
        np.testing.assert_allclose(got, ref_normalize(x.astype(np.float64), ref, 1e-8, 5.0), rtol=1e-5, atol=1e-6)
    assert int(rms.nonfinite.item()) == 0
    x[5, 3] = np.nan
    rms.update(torch.from_numpy(x).cuda())
    assert int(rms.nonfinite.item()) == 1


def _make_pair(N, A, seed, **kw):
    from atc_reinforcement_learning_b200 import BatchedAtcEnv, LOWW, SimParameters
    from atc_reinforcement_learning_b200.vec_normalize import VecNormalize
    env = BatchedAtcEnv(N, A, SimParameters(1), LOWW(random_entrypoints=True), seed=seed)
    raw = BatchedAtcEnv(N, A, SimParameters(1), LOWW(random_entrypoints=True), seed=seed)
    return VecNormalize(env, check_nan=True, **kw), raw


def test_vec_normalize_wrapper_on_the_env():
    """step(): env launch + ONE fused launch == stable-baselines' step_wait order (numpy restatement, UNPINNED:
    stable-baselines 2.8.0 is not under /root/reference)."""
    N, A, T = 512, 2, 150
    venv, raw = _make_pair(N, A, 1, gamma=0.99)
    ref = VecNormalizeRef(N, gamma=0.99)
    o = venv.reset()
    ro = raw.reset().cpu().numpy()
    np.testing.assert_allclose(o.cpu().numpy(), ref.reset(ro), rtol=1e-5, atol=1e-5)
    g = torch.Generator(device='cuda').manual_seed(3)
    n_done = 0
    for t in range(T):
        a = torch.rand(N, A, 3, device='cuda', generator=g) * 2 - 1
        o, r, d, info = venv.step(a)
        ro, rr, rd, rinfo = raw.step(a)
        eo, er = ref.step(ro.cpu().numpy(), rr.cpu().numpy().astype(np.float64), rd.cpu().numpy())
        np.testing.assert_allclose(o.cpu().numpy(), eo, rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(r.cpu().numpy(), er, rtol=1e-5, atol=1e-5)
        assert torch.equal(d, rd)
        assert torch.equal(venv.get_original_obs(info), rinfo['original_state'])
        n_done += int(rd.sum())
    np.testing.assert_allclose(venv.ret.cpu().numpy(), ref.ret, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(venv.obs_rms.mean.cpu().numpy(), ref.obs_rms.mean, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(venv.obs_rms.var.cpu().numpy(), ref.obs_rms.var, rtol=1e-8)
    np.testing.assert_allclose(float(venv.ret_rms.var[0]), ref.ret_rms.var, rtol=1e-8)
    np.testing.assert_allclose(float(venv.obs_rms.count), ref.obs_rms.count, rtol=1e-12)
    assert n_done > 0
    with pytest.raises(ValueError):
        from atc_reinforcement_learning_b200.vec_normalize import RunningMeanStd
        RunningMeanStd(64)


@pytest.mark.parametrize('N,A,T', [(16384, 4, 48), (300, 3, 130), (1, 1, 40)])
def test_vec_normalize_rollout_is_t_wrapped_steps(N, A, T):
    """rollout(): the T normalisation steps run inside one cooperative launch (a grid barrier per step) and give what T
    wrapped step() calls give — at the bench batch (148 CTAs), a ragged batch and a single env (1 CTA)."""
    venv, raw = _make_pair(N, A, 5, gamma=0.97, clip_obs=4.0, clip_reward=2.5)
    ref = VecNormalizeRef(N, gamma=0.97, clip_obs=4.0, clip_reward=2.5)
    venv.reset(); ref.reset(raw.reset().cpu().numpy())
    g = torch.Generator(device='cuda').manual_seed(9)
    for chunk in range(2):                                # two launches: the scratch rotation carries over
        acts = (torch.rand(T // 8 + 1, N, A, 3, device='cuda', generator=g) * 2 - 1).repeat_interleave(8, 0)[:T].contiguous()
        o, r, d, _ = venv.rollout(acts)
        ro, rr, rd, _ = raw.rollout(acts)
        ro, rr, rd = ro.cpu().numpy(), rr.cpu().numpy().astype(np.float64), rd.cpu().numpy()
        o, r = o.cpu().numpy(), r.cpu().numpy()
        for t in range(T):
            eo, er = ref.step(ro[t], rr[t], rd[t])
            np.testing.assert_allclose(o[t], eo, rtol=1e-5, atol=1e-5, err_msg='obs, step %d' % t)
            np.testing.assert_allclose(r[t], er, rtol=1e-5, atol=1e-5, err_msg='reward, step %d' % t)
    np.testing.assert_allclose(venv.ret.cpu().numpy(), ref.ret, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(venv.obs_rms.var.cpu().numpy(), ref.obs_rms.var, rtol=1e-8)
    assert int(venv._sync[2].item()) == 0                # no barrier time-out


def test_vec_normalize_eval_mode_and_nan_check():
    """training=False freezes the moments (no barrier, one pass); VecCheckNan raises on a NaN observation."""
    N, A = 256, 2
    venv, raw = _make_pair(N, A, 2)
    venv.reset(); raw.reset()
    a = torch.rand(N, A, 3, device='cuda') * 2 - 1
    for _ in range(5):
        venv.step(a); raw.step(a)
    venv.training = False
    mean, var = venv.obs_rms.mean.clone(), venv.obs_rms.var.clone()
    o, r, d, _ = venv.step(a)
    ro, rr, rd, _ = raw.step(a)
    assert torch.equal(mean, venv.obs_rms.mean) and torch.equal(var, venv.obs_rms.var)
    exp = torch.clamp((ro.double() - mean) / torch.sqrt(var + 1e-8), -10, 10).float()
    np.testing.assert_allclose(o.cpu().numpy(), exp.cpu().numpy(), rtol=1e-5, atol=1e-6)
    # VecCheckNan: a NaN in what the env hands out raises (injected directly: the env itself re-spawns a NaN aircraft)
    bad = ro.clone()
    bad[3, 1, 2] = float('nan')
    with pytest.raises(ValueError):
        venv._run(1, bad, rr.clone(), rd.view(torch.uint8).clone())
