"""§8f rank 1: on-device VecNormalize / VecCheckNan against the numpy restatement of stable-baselines' algorithm."""
import numpy as np
import pytest
import torch

from oracle.vecnorm_ref import RunningMeanStd as RefRMS, normalize as ref_normalize

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


def test_vec_normalize_wrapper_on_the_env():
    from atc_reinforcement_learning_b200 import BatchedAtcEnv, LOWW, SimParameters
    from atc_reinforcement_learning_b200.vec_normalize import VecNormalize
    N, A, T = 512, 2, 60
    env = BatchedAtcEnv(N, A, SimParameters(1), LOWW(random_entrypoints=True), seed=1)
    raw = BatchedAtcEnv(N, A, SimParameters(1), LOWW(random_entrypoints=True), seed=1)
    venv = VecNormalize(env, gamma=0.99, check_nan=True)
    obs_rms, ret_rms, ret = RefRMS(shape=(10,)), RefRMS(shape=()), np.zeros(N)
    o = venv.reset()
    ro = raw.reset().cpu().numpy().astype(np.float64)
    obs_rms.update(ro.reshape(-1, 10))
    np.testing.assert_allclose(o.cpu().numpy(), ref_normalize(ro, obs_rms), rtol=1e-5, atol=1e-5)
    g = torch.Generator(device='cuda').manual_seed(3)
    for t in range(T):
        a = torch.rand(N, A, 3, device='cuda', generator=g) * 2 - 1
        o, r, d, info = venv.step(a)
        ro, rr, rd, _ = raw.step(a)
        ro, rr, rd = ro.cpu().numpy().astype(np.float64), rr.cpu().numpy().astype(np.float64), rd.cpu().numpy()
        obs_rms.update(ro.reshape(-1, 10))
        ret = ret * 0.99 + rr
        ret_rms.update(ret)
        np.testing.assert_allclose(o.cpu().numpy(), ref_normalize(ro, obs_rms), rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(r.cpu().numpy(), np.clip(rr / np.sqrt(ret_rms.var + 1e-8), -10.0, 10.0), rtol=1e-4, atol=1e-4)
        ret[rd] = 0.0
        assert torch.equal(d, torch.from_numpy(rd).cuda())
    with pytest.raises(ValueError):
        from atc_reinforcement_learning_b200.vec_normalize import RunningMeanStd
        RunningMeanStd(64)
