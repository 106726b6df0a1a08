"""Chained rollout launches (atc_rollout_chained): back-to-back rollouts linked by the device-side ready
queue and submitted as programmatic dependent launches must give bit-identical results to ordinary launches, whatever
sits between them, and the float32 snapshot of the return log the launch leaves behind must equal the log."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _env(N, A, seed, **kw):
    from atc_reinforcement_learning_b200 import BatchedAtcEnv, LOWW, SimParameters
    return BatchedAtcEnv(N, A, SimParameters(1), LOWW(random_entrypoints=True), seed=seed, **kw)


def _acts(T, N, A, g):
    return (torch.rand((T + 9) // 10, N, A, 3, device='cuda', generator=g) * 2 - 1).repeat_interleave(10, 0)[:T].contiguous()


def _same(a, b):
    for x, y in zip(a[:3], b[:3]):
        assert torch.equal(x, y)
    assert torch.equal(a[3]['original_state'], b[3]['original_state']) and torch.equal(a[3]['term_code'], b[3]['term_code'])


@pytest.mark.parametrize('N,A', [(16384, 4), (5000, 3), (640, 8)])
def test_chained_rollouts_are_bit_identical(N, A):
    """Six launches of different lengths, chained against unchained: every output of every launch, the final state and
    all counters equal.  16384 x 4 fills the machine (147 CTAs); 5000 x 3 is ragged; 640 x 8 has fewer CTAs than SMs, so
    the CTAs of a chained launch start on idle SMs and really wait for their queue entries."""
    e1, e2 = _env(N, A, 21), _env(N, A, 21)
    g = torch.Generator(device='cuda').manual_seed(4)
    lens = [64, 200, 33, 120, 64, 500]
    streams = [_acts(T, N, A, g) for T in lens]
    res1 = [e1.rollout(a) for a in streams]
    res2 = []
    for k, a in enumerate(streams):
        res2.append(e2.rollout(a, chain=True))
        ll = e2.last_launch
        assert ll['kernel'] == 3
        assert ll['chained'] == (1 if k > 0 else 0), (k, ll)
    assert e2.chain_status() == len(lens) - 1
    for a, b in zip(res1, res2):
        _same(a, b)
    s1, t1 = e1.get_state(); s2, t2 = e2.get_state()
    assert torch.equal(s1, s2) and torch.equal(t1, t2)
    for name in ('episodes', 'ep_return', 'last_ep_return', 'last_ep_len', 'win_ring'):
        assert torch.equal(getattr(e1, name), getattr(e2, name)), name
    assert torch.equal(e2.ret_log, e2.last_ep_return.float())
    assert int((e2.episodes > 1).sum()) > 0


def test_chain_is_broken_and_restarted_by_anything_in_between():
    """step(), reset(mask), a rollout of the other layout (too short), another stream: each starts a new chain; results
    stay identical to ordinary launches."""
    N, A = 4096, 4
    e1, e2 = _env(N, A, 9), _env(N, A, 9)
    g = torch.Generator(device='cuda').manual_seed(8)
    mask = (torch.arange(N, device='cuda') % 7 == 0).to(torch.uint8)
    side = torch.cuda.Stream()

    def script(env, chain):
        out, flags = [], []
        kw = {'chain': True} if chain else {}
        for step in range(3):
            out.append(env.rollout(_a[0 + 4 * step], **kw)); flags.append(env.last_launch['chained'])
            out.append(env.rollout(_a[1 + 4 * step], **kw)); flags.append(env.last_launch['chained'])
            if step == 0:
                out.append(env.step(_a[2][0]))
            elif step == 1:
                env.reset(mask=mask)
            else:
                out.append(env.rollout(_a[2][:8], **kw)); flags.append(env.last_launch['chained'])   # short: other layout
            out.append(env.rollout(_a[3 + 4 * step], **kw)); flags.append(env.last_launch['chained'])
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            out.append(env.rollout(_a[12], **kw)); flags.append(env.last_launch['chained'])
        torch.cuda.current_stream().wait_stream(side)
        out.append(env.rollout(_a[13], **kw)); flags.append(env.last_launch['chained'])
        return out, flags

    _a = [_acts(T, N, A, g) for T in (64, 96, 40, 64, 80, 64, 40, 50, 64, 70, 40, 64, 48, 64)]
    r1, _ = script(e1, False)
    r2, flags = script(e2, True)
    for a, b in zip(r1, r2):
        _same(a, b)
    #        r   r   r(after step)  r  r  r(after reset)  r  r  short  r(after short)  side  main
    assert flags == [0, 1, 0, 1, 1, 0, 1, 1, 0, 0, 0, 0], flags
    assert e2.chain_status() == 5
    s1, _ = e1.get_state(); s2, _ = e2.get_state()
    assert torch.equal(s1, s2)


def test_chained_rollouts_against_the_oracle_and_the_return_log_buffer():
    from oracle.oracle import Oracle
    N, A, seed = 2048, 4, 13
    env = _env(N, A, seed)
    ora = Oracle('LOWW', True, n_env=N, n_ac=A, seed=seed)
    ora.reset(); ora.reset()
    rng = np.random.RandomState(2)
    logs = [torch.zeros(N, dtype=torch.float32, device='cuda') for _ in range(2)]
    for k in range(4):
        T = 150
        acts = np.repeat(rng.uniform(-1, 1, (T // 15, N, A, 3)).astype(np.float32), 15, 0)
        obs, rew, done, info = env.rollout(torch.from_numpy(acts).cuda(), chain=True, ret_log=logs[k & 1])
        o_obs, o_raw, o_rew, o_done, o_term = ora.rollout(acts, raw=True)
        np.testing.assert_array_equal(done.cpu().numpy().astype(np.uint8), o_done)
        np.testing.assert_array_equal(info['term_code'].cpu().numpy(), o_term)
        np.testing.assert_allclose(obs.cpu().numpy(), o_obs, rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(info['original_state'].cpu().numpy(), o_raw, rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(rew.cpu().numpy(), o_rew, rtol=1e-5, atol=1e-5)
        assert torch.equal(logs[k & 1], env.last_ep_return.float())
    assert env.chain_status() == 3
    with pytest.raises(ValueError):
        env.rollout(torch.zeros(8, N, A, 3, device='cuda'), ret_log=torch.zeros(N, device='cuda', dtype=torch.float64))
