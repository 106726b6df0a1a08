"""Replay of the golden traces (recorded from the live reference by oracle/make_golden.py) through any
implementation exposing the small adaptor interface below.  Used for the oracle (CPU) and the CUDA env (GPU)."""
import glob
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
F_BELOW, F_OUTSIDE, F_WIN, F_TIMEOUT = 1, 2, 4, 8


def trace_names():
    return sorted(os.path.basename(p)[len('trace_'):-len('.npz')] for p in glob.glob(os.path.join(GOLDEN, 'trace_*.npz')))


def load_trace(name):
    z = np.load(os.path.join(GOLDEN, 'trace_%s.npz' % name))
    d = {k: z[k] for k in z.files if k != 'meta'}
    d['meta'] = json.loads(str(z['meta']))
    return d


def kat():
    with open(os.path.join(GOLDEN, 'kat.json')) as f:
        return json.load(f)


def expected_term(flags):
    """Reference override order (atc_gym.py:145-173): MVA branch, then capture, then timeout."""
    code = np.zeros(flags.shape, np.int32)
    code[(flags & F_BELOW) > 0] = 1
    code[(flags & F_OUTSIDE) > 0] = 2
    code[(flags & F_WIN) > 0] = 3
    code[(flags & F_TIMEOUT) > 0] = 4
    return code


def replay(tr, impl, state_atol=1e-9, reward_rtol=1e-9, obs_rtol=2e-7, obs_atol=1e-9, check_metrics=True,
           reward_atol=1e-12, return_rtol=1e-9, return_atol=1e-9):
    """impl interface (all numpy, n_ac == 1):
         impl.reset(mask[E] or None, spawn[E,1,5]) -> obs[E,1,10]
         impl.set_state(state[E,1,5], timesteps[E])
         impl.step(actions[E,1,3]) -> obs, raw_obs, reward[E], done[E], term[E]   (no autoreset)
         impl.get_state() -> state[E,1,5], timesteps[E]
         impl.metrics() -> dict
    Returns a dict of the worst deviations seen."""
    T, E = tr['action'].shape[:2]
    worst = {'state': 0.0, 'reward': 0.0, 'obs': 0.0, 'raw_obs': 0.0}
    obs0 = impl.reset(None, tr['init_state'].reshape(E, 1, 5))
    impl.set_state(tr['init_state'].reshape(E, 1, 5), tr['init_timesteps'])
    np.testing.assert_allclose(obs0.reshape(E, 10), tr['init_obs'], rtol=obs_rtol, atol=obs_atol)
    wins_seen = np.zeros(E, np.int64)
    for t in range(T):
        obs, raw, rew, done, term = impl.step(tr['action'][t].reshape(E, 1, 3))
        st, ts = impl.get_state()
        msg = 'step %d' % t
        np.testing.assert_array_equal(done.astype(np.uint8), tr['done'][t], err_msg=msg)
        np.testing.assert_array_equal(ts, tr['timesteps'][t], err_msg=msg)
        np.testing.assert_array_equal(term & 0xFF, expected_term(tr['flags'][t]), err_msg=msg)
        np.testing.assert_allclose(st.reshape(E, 5), tr['state'][t], rtol=0, atol=state_atol, err_msg=msg)
        np.testing.assert_allclose(rew, tr['reward'][t], rtol=reward_rtol, atol=reward_atol, err_msg=msg)
        np.testing.assert_allclose(obs.reshape(E, 10), tr['obs'][t], rtol=obs_rtol, atol=obs_atol, err_msg=msg)
        np.testing.assert_allclose(raw.reshape(E, 10), tr['raw_obs'][t], rtol=obs_rtol, atol=obs_atol, err_msg=msg)
        worst['state'] = max(worst['state'], float(np.abs(st.reshape(E, 5) - tr['state'][t]).max()))
        worst['reward'] = max(worst['reward'], float(np.abs(rew - tr['reward'][t]).max()))
        worst['obs'] = max(worst['obs'], float(np.abs(obs.reshape(E, 10) - tr['obs'][t]).max()))
        worst['raw_obs'] = max(worst['raw_obs'], float(np.abs(raw.reshape(E, 10) - tr['raw_obs'][t]).max()))
        if check_metrics:
            m = impl.metrics()
            np.testing.assert_array_equal(m['actions_taken'], tr['actions_taken'][t], err_msg=msg)
            np.testing.assert_allclose(m['ep_return'], tr['total_reward'][t], rtol=return_rtol, atol=return_atol, err_msg=msg)
        was_reset = done.astype(bool) & (np.abs(tr['spawn'][t]).sum(-1) > 0)
        if was_reset.any():
            ro = impl.reset(was_reset.astype(np.uint8), tr['spawn'][t].reshape(E, 1, 5))
            np.testing.assert_allclose(ro.reshape(E, 10)[was_reset], tr['reset_obs'][t][was_reset],
                                       rtol=obs_rtol, atol=obs_atol, err_msg=msg)
            if check_metrics and not (tr['flags'] & F_TIMEOUT).any():
                # reference sliding window (atc_gym.py:359-363): 0.1 * wins among the last 9 finished episodes
                m = impl.metrics()
                ratio = np.array([bin(int(w) & 0x1FF).count('1') for w in m['win_ring']]) * 0.1
                np.testing.assert_allclose(ratio[was_reset], tr['winning_ratio'][t][was_reset], atol=1e-9, err_msg=msg)
    return worst
