"""numpy restatement of stable-baselines 2.8.0 `RunningMeanStd` (common/running_mean_std.py) and the normalisation of
`VecNormalize` (common/vec_env/vec_normalize.py) — TEST INFRASTRUCTURE.  stable-baselines is not vendored under
/root/reference (requirements.txt:117), so this follows its published algorithm: parity UNPINNED."""
import numpy as np


class RunningMeanStd(object):
    def __init__(self, epsilon=1e-4, shape=()):
        self.mean = np.zeros(shape, 'float64')
        self.var = np.ones(shape, 'float64')
        self.count = epsilon

    def update(self, arr):
        batch_mean = np.mean(arr, axis=0)
        batch_var = np.var(arr, axis=0)
        batch_count = arr.shape[0]
        delta = batch_mean - self.mean
        tot_count = self.count + batch_count
        new_mean = self.mean + delta * batch_count / tot_count
        m_a = self.var * self.count
        m_b = batch_var * batch_count
        m_2 = m_a + m_b + np.square(delta) * self.count * batch_count / (self.count + batch_count)
        self.mean, self.var, self.count = new_mean, m_2 / (self.count + batch_count), tot_count


def normalize(x, rms, epsilon=1e-8, clip=10.0):
    return np.clip((x - rms.mean) / np.sqrt(rms.var + epsilon), -clip, clip)


class VecNormalizeRef(object):
    """`VecNormalize.step_wait` / `reset` / `_normalize_observation` of stable-baselines 2.8.0
    (common/vec_env/vec_normalize.py), on numpy arrays the caller obtained from the wrapped env:
        ret = ret * gamma + rews;  obs = _normalize_observation(obs)  [obs_rms.update(obs) when training, then clip];
        ret_rms.update(ret) when training;  rews = clip(rews / sqrt(ret_rms.var + epsilon));  ret[news] = 0.
    Observations are (rows, dim): the batched env hands in one row per aircraft.  UNPINNED like the rest of this file."""

    def __init__(self, n_env, dim=10, training=True, norm_obs=True, norm_reward=True, clip_obs=10.0, clip_reward=10.0,
                 gamma=0.99, epsilon=1e-8):
        self.obs_rms, self.ret_rms = RunningMeanStd(shape=(dim,)), RunningMeanStd(shape=())
        self.ret = np.zeros(n_env)
        self.training, self.norm_obs, self.norm_reward = training, norm_obs, norm_reward
        self.clip_obs, self.clip_reward, self.gamma, self.epsilon = clip_obs, clip_reward, gamma, epsilon
        self.dim = dim

    def _normalize_observation(self, obs):
        if not self.norm_obs:
            return obs
        rows = obs.reshape(-1, self.dim).astype(np.float64)
        if self.training:
            self.obs_rms.update(rows)
        return normalize(rows, self.obs_rms, self.epsilon, self.clip_obs).reshape(obs.shape)

    def reset(self, obs):
        self.ret = np.zeros_like(self.ret)
        return self._normalize_observation(obs)

    def step(self, obs, rews, news):
        self.ret = self.ret * self.gamma + rews
        obs = self._normalize_observation(obs)
        if self.norm_reward:
            if self.training:
                self.ret_rms.update(self.ret)
            rews = np.clip(rews / np.sqrt(self.ret_rms.var + self.epsilon), -self.clip_reward, self.clip_reward)
        self.ret[news] = 0.0
        return obs, rews
