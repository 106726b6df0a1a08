"""On-device running normalisation of observations and returns, plus a finiteness check — the part of the reference's
training setup that sits directly around the env (`VecNormalize` + `VecCheckNan` from stable-baselines 2.8.0 in
/root/reference/learning/tune_hyperparameters.py:94-97).  SURVEY.md §8f rank 1.  The statistics live in device memory
and are updated by two small kernels through the C ABI (atc_obs_stats_update / atc_obs_normalize); nothing crosses to
the host per step."""
import ctypes as C

import torch

from . import _native as nat


def _p(t):
    return C.c_void_p(t.data_ptr())


class RunningMeanStd(object):
    """stable-baselines `RunningMeanStd` (mean 0, var 1, count epsilon; batched parallel-variance update) on device."""

    def __init__(self, dim, device='cuda:0', epsilon=1e-4):
        if not 1 <= int(dim) <= 32:
            raise ValueError("dim must be in 1..32")
        self.dim = int(dim)
        self.device = torch.device(device)
        self.rms = torch.zeros(2 * self.dim + 1, dtype=torch.float64, device=self.device)
        self.rms[self.dim:2 * self.dim] = 1.0
        self.rms[2 * self.dim] = float(epsilon)
        self._scratch = torch.zeros(2 * self.dim + 2, dtype=torch.float64, device=self.device)
        self.nonfinite = torch.zeros(1, dtype=torch.int32, device=self.device)

    @property
    def mean(self):
        return self.rms[:self.dim]

    @property
    def var(self):
        return self.rms[self.dim:2 * self.dim]

    @property
    def count(self):
        return self.rms[2 * self.dim]

    def _rows(self, x):
        if not (torch.is_tensor(x) and x.device == self.device and x.dtype == torch.float32):
            raise ValueError("expected a float32 tensor on %s" % self.device)
        if x.shape[-1] != self.dim and not (self.dim == 1 and x.numel() > 0):
            raise ValueError("last dimension must be %d" % self.dim)
        x = x.contiguous()
        return x, x.numel() // self.dim

    def update(self, x):
        x, n = self._rows(x)
        with torch.cuda.device(self.device):
            st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            nat.check(None, nat.lib().atc_obs_stats_update(_p(x), n, self.dim, _p(self.rms), _p(self._scratch),
                                                           _p(self.nonfinite), st))

    def normalize(self, x, epsilon=1e-8, clip=10.0, out=None):
        x, n = self._rows(x)
        out = torch.empty_like(x) if out is None else out
        with torch.cuda.device(self.device):
            st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            nat.check(None, nat.lib().atc_obs_normalize(_p(x), n, self.dim, _p(self.rms), float(epsilon), float(clip),
                                                        _p(out), st))
        return out


class VecNormalize(object):
    """Wraps a BatchedAtcEnv like stable-baselines' VecNormalize(norm_obs, norm_reward, clip_obs, clip_reward, gamma,
    epsilon) + VecCheckNan(raise_exception): step() / rollout() outputs are normalised on the device."""

    def __init__(self, env, training=True, norm_obs=True, norm_reward=True, clip_obs=10.0, clip_reward=10.0, gamma=0.99,
                 epsilon=1e-8, check_nan=False):
        self.env = env
        self.training, self.norm_obs, self.norm_reward = training, norm_obs, norm_reward
        self.clip_obs, self.clip_reward, self.gamma, self.epsilon = clip_obs, clip_reward, gamma, epsilon
        self.check_nan = check_nan
        self.obs_rms = RunningMeanStd(10, env.device)
        self.ret_rms = RunningMeanStd(1, env.device)
        self.ret = torch.zeros(env.num_envs, dtype=torch.float32, device=env.device)
        self.num_envs, self.num_aircraft = env.num_envs, env.num_aircraft
        self.action_space, self.observation_space = env.action_space, env.observation_space

    def _obs(self, obs):
        if self.training and self.norm_obs:
            self.obs_rms.update(obs)
        return self.obs_rms.normalize(obs, self.epsilon, self.clip_obs) if self.norm_obs else obs

    def reset(self, *a, **kw):
        self.ret.zero_()
        return self._obs(self.env.reset(*a, **kw))

    def step(self, actions):
        obs, reward, done, info = self.env.step(actions)
        if not torch.is_tensor(obs):
            raise TypeError("VecNormalize works on the device path (pass cuda tensors)")
        obs = self._obs(obs)
        if self.norm_reward:
            self.ret = self.ret * self.gamma + reward
            if self.training:
                self.ret_rms.update(self.ret)
            # stable-baselines scales the reward by the std of the discounted return; it does not centre it
            scale = torch.sqrt(self.ret_rms.var + self.epsilon).to(torch.float32)
            reward = torch.clamp(reward / scale, -self.clip_reward, self.clip_reward)
        self.ret = torch.where(done, torch.zeros_like(self.ret), self.ret)
        if self.check_nan and (int(self.obs_rms.nonfinite.item()) or int(self.ret_rms.nonfinite.item())):
            raise ValueError("NaN or Inf in the observation / reward (VecCheckNan)")
        return obs, reward, done, info

    def get_attr(self, name):
        return self.env.get_attr(name)

    def close(self):
        self.env.close()
