"""On-device running normalisation of observations and returns, plus a finiteness check — the part of the reference's
training setup that sits directly around the env (`VecNormalize` + `VecCheckNan` from stable-baselines 2.8.0 in
/root/reference/learning/tune_hyperparameters.py:94-97).  SURVEY.md §8f rank 1.  The statistics live in device memory
(`RunningMeanStd`: atc_obs_stats_update / atc_obs_normalize for stand-alone use; `VecNormalize`: the fused
atc_vecnorm_run, one launch per step() / rollout()); nothing crosses to the host per step."""
import contextlib
import ctypes as C

import torch

from . import _native as nat


_NULL = contextlib.nullcontext()


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
    epsilon) + VecCheckNan(raise_exception) (learning/tune_hyperparameters.py:94-97).  step() and rollout() are each the
    env's launch plus ONE fused launch (csrc/atc_vecnorm.cu, atc_vecnorm_run): per env step, in stable-baselines' order,
    ret = ret * gamma + reward; obs_rms / ret_rms updated with the step's batch (training); obs and reward normalised
    with the UPDATED moments and clipped; ret[done] = 0.  A T-step rollout applies the T steps one after the other
    inside the launch (a grid barrier per step), i.e. exactly what T wrapped step() calls give.  The running moments,
    the discounted-return accumulators and the NaN flag live in device memory; nothing crosses to the host per step.
    obs / reward are normalised in place (the env's output tensors are returned)."""

    def __init__(self, env, training=True, norm_obs=True, norm_reward=True, clip_obs=10.0, clip_reward=10.0, gamma=0.99,
                 epsilon=1e-8, check_nan=False):
        self.env = env
        self.training, self.norm_obs, self.norm_reward = training, norm_obs, norm_reward
        self.clip_obs, self.clip_reward, self.gamma, self.epsilon = clip_obs, clip_reward, gamma, epsilon
        self.check_nan = check_nan
        dev = env.device
        self.obs_rms = RunningMeanStd(10, dev)
        self.ret_rms = RunningMeanStd(1, dev)
        self.ret = torch.zeros(env.num_envs, dtype=torch.float64, device=dev)    # np.zeros(num_envs) in stable-baselines
        self._scratch = None                                 # work space of the fused launch, grown on demand
        self._sync = torch.zeros(4, dtype=torch.int32, device=dev)
        self.nonfinite = torch.zeros(1, dtype=torch.int32, device=dev)
        self._max_steps = int(nat.lib().atc_vecnorm_max_steps(dev.index))
        self._need = {}                                      # scratch size per launch length
        self.num_envs, self.num_aircraft = env.num_envs, env.num_aircraft
        self.action_space, self.observation_space = env.action_space, env.observation_space

    def _run(self, n_steps, obs, reward=None, done=None):
        if not (torch.is_tensor(obs) and obs.device == self.env.device and obs.dtype == torch.float32 and obs.is_contiguous()):
            raise TypeError("VecNormalize works on the device path (contiguous float32 cuda tensors)")
        N, A = self.num_envs, self.num_aircraft
        if obs.numel() != n_steps * N * A * 10:
            raise ValueError("obs must hold %d x %d x %d x 10 values" % (n_steps, N, A))
        if reward is not None:
            if not (reward.dtype == torch.float32 and reward.is_contiguous() and reward.numel() == n_steps * N and
                    done.dtype in (torch.uint8, torch.bool) and done.is_contiguous() and done.numel() == n_steps * N):
                raise ValueError("reward / done must be contiguous [T, N] float32 / uint8 tensors")
        p = nat.AtcVecNormParams(training=int(bool(self.training)), norm_obs=int(bool(self.norm_obs)),
                                 norm_reward=int(bool(self.norm_reward)), reserved=0, clip_obs=float(self.clip_obs),
                                 clip_reward=float(self.clip_reward), gamma=float(self.gamma), epsilon=float(self.epsilon))
        dev = self.env.device
        L = nat.lib()
        with (_NULL if torch.cuda.current_device() == dev.index else torch.cuda.device(dev)):
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            for t0 in range(0, n_steps, self._max_steps):      # (one launch unless the rollout is thousands of steps long)
                tc = min(self._max_steps, n_steps - t0)
                need = self._need.get(tc)
                if need is None:
                    need = self._need[tc] = int(L.atc_vecnorm_scratch_doubles(dev.index, tc, N, A))
                if need < 0:
                    raise nat.AtcError(L.atc_vecnorm_last_error().decode())
                if self._scratch is None or self._scratch.numel() < need:
                    self._scratch = torch.empty(need, dtype=torch.float64, device=dev)
                state = nat.AtcVecNormState(obs_rms=self.obs_rms.rms.data_ptr(), ret_rms=self.ret_rms.rms.data_ptr(),
                                            ret=self.ret.data_ptr(), scratch=self._scratch.data_ptr(),
                                            scratch_doubles=self._scratch.numel(), sync=self._sync.data_ptr(),
                                            nonfinite=self.nonfinite.data_ptr())
                o = obs.reshape(-1)[t0 * N * A * 10:]
                rp = dp = C.c_void_p(None)
                if reward is not None:
                    rp, dp = _p(reward.reshape(-1)[t0 * N:]), _p(done.reshape(-1)[t0 * N:])
                rc = L.atc_vecnorm_run(C.byref(state), C.byref(p), tc, N, A, _p(o), _p(o), rp, rp, dp, dev.index, stream)
                if rc != 0:
                    raise nat.AtcError(L.atc_vecnorm_last_error().decode())
        if self.check_nan:                                   # VecCheckNan(raise_exception=True): one 4-byte read back
            flags = torch.stack([self.nonfinite[0], self._sync[2]]).tolist()
            if flags[1]:
                raise RuntimeError("atc_vecnorm_run: grid barrier timed out")
            if flags[0]:
                raise ValueError("NaN or Inf in the observation / reward (VecCheckNan)")

    def reset(self, *a, **kw):
        """VecNormalize.reset: zero the return accumulators, update obs_rms with the reset observation, normalise it."""
        self.ret.zero_()
        obs = self.env.reset(*a, **kw)
        self._run(1, obs)
        return obs

    def step(self, actions, out=None):
        obs, reward, done, info = self.env.step(actions, out=out)
        if not torch.is_tensor(obs):
            raise TypeError("VecNormalize works on the device path (pass cuda tensors)")
        d8 = done.view(torch.uint8) if done.dtype == torch.bool else done
        self._run(1, obs, reward, d8)
        return obs, reward, done, info

    def rollout(self, actions, out=None):
        """T fused env steps (BatchedAtcEnv.rollout) followed by the T normalisation steps in one launch."""
        obs, reward, done, info = self.env.rollout(actions, out=out)
        if not torch.is_tensor(obs):
            raise TypeError("VecNormalize works on the device path (pass cuda tensors)")
        d8 = done.view(torch.uint8) if done.dtype == torch.bool else done
        self._run(int(obs.shape[0]), obs, reward, d8)
        return obs, reward, done, info

    def get_original_obs(self, info):
        """stable-baselines' get_original_obs(): the un-normalised observation is what the env reports as
        info['original_state'] (atc_gym.py:192)."""
        return info.get('original_state')

    def get_attr(self, name, indices=None):
        return self.env.get_attr(name, indices)

    def step_async(self, actions):
        self._pending = self.step(actions)

    def step_wait(self):
        res, self._pending = self._pending, None
        return res

    def close(self):
        self.env.close()
