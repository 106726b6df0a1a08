"""Batched ATC approach-control environment: the reference's gym.Env surface (AtcGym, /root/reference/
envs/atc/atc_gym.py:22-192, 337-365) with a leading batch dimension, running on one B200 through the C ABI in
include/atc_b200.h.  PyTorch owns every device tensor; the kernels live in csrc/atc_kernels.cu.

    env = BatchedAtcEnv(num_envs=16384, num_aircraft=4, scenario=LOWW(random_entrypoints=True))
    obs = env.reset()                                  # [N, A, 10] float32, cuda
    obs, reward, done, info = env.step(actions)        # actions [N, A, 3] float32 (cuda tensor or numpy)
    obs, reward, done, info = env.rollout(actions_T)   # actions [T, N, A, 3]: T fused steps, one launch

`AtcGym` is the call-compatible single-env adaptor (N = 1, one aircraft, numpy in / numpy out, no auto-reset).
"""
import contextlib
import copy
import ctypes as C

import numpy as np
import torch

from . import _native as nat
from . import model, scenarios, spaces
from .sector import CompiledSector

TERM_NAMES = ('running', 'below_mva', 'left_airspace', 'captured', 'timeout', 'separation')


_SECTOR_CACHE = {}


_COMPACT_CACHE = {}
_NULL_CTX = contextlib.nullcontext()


def compile_sector_cached(scenario, cell, wind=None):
    """CompiledSector is a pure function of the scenario data; building the fine MVA grid takes a couple of seconds,
    so identical (scenario, cell) pairs share one compiled sector (wind grids are not cached)."""
    if wind is not None:
        return CompiledSector(scenario, cell=cell, wind=wind)
    key = (float(cell),
           tuple((m.height, tuple(m.area_as_list)) for m in scenario.mvas),
           (scenario.runway.x, scenario.runway.y, scenario.runway.h, scenario.runway.phi_from_runway),
           tuple((e.x, e.y, e.phi, tuple(e.levels)) for e in scenario.entrypoints))
    cs = _SECTOR_CACHE.get(key)
    if cs is None:
        cs = _SECTOR_CACHE[key] = CompiledSector(scenario, cell=cell)
    return cs


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(None)


class BatchedAtcEnv(object):
    """N independent envs x A aircraft, one fused CUDA kernel per step (or per T-step rollout).

    Differences from the reference env, all forced by batching:
      * auto-reset: a finished env is re-spawned inside the step and `obs` holds its reset observation
        (VecEnv convention); the terminal raw observation is in info['original_state'].  Pass autoreset=False to
        get the reference's behaviour (keeps simulating until reset(mask) is called).
      * the reset observation is RAW (un-normalised) like the reference's reset() (atc_gym.py:351,365) unless
        normalize_reset_obs=True.
      * spawns come from a counter-based device RNG (or are injected with reset(spawn=...)), not CPython's `random`.
      * num_aircraft > 1, 3 nm / 1000 ft separation and wind are extensions; with 1 aircraft and no wind every
        output is the reference's.
    """
    metadata = {'render.modes': []}
    reward_range = (-3000.0, 23000.0)          # atc_gym.py:115 (kept verbatim; a first-step win exceeds it)

    def __init__(self, num_envs, num_aircraft=1, sim_parameters=None, scenario=None, device='cuda:0', seed=0,
                 wind=None, autoreset=True, track_actions=False, return_raw_obs=True, normalize_reset_obs=False,
                 env_index_base=0, grid_cell=0.0625, exact_math=False, compact_grid=True):
        self._handle = None
        if sim_parameters is None:
            sim_parameters = model.SimParameters(1)
        if scenario is None:
            scenario = scenarios.LOWW()
        if not 1 <= int(num_aircraft) <= nat.MAX_AIRCRAFT:
            raise ValueError("num_aircraft must be in 1..%d" % nat.MAX_AIRCRAFT)
        if int(num_envs) < 1:
            raise ValueError("num_envs must be >= 1")
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise ValueError("BatchedAtcEnv runs on a CUDA device only (there is no CPU path)")
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device available (there is no CPU path)")
        if self.device.index is None:
            self.device = torch.device('cuda', torch.cuda.current_device())
        self._dev_index = self.device.index
        self.num_envs, self.num_aircraft = int(num_envs), int(num_aircraft)
        self._sim_parameters = sim_parameters
        self._scenario = scenario
        self.sector = compile_sector_cached(scenario, grid_cell, wind)
        if compact_grid and not hasattr(self.sector, 'compact'):
            # coarse copy of the MVA grid for the shared memory of one SM (sector.CompactGrid); None if it cannot be built.
            # A pure function of the polygons and the budget: shared between envs (building it takes seconds).
            from .sector import build_compact_grid
            budget = int(nat.lib().atc_compact_grid_budget())
            key = (budget, tuple((m.height, tuple(m.area_as_list)) for m in scenario.mvas))
            if key not in _COMPACT_CACHE:
                _COMPACT_CACHE[key] = build_compact_grid(scenario, budget)
            self.sector.compact = _COMPACT_CACHE[key]
        elif not compact_grid:
            self.sector = copy.copy(self.sector)
            self.sector.compact = None
        self.autoreset = bool(autoreset)
        self.track_actions = bool(track_actions)
        self.return_raw_obs = bool(return_raw_obs)
        self.timestep_limit = 6000
        self._seed = int(seed)
        self._env_index_base = int(env_index_base)
        self._normalize_reset_obs = bool(normalize_reset_obs)
        self.exact_math = bool(exact_math)

        if sim_parameters.discrete_action_space:            # atc_gym.py:66-82
            self.action_space = spaces.MultiDiscrete([int((model.V_MAX - model.V_MIN) / 10), int(model.H_MAX / 100), 360])
        else:
            self.action_space = spaces.Box(low=np.array([-1, -1, -1]), high=np.array([1, 1, 1]))
        self.observation_space = spaces.Box(low=-1.0, high=1.0, shape=(10,))       # atc_gym.py:113
        self.normalization_state_min = self.sector.norm_min.copy()
        self.normalization_state_max = self.sector.norm_max.copy()

        N, A, dev = self.num_envs, self.num_aircraft, self.device
        self.state = torch.zeros(5, N * A, dtype=torch.float64, device=dev)
        self.timesteps = torch.zeros(N, dtype=torch.int32, device=dev)
        self.episodes = torch.zeros(N, dtype=torch.int32, device=dev)
        self.ep_return = torch.zeros(N, dtype=torch.float64, device=dev)
        self.last_ep_return = torch.zeros(N, dtype=torch.float64, device=dev)
        self.last_ep_len = torch.zeros(N, dtype=torch.int32, device=dev)
        self.win_ring = torch.zeros(N, dtype=torch.int32, device=dev)
        self.last_action = torch.zeros(3, N * A, dtype=torch.float64, device=dev) if track_actions else None
        self.actions_taken = torch.zeros(N, dtype=torch.int32, device=dev) if track_actions else None
        self._host = None          # pinned staging for the host-buffer (end-to-end) path, allocated on first use
        self._create()
        self.reset()

    # ------------------------------------------------------------------------------------------ native plumbing
    def _create(self):
        p = nat.AtcSimParams()
        sp = self._sim_parameters
        p.timestep = float(sp.timestep)
        p.reward_shaping = int(bool(sp.reward_shaping))
        p.normalize_state = int(bool(sp.normalize_state))
        p.discrete_action_space = int(bool(sp.discrete_action_space))
        p.normalize_reset_obs = int(self._normalize_reset_obs)
        p.n_env, p.n_aircraft = self.num_envs, self.num_aircraft
        p.track_actions = int(self.track_actions)
        p.exact_math = int(self.exact_math)
        p.seed = self._seed & 0xFFFFFFFFFFFFFFFF
        p.env_index_base = self._env_index_base
        desc = nat.sector_desc(self.sector)
        h = C.c_void_p()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        nat.check(None, nat.lib().atc_create(C.byref(desc), C.byref(p), dev_index, C.byref(h)))
        self._handle = h
        self._buffers = nat.AtcBuffers(
            state=self.state.data_ptr(), last_action=self.last_action.data_ptr() if self.track_actions else None,
            timesteps=self.timesteps.data_ptr(), episodes=self.episodes.data_ptr(), ep_return=self.ep_return.data_ptr(),
            actions_taken=self.actions_taken.data_ptr() if self.track_actions else None,
            last_ep_return=self.last_ep_return.data_ptr(), last_ep_len=self.last_ep_len.data_ptr(),
            win_ring=self.win_ring.data_ptr())

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _on_device(self):
        """Context that makes the env's GPU current; free when it already is (the common case: the context manager
        alone costs several microseconds of the ~20 us a single-step call takes on the host)."""
        if torch.cuda.current_device() == self._dev_index:
            return _NULL_CTX
        return torch.cuda.device(self.device)

    def close(self):
        if getattr(self, '_handle', None):
            nat.lib().atc_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def seed(self, seed=None):
        """atc_gym.py:117-126.  Re-keys the spawn RNG (takes effect from the next reset)."""
        self._seed = 0 if seed is None else int(seed)
        self.close()
        self._create()
        return [seed]

    @property
    def launch_count(self):
        return int(nat.lib().atc_launch_count(self._handle))

    @property
    def last_launch(self):
        """What the last step() / rollout() actually launched (atc_last_launch_info): kernel name with its template
        arguments, grid, block, pairs per CTA, steps, dynamic shared memory — for truthful benchmark labels."""
        info = nat.AtcLaunchInfo()
        nat.check(self._handle, nat.lib().atc_last_launch_info(self._handle, C.byref(info)))
        d = {k: int(getattr(info, k)) for k, _ in nat.AtcLaunchInfo._fields_}
        b = lambda v: 'true' if v else 'false'
        name = nat.KERNEL_NAMES.get(d['kernel'], '?')
        if d['kernel'] == 1:
            name += '<%d,%s,%s,%s>' % (d['lanes_per_env'], b(d['wind']), b(d['track_actions']), b(d['exact_math']))
        elif d['kernel'] in (2, 3):
            name += '<%d,%s,%s,%d,%d>' % (d['lanes_per_env'], b(d['wind']), b(d['track_actions']), d['cfg'],
                                          14 if d['kernel'] == 3 else 1)
        d['name'] = name
        d['layout'] = {1: 'fused step kernel, 64-thread CTAs',
                       2: 'one mover + observer warp pair per 64-thread CTA, MVA grid through L1 / L2',
                       3: 'one CTA per SM, %d mover + observer warp pairs, compact MVA grid in shared memory'
                          % d['pairs_per_cta']}.get(d['kernel'], 'none')
        return d

    # ------------------------------------------------------------------------------------------ validation helpers
    def _as_actions(self, actions, lead):
        shape = tuple(lead) + (self.num_envs, self.num_aircraft, 3)
        if isinstance(actions, np.ndarray):
            return None, np.ascontiguousarray(actions, dtype=np.float32).reshape(shape)
        if not torch.is_tensor(actions):
            raise TypeError("actions must be a torch tensor or a numpy array")
        if actions.device != self.device:
            raise ValueError("actions live on %s, env on %s" % (actions.device, self.device))
        a = actions
        if a.dtype != torch.float32:
            a = a.to(torch.float32)
        n = self.num_envs * self.num_aircraft * 3
        for d in lead:
            n *= d
        if a.numel() != n:
            raise ValueError("actions must have shape %s, got %s" % (shape, tuple(actions.shape)))
        if a.shape != shape or not a.is_contiguous():
            a = a.reshape(shape).contiguous()
        return a, None

    def _alloc_io(self, lead, out=None):
        N, A, dev = self.num_envs, self.num_aircraft, self.device
        lead = tuple(lead)
        if out is not None:
            self._check_io(out, lead, host=False)
            return out
        io = {'obs': torch.empty(lead + (N, A, 10), dtype=torch.float32, device=dev),
              'reward': torch.empty(lead + (N,), dtype=torch.float32, device=dev),
              'done': torch.empty(lead + (N,), dtype=torch.uint8, device=dev),
              'term': torch.empty(lead + (N,), dtype=torch.int32, device=dev)}
        if self.return_raw_obs:
            io['raw_obs'] = torch.empty(lead + (N, A, 10), dtype=torch.float32, device=dev)
        return io

    _IO_SPEC = {'obs': (torch.float32, 'A10'), 'raw_obs': (torch.float32, 'A10'), 'reward': (torch.float32, ''),
                'done': (torch.uint8, ''), 'term': (torch.int32, '')}

    def _check_io(self, io, lead, host):
        """The C ABI receives raw pointers: every caller-supplied output tensor must have exactly the element count,
        dtype, placement and contiguity the kernels / copies assume, or the call would write out of bounds."""
        N, A = self.num_envs, self.num_aircraft
        n_lead = 1
        for d in lead:
            n_lead *= int(d)
        for k in ('obs', 'reward', 'done'):
            if io.get(k) is None:
                raise ValueError("out[%r] is required" % k)
        for k, t in io.items():
            if t is None:
                continue
            if k not in self._IO_SPEC:
                raise ValueError("unknown output %r" % k)
            dt, kind = self._IO_SPEC[k]
            want = n_lead * N * (A * 10 if kind else 1)
            if not torch.is_tensor(t) or t.dtype != dt or not t.is_contiguous() or t.numel() != want:
                raise ValueError("out[%r] must be a contiguous %s tensor with %d elements (lead %s)"
                                 % (k, dt, want, tuple(lead)))
            if host:
                if t.device.type != 'cpu' or not t.is_pinned():
                    raise ValueError("out[%r] must be a pinned host tensor" % k)
            elif t.device != self.device:
                raise ValueError("out[%r] lives on %s, env on %s" % (k, t.device, self.device))

    def _step_io(self, actions, io):
        return nat.AtcStepIO(actions=actions.data_ptr(), obs=io['obs'].data_ptr(),
                             raw_obs=io['raw_obs'].data_ptr() if io.get('raw_obs') is not None else None,
                             reward=io['reward'].data_ptr(), done=io['done'].data_ptr(),
                             term=io['term'].data_ptr() if io.get('term') is not None else None)

    def _info(self, io):
        info = {'term_code': io.get('term')}
        if io.get('raw_obs') is not None:
            info['original_state'] = io['raw_obs']          # atc_gym.py:192
        return info

    # ------------------------------------------------------------------------------------------ gym surface
    def reset(self, mask=None, spawn=None):
        """AtcGym.reset (atc_gym.py:337-365) for all envs, or those with mask != 0.  `spawn` [N, A, 5]
        (x, y, h, phi, v) overrides the RNG.  Returns obs [N, A, 10]; rows of envs not reset are zero."""
        N, A = self.num_envs, self.num_aircraft
        with torch.cuda.device(self.device):
            obs = torch.zeros(N, A, 10, dtype=torch.float32, device=self.device)
            m = None
            if mask is not None:
                m = torch.as_tensor(mask).to(device=self.device, dtype=torch.uint8).contiguous()
                if m.numel() != N:
                    raise ValueError("mask must have %d elements" % N)
            s = None
            if spawn is not None:
                s = torch.as_tensor(spawn).to(device=self.device, dtype=torch.float64).reshape(N, A, 5).contiguous()
                sel = s if m is None else s[m.bool()]
                if ((sel[..., 2] < model.H_MIN) | (sel[..., 2] > model.H_MAX)).any():
                    raise ValueError("invalid altitude")        # Airplane.__init__, model.py:35-36
                if ((sel[..., 4] < model.V_MIN) | (sel[..., 4] > model.V_MAX)).any():
                    raise ValueError("invalid velocity")        # model.py:38-39
            nat.check(self._handle, nat.lib().atc_reset(self._handle, C.byref(self._buffers), _ptr(m), _ptr(s),
                                                        _ptr(obs), self._stream()))
        return obs

    def step(self, actions, out=None):
        """AtcGym.step (atc_gym.py:128-192), batched.  actions [N, A, 3] (float32; MultiDiscrete indices for the
        discrete space).  A cuda tensor runs device-to-device; a numpy array takes the end-to-end host path
        (pinned H2D, kernel, D2H) and returns numpy arrays."""
        dev_a, host_a = self._as_actions(actions, ())
        with self._on_device():
            if host_a is not None:
                return self._run_host(host_a, 1, self.autoreset, ())
            io = self._alloc_io((), out)
            sio = self._step_io(dev_a, io)
            nat.check(self._handle, nat.lib().atc_step(self._handle, C.byref(self._buffers), C.byref(sio),
                                                       int(self.autoreset), self._stream()))
        return io['obs'], io['reward'], io['done'].bool() if out is None else io['done'], self._info(io)

    def rollout(self, actions, out=None):
        """T fused steps in one kernel launch (state stays in registers, auto-reset on).  actions [T, N, A, 3].
        Returns obs [T, N, A, 10], reward [T, N], done [T, N], info."""
        if actions.ndim != 4:
            raise ValueError("rollout actions must have shape [T, N, A, 3]")
        T = int(actions.shape[0])
        if T < 1:
            raise ValueError("rollout needs at least one step")
        dev_a, host_a = self._as_actions(actions, (T,))
        with self._on_device():
            if host_a is not None:
                return self._run_host(host_a, T, True, (T,))
            io = self._alloc_io((T,), out)
            sio = self._step_io(dev_a, io)
            nat.check(self._handle, nat.lib().atc_rollout(self._handle, C.byref(self._buffers), C.byref(sio), T,
                                                          self._stream()))
        return io['obs'], io['reward'], io['done'].bool() if out is None else io['done'], self._info(io)

    # ------------------------------------------------------------------------------------------ end-to-end host path
    def _host_buffers(self, T):
        if self._host is not None and self._host['T'] == T:
            return self._host
        N, A = self.num_envs, self.num_aircraft
        lead = (T,)

        def pin(shape, dtype):
            return torch.empty(shape, dtype=dtype).pin_memory()

        h = {'T': T,
             'h_actions': pin(lead + (N, A, 3), torch.float32),
             'h': {'obs': pin(lead + (N, A, 10), torch.float32), 'reward': pin(lead + (N,), torch.float32),
                   'done': pin(lead + (N,), torch.uint8), 'term': pin(lead + (N,), torch.int32)}}
        st = self._device_staging(T)
        h['d_actions'], h['d'] = st['d_actions'], st['d']
        if self.return_raw_obs:
            h['h']['raw_obs'] = pin(lead + (N, A, 10), torch.float32)
        self._host = h
        return h

    def _device_staging(self, T):
        """Device-side staging of the host-buffer path (actions in, outputs out), cached per T.  Separate from the
        pinned host buffers: rollout_pinned() with caller-owned pinned tensors must not touch those."""
        st = getattr(self, '_staging', None)
        if st is None or st['T'] != T:
            N, A = self.num_envs, self.num_aircraft
            st = self._staging = {'T': T, 'd_actions': torch.empty((T, N, A, 3), dtype=torch.float32, device=self.device),
                                  'd': self._alloc_io((T,))}
        return st

    def _run_host(self, host_actions, T, autoreset, lead):
        hb = self._host_buffers(T)
        hb['h_actions'].numpy()[...] = host_actions.reshape(hb['h_actions'].shape)
        hio = self._step_io(hb['h_actions'], hb['h'])
        dio = self._step_io(hb['d_actions'], hb['d'])
        L = nat.lib()
        if len(lead) == 0:
            rc = L.atc_step_host(self._handle, C.byref(self._buffers), C.byref(hio), C.byref(dio), int(autoreset),
                                 self._stream())
        else:
            rc = L.atc_rollout_host(self._handle, C.byref(self._buffers), C.byref(hio), C.byref(dio), T, self._stream())
        nat.check(self._handle, rc)
        N, A = self.num_envs, self.num_aircraft
        h = hb['h']
        obs = h['obs'].numpy().reshape(lead + (N, A, 10)).copy()
        info = {'term_code': h['term'].numpy().reshape(lead + (N,)).copy()}
        if 'raw_obs' in h:
            info['original_state'] = h['raw_obs'].numpy().reshape(lead + (N, A, 10)).copy()
        return (obs, h['reward'].numpy().reshape(lead + (N,)).copy(),
                h['done'].numpy().reshape(lead + (N,)).astype(bool), info)

    def alloc_pinned_io(self, T):
        """Pinned host buffers for rollout_pinned(): (actions [T, N, A, 3], out dict)."""
        hb = self._host_buffers(int(T))
        return hb['h_actions'], hb['h']

    def rollout_pinned(self, h_actions, h_out):
        """End-to-end rollout straight through the C ABI's host-buffer entry point: actions are read from the pinned
        host tensor `h_actions` [T, N, A, 3], results land in the pinned host tensors of `h_out`
        (obs / reward / done / term [/ raw_obs]); H2D copy, the fused kernel and the D2H copies all run on the
        current stream and the call returns after they completed.  No intermediate host copies."""
        T = int(h_actions.shape[0])
        if not (h_actions.is_pinned() and h_actions.dtype == torch.float32 and h_actions.is_contiguous()):
            raise ValueError("h_actions must be a pinned, contiguous float32 host tensor")
        if h_actions.numel() != T * self.num_envs * self.num_aircraft * 3:
            raise ValueError("h_actions must have shape [T, N, A, 3]")
        self._check_io(h_out, (T,), host=True)
        if h_out.get('raw_obs') is not None and not self.return_raw_obs:
            raise ValueError("h_out['raw_obs'] given but the env was created with return_raw_obs=False")
        hb = self._device_staging(T)
        hio = self._step_io(h_actions, h_out)
        dio = self._step_io(hb['d_actions'], hb['d'])
        with torch.cuda.device(self.device):
            nat.check(self._handle, nat.lib().atc_rollout_host(self._handle, C.byref(self._buffers), C.byref(hio),
                                                               C.byref(dio), T, self._stream()))
        return h_out

    # ------------------------------------------------------------------------------------------ state access / metrics
    def get_state(self):
        """(state [N, A, 5] float64 = x, y, h, phi, v ; timesteps [N]) as new tensors."""
        N, A = self.num_envs, self.num_aircraft
        return self.state.reshape(5, N, A).permute(1, 2, 0).contiguous(), self.timesteps.clone()

    def set_state(self, state, timesteps=None):
        N, A = self.num_envs, self.num_aircraft
        s = torch.as_tensor(state).to(device=self.device, dtype=torch.float64).reshape(N, A, 5)
        self.state.copy_(s.permute(2, 0, 1).reshape(5, N * A))
        if timesteps is not None:
            self.timesteps.copy_(torch.as_tensor(timesteps).to(device=self.device, dtype=torch.int32))

    @property
    def winning_ratio(self):
        """atc_gym.py:359-363: 0.1 x wins among the last 9 finished episodes (the reference's window), per env."""
        w = self.win_ring & 0x1FF
        cnt = torch.zeros_like(w)
        for k in range(9):
            cnt += (w >> k) & 1
        return cnt.to(torch.float64) * 0.1

    @property
    def actions_per_timestep(self):
        """atc_gym.py:197 (needs track_actions=True)."""
        if not self.track_actions:
            raise AttributeError("actions_per_timestep needs track_actions=True")
        return self.actions_taken.to(torch.float64) / self.timesteps.clamp(min=1).to(torch.float64)

    @property
    def total_reward(self):
        return self.ep_return

    # ---- the stable-baselines VecEnv surface the reference's training scripts drive through SubprocVecEnv /
    # DummyVecEnv (learning/atc-gym-stable-baselines.py:76-85, learning/tune_hyperparameters.py:89-97).  SB 2.8.0 is not
    # under /root/reference: names and argument meaning are from its published VecEnv base class (unpinned).
    def _indices(self, indices):
        if indices is None:
            return list(range(self.num_envs))
        if isinstance(indices, int):
            indices = [indices]
        idx = [int(i) for i in indices]
        for i in idx:
            if not -self.num_envs <= i < self.num_envs:
                raise IndexError("env index %d out of range" % i)
        return idx

    def get_attr(self, name, indices=None):
        """VecEnv.get_attr — used by the reference's training callback (learning/atc-gym-stable-baselines.py:34-36).
        Per-env device counters come back as a list with one entry per (selected) env."""
        v = getattr(self, name)
        idx = self._indices(indices)
        if torch.is_tensor(v) and v.ndim >= 1 and v.shape[0] == self.num_envs:
            v = v.tolist()
            return v if indices is None else [v[i] for i in idx]
        return [v] * len(idx)

    def set_attr(self, name, value, indices=None):
        """VecEnv.set_attr.  The batch shares one configuration: an attribute can only be set for all envs at once;
        per-env device counters (timesteps, ep_return, ...) accept a scalar or a sequence for the selected envs."""
        cur = getattr(self, name, None)
        if torch.is_tensor(cur) and cur.ndim >= 1 and cur.shape[0] == self.num_envs:
            idx = torch.as_tensor(self._indices(indices), device=self.device, dtype=torch.long)
            cur[idx] = torch.as_tensor(value, device=self.device).to(cur.dtype)
            return
        if indices is not None and len(self._indices(indices)) != self.num_envs:
            raise ValueError("%r is shared by the whole batch: set it for all envs (indices=None)" % name)
        setattr(self, name, value)

    def env_method(self, method_name, *method_args, indices=None, **method_kwargs):
        """VecEnv.env_method: the batched env IS every sub-env, so the method runs once on the batch and its result is
        repeated per selected env (tensors with a leading env dimension are split per env)."""
        res = getattr(self, method_name)(*method_args, **method_kwargs)
        idx = self._indices(indices)
        if torch.is_tensor(res) and res.ndim >= 1 and res.shape[0] == self.num_envs:
            return [res[i] for i in idx]
        return [res] * len(idx)

    def step_async(self, actions):
        """VecEnv.step_async: enqueue the step on the current CUDA stream and return at once (kernel launches are
        asynchronous by nature); step_wait() hands out the result tensors."""
        if getattr(self, '_pending', None) is not None:
            raise RuntimeError("step_async called twice without step_wait")       # SB's AlreadySteppingError
        self._pending = self.step(actions)

    def step_wait(self):
        """VecEnv.step_wait: the (obs, reward, done, info) of the step enqueued by step_async().  The tensors are
        ordered on the current stream like any other result of step(); nothing blocks the host here."""
        if getattr(self, '_pending', None) is None:
            raise RuntimeError("step_wait called without step_async")              # SB's NotSteppingError
        res, self._pending = self._pending, None
        return res

    def query_mva(self, xy):
        """MVA height [ft] (or -1 outside) for points [n, 2] — the kernel's find_mva (model.py:282-292)."""
        xy = torch.as_tensor(xy).to(device=self.device, dtype=torch.float64).reshape(-1, 2).contiguous()
        out = torch.empty(xy.shape[0], dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            nat.check(self._handle, nat.lib().atc_query_mva(self._handle, xy.shape[0], _ptr(xy), _ptr(out), self._stream()))
        return out

    def query_corridor(self, xyhphi):
        """Runway.inside_corridor (model.py:188-231, 248) for rows [x, y, h, phi]."""
        q = torch.as_tensor(xyhphi).to(device=self.device, dtype=torch.float64).reshape(-1, 4).contiguous()
        out = torch.empty(q.shape[0], dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            nat.check(self._handle, nat.lib().atc_query_corridor(self._handle, q.shape[0], _ptr(q), _ptr(out), self._stream()))
        return out.bool()

    def render(self, mode='rgb_array', env_index=0, trail_xy=None, labels=True, last_reward=None):
        """AtcGym.render (atc_gym.py:367-552) for one env of the batch, headless: mode 'rgb_array' returns a numpy
        uint8 image [H, W, 3] drawn by CUDA kernels (render.py): sector, runway, approach, aircraft, trail and — with
        labels=True — the reference's text labels in a 5 x 7 bitmap font.  There is no window system here: mode 'human'
        raises."""
        if mode != 'rgb_array':
            raise NotImplementedError("only mode='rgb_array' is available (headless; SURVEY.md §2 row 7, §8f rank 4)")
        from .render import render_rgb
        return render_rgb(self, env_index, trail_xy, labels=labels, last_reward=last_reward).cpu().numpy()


class AtcGym(object):
    """Call-compatible stand-in for the reference's single env (atc_gym.py:22): numpy in, numpy out, one aircraft,
    no auto-reset, reset() returns the raw observation, step() returns (obs f32[10], float, bool, {"original_state"})."""
    metadata = {'render.modes': ['human', 'rgb_array'], 'video.frames_per_second': 50}

    def __init__(self, sim_parameters=None, scenario=None, device='cuda:0'):
        self._env = BatchedAtcEnv(1, 1, sim_parameters, scenario, device=device, autoreset=False, track_actions=True,
                                  return_raw_obs=True)
        self.action_space = self._env.action_space
        self.observation_space = self._env.observation_space
        self.reward_range = self._env.reward_range
        self.timestep_limit = self._env.timestep_limit
        self.last_reward = 0
        self.done = True
        self.state = self.reset()

    def seed(self, seed=None):
        return self._env.seed(seed)

    def reset(self):
        self.done = False
        self.last_reward = 0
        self.state = self._env.reset().cpu().numpy().reshape(10)
        self._history = []                               # Airplane.position_history (model.py:51), render only
        return self.state

    def step(self, action):
        a = np.asarray(action, dtype=np.float32).reshape(1, 1, 3)
        self._history.append((float(self.state[0]), float(self.state[1])))     # model.py:123
        obs, reward, done, info = self._env.step(a)
        self.state = info['original_state'].reshape(10)
        self.done = bool(done[0])
        self.last_reward = float(reward[0])
        return obs.reshape(10), self.last_reward, self.done, {"original_state": self.state}

    @property
    def timesteps(self):
        return int(self._env.timesteps[0])

    @property
    def total_reward(self):
        return float(self._env.ep_return[0])

    @property
    def actions_taken(self):
        return int(self._env.actions_taken[0])

    @property
    def actions_per_timestep(self):
        return float(self._env.actions_per_timestep[0])

    @property
    def winning_ratio(self):
        return float(self._env.winning_ratio[0])

    def render(self, mode='human'):
        if mode == 'rgb_array':                          # trail: every 5th of the last 25 positions (atc_gym.py:440-447)
            n = len(self._history)
            idx = [i for i in range(n - 5, max(0, n - 25), -1) if i % 5 == 0]
            return self._env.render(mode, 0, np.asarray([self._history[i] for i in idx], np.float64).reshape(-1, 2),
                                    last_reward=self.last_reward)
        return self._env.render(mode)

    def close(self):
        self._env.close()


_REGISTRY = {'AtcEnv-v0': AtcGym}        # envs/__init__.py:3-5


def make(env_id, **kwargs):
    """gym.make look-alike for the id the reference registers."""
    return _REGISTRY[env_id](**kwargs)
