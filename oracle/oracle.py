"""ctypes front-end of the CPU oracle (oracle/atc_oracle.c).  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg import this
module; the product package never does.  It reads the sector JSON files itself (a deliberately
independent ~20-line reader) so that the product's sector compiler is not on the oracle's path.
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, '_build', 'libatc_oracle.so')
SECTOR_DIR = os.path.join(os.path.dirname(HERE), 'atc_reinforcement_learning_b200', 'sectors')

TERM_NAMES = {0: 'running', 1: 'below_mva', 2: 'left_airspace', 3: 'captured', 4: 'timeout', 5: 'separation'}


def build(force=False):
    src = os.path.join(HERE, 'atc_oracle.c')
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(['make', '-s', '-C', HERE])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.atc_oracle_create.restype = C.c_void_p
        _lib.atc_oracle_num_threads.restype = C.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def load_sector(name_or_path, random_entrypoints=False):
    path = name_or_path if os.path.exists(name_or_path) else os.path.join(SECTOR_DIR, name_or_path + '.json')
    with open(path) as f:
        doc = json.load(f)
    eps = doc['entrypoints_random'] if random_entrypoints else doc['entrypoints']
    return doc, eps


class Oracle(object):
    def __init__(self, sector='LOWW', random_entrypoints=False, n_env=1, n_ac=1, dt=1.0, reward_shaping=True,
                 normalize_state=True, discrete=False, normalize_reset_obs=False, seed=0, env_index_base=0,
                 wind=None):
        doc, eps = load_sector(sector, random_entrypoints)
        rings = [np.asarray(m['ring'], np.float64) for m in doc['mvas']]
        self.ring_xy = np.ascontiguousarray(np.concatenate(rings, 0))
        self.ring_off = np.cumsum([0] + [len(r) for r in rings]).astype(np.int32)
        self.height = np.asarray([m['height'] for m in doc['mvas']], np.float64)
        r = doc['runway']
        self.runway = np.asarray([r['x'], r['y'], r['h'], r['phi_from_runway']], np.float64)
        self.entry = np.ascontiguousarray(np.asarray([[e['x'], e['y'], e['phi']] for e in eps], np.float64))
        self.level_off = np.cumsum([0] + [len(e['levels']) for e in eps]).astype(np.int32)
        self.levels = np.concatenate([np.asarray(e['levels'], np.int32) for e in eps]).astype(np.int32)
        self.n_env, self.n_ac = n_env, n_ac
        gx = gy = 0
        w = None
        if wind is not None:
            w = np.ascontiguousarray(wind, np.float32)
            gy, gx = w.shape[0], w.shape[1]
        self._wind = w
        self._h = lib().atc_oracle_create(
            C.c_int(len(rings)), _p(self.ring_xy, C.c_double), _p(self.ring_off, C.c_int), _p(self.height, C.c_double),
            _p(self.runway, C.c_double), C.c_int(len(eps)), _p(self.entry, C.c_double), _p(self.level_off, C.c_int),
            _p(self.levels, C.c_int), C.c_double(dt), C.c_int(reward_shaping), C.c_int(normalize_state),
            C.c_int(discrete), C.c_int(normalize_reset_obs), C.c_int(n_env), C.c_int(n_ac), C.c_uint64(seed),
            C.c_int64(env_index_base), C.c_int(gx), C.c_int(gy), _p(w, C.c_float))
        if not self._h:
            raise ValueError('atc_oracle_create failed')
        self._h = C.c_void_p(self._h)

    def __del__(self):
        if getattr(self, '_h', None) and _lib is not None:
            _lib.atc_oracle_destroy(self._h)
            self._h = None

    def constants(self):
        out = np.zeros(61, np.float64)
        lib().atc_oracle_constants(self._h, _p(out, C.c_double))
        return {'faf': out[0:2], 'iaf': out[2:4], 'corner1': out[4:6], 'corner2': out[6:8], 'normal': out[8:10],
                'bbox': out[10:14], 'dmax': out[14], 'faf_mva': out[15], 'phi_to': out[16], 'nmin': out[17:27],
                'nmax': out[27:37], 'tri_h': out[37:45].reshape(4, 2), 'tri_1': out[45:53].reshape(4, 2),
                'tri_2': out[53:61].reshape(4, 2)}

    def mva(self, xy):
        xy = np.ascontiguousarray(xy, np.float64)
        out = np.zeros(len(xy), np.int32)
        lib().atc_oracle_mva(self._h, C.c_int(len(xy)), _p(xy, C.c_double), _p(out, C.c_int32))
        return out

    def mva_index(self, xy):
        xy = np.ascontiguousarray(xy, np.float64)
        out = np.zeros(len(xy), np.int32)
        lib().atc_oracle_mva_index(self._h, C.c_int(len(xy)), _p(xy, C.c_double), _p(out, C.c_int32))
        return out

    def inside_corridor(self, xyhphi):
        a = np.ascontiguousarray(xyhphi, np.float64)
        out = np.zeros(len(a), np.uint8)
        lib().atc_oracle_inside_corridor(self._h, C.c_int(len(a)), _p(a, C.c_double), _p(out, C.c_uint8))
        return out

    def inside_corridor_angle(self, xyphi):
        a = np.ascontiguousarray(xyphi, np.float64)
        out = np.zeros(len(a), np.uint8)
        lib().atc_oracle_inside_corridor_angle(self._h, C.c_int(len(a)), _p(a, C.c_double), _p(out, C.c_uint8))
        return out

    def reset(self, mask=None, spawn=None):
        """spawn: None (spec RNG) or [n_env, n_ac, 5] explicit (x, y, h, phi, v).  Returns obs [n_env, n_ac, 10]
        (rows of un-reset envs are left zero)."""
        obs = np.zeros((self.n_env, self.n_ac, 10), np.float32)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        s = None if spawn is None else np.ascontiguousarray(spawn, np.float64).reshape(self.n_env, self.n_ac, 5)
        lib().atc_oracle_reset(self._h, _p(m, C.c_uint8), _p(s, C.c_double), _p(obs, C.c_float))
        return obs

    def step(self, actions, autoreset=False):
        a = np.ascontiguousarray(actions, np.float32).reshape(self.n_env, self.n_ac, 3)
        obs = np.zeros((self.n_env, self.n_ac, 10), np.float32)
        raw = np.zeros((self.n_env, self.n_ac, 10), np.float32)
        rew = np.zeros(self.n_env, np.float64)
        done = np.zeros(self.n_env, np.uint8)
        term = np.zeros(self.n_env, np.int32)
        lib().atc_oracle_step(self._h, _p(a, C.c_float), _p(obs, C.c_float), _p(raw, C.c_float), _p(rew, C.c_double),
                              _p(done, C.c_uint8), _p(term, C.c_int32), C.c_int(autoreset))
        return obs, raw, rew, done, term

    def rollout(self, actions, raw=False):
        """actions [T, n_env, n_ac, 3] float32, autoreset on.  Returns (obs, reward, done, term); with raw=True
        (obs, raw_obs, reward, done, term), raw_obs = info["original_state"] of every step (atc_gym.py:192)."""
        a = np.ascontiguousarray(actions, np.float32)
        T = a.shape[0]
        obs = np.zeros((T, self.n_env, self.n_ac, 10), np.float32)
        rew = np.zeros((T, self.n_env), np.float64)
        done = np.zeros((T, self.n_env), np.uint8)
        term = np.zeros((T, self.n_env), np.int32)
        if raw:
            raw_obs = np.zeros((T, self.n_env, self.n_ac, 10), np.float32)
            lib().atc_oracle_rollout_raw(self._h, C.c_int(T), _p(a, C.c_float), _p(obs, C.c_float),
                                         _p(raw_obs, C.c_float), _p(rew, C.c_double), _p(done, C.c_uint8),
                                         _p(term, C.c_int32))
            return obs, raw_obs, rew, done, term
        lib().atc_oracle_rollout(self._h, C.c_int(T), _p(a, C.c_float), _p(obs, C.c_float), _p(rew, C.c_double),
                                 _p(done, C.c_uint8), _p(term, C.c_int32))
        return obs, rew, done, term

    def get_state(self):
        st = np.zeros((self.n_env, self.n_ac, 5), np.float64)
        ts = np.zeros(self.n_env, np.int32)
        lib().atc_oracle_get_state(self._h, _p(st, C.c_double), _p(ts, C.c_int32))
        return st, ts

    def set_state(self, state, timesteps=None):
        st = np.ascontiguousarray(state, np.float64).reshape(self.n_env, self.n_ac, 5)
        ts = None if timesteps is None else np.ascontiguousarray(timesteps, np.int32)
        lib().atc_oracle_set_state(self._h, _p(st, C.c_double), _p(ts, C.c_int32))

    def metrics(self):
        n = self.n_env
        er, ler = np.zeros(n), np.zeros(n)
        at, ep, wr, ll = (np.zeros(n, np.int32) for _ in range(4))
        lib().atc_oracle_get_metrics(self._h, _p(er, C.c_double), _p(ler, C.c_double), _p(at, C.c_int32),
                                     _p(ep, C.c_int32), _p(wr, C.c_int32), _p(ll, C.c_int32))
        return {'ep_return': er, 'last_ep_return': ler, 'actions_taken': at, 'episodes': ep, 'win_ring': wr,
                'last_ep_len': ll}


def num_threads():
    return int(lib().atc_oracle_num_threads())


def set_num_threads(n):
    lib().atc_oracle_set_num_threads(C.c_int(int(n)))
