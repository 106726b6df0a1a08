#!/usr/bin/env python
"""Generate golden vectors from the LIVE, UNMODIFIED reference (test infrastructure).

Runs only in the build container (needs /root/reference); output is committed under
tests/golden/ so the GPU box, which has no /root/reference, can check against it.

    PYTHONPATH=oracle/standins:/root/reference python oracle/make_golden.py

What is recorded (everything is an output of reference code, nothing is re-derived):
  * kat.json            — the reference's own 8 unit-test cases (envs/atc/model_test.py:10-92)
                          and the constants of AtcGym.__init__ / Corridor.__init__ for both scenarios
  * geometry_*.npz      — Airspace.get_mva_height (model.py:282-292) and Runway.inside_corridor
                          (model.py:188-231,248) on random + adversarial points
  * trace_*.npz         — AtcGym.step()/reset() (atc_gym.py:128-192,337-365) driven in lockstep over E
                          independent env objects with recorded float32 action streams; per step the raw
                          float64 airplane state, both observations, reward, done, the terminal-branch
                          flags, metrics, and the spawn the reference chose on reset
"""
import json
import math
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'standins'))
sys.path.insert(1, os.environ.get('ATC_REFERENCE_ROOT', '/root/reference'))

import envs.atc.atc_gym as atc_gym  # noqa: E402
from envs.atc import model, scenarios  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')
os.makedirs(OUT, exist_ok=True)

F_BELOW, F_OUTSIDE, F_WIN, F_TIMEOUT = 1, 2, 4, 8


def mva_or_minus1(airspace, x, y):
    try:
        return int(airspace.get_mva_height(x, y))
    except ValueError:
        return -1


def flags_of(env):
    """Terminal-branch flags, evaluated by reference code on the post-step airplane."""
    a = env._airplane
    f = 0
    m = mva_or_minus1(env._airspace, a.x, a.y)
    if m < 0:
        f |= F_OUTSIDE
    elif a.h < m:
        f |= F_BELOW
    if env._runway.inside_corridor(a.x, a.y, a.h, a.phi):
        f |= F_WIN
    if env.timesteps > env.timestep_limit:
        f |= F_TIMEOUT
    return f


def scenario_constants(scn, env):
    c = scn.runway.corridor
    return {
        'mva_heights': [int(m.height) for m in scn.mvas],
        'mva_ring_sizes': [int(len(m.area_as_list)) for m in scn.mvas],
        'mva_bounds': [list(map(float, m.outer_bounds)) for m in scn.mvas],
        'runway': [float(scn.runway.x), float(scn.runway.y), float(scn.runway.h), float(scn.runway.phi_from_runway)],
        'phi_to_runway': float(scn.runway.phi_to_runway),
        'faf': [float(c.faf[0][0]), float(c.faf[1][0])],
        'iaf': [float(c.iaf[0][0]), float(c.iaf[1][0])],
        'corner1': [float(c.corner1[0][0]), float(c.corner1[1][0])],
        'corner2': [float(c.corner2[0][0]), float(c.corner2[1][0])],
        'faf_iaf_normal': [float(c._faf_iaf_normal[0][0]), float(c._faf_iaf_normal[1][0])],
        'corridor_horizontal': np.asarray(c.corridor_horizontal_list, dtype=np.float64).tolist(),
        'corridor1': np.asarray(c.corridor1_list, dtype=np.float64).tolist(),
        'corridor2': np.asarray(c.corridor2_list, dtype=np.float64).tolist(),
        'bbox': [float(env._world_x_min), float(env._world_y_min), float(env._world_x_max), float(env._world_y_max)],
        'world_max_distance': float(env._world_max_distance),
        'faf_mva': float(env._faf_mva),
        'norm_min': [float(v) for v in env.normalization_state_min],
        'norm_max': [float(v) for v in env.normalization_state_max],
        'entrypoints': [[float(e.x), float(e.y), float(e.phi), [int(l) for l in e.levels]] for e in scn.entrypoints],
    }


# --------------------------------------------------------------------------------------- KATs
def make_kat():
    kat = {}
    # the reference's own unit tests (model_test.py) — fixture: 5 MVAs, runway (20, 20, 0, 180)
    mvas = [
        model.MinimumVectoringAltitude(scenarios.shape.Polygon([(15, 0), (35, 0), (35, 26)]), 3500),
        model.MinimumVectoringAltitude(scenarios.shape.Polygon([(15, 0), (35, 26), (35, 30), (15, 30), (15, 27.8)]), 2400),
        model.MinimumVectoringAltitude(scenarios.shape.Polygon([(15, 30), (35, 30), (35, 40), (15, 40)]), 4000),
        model.MinimumVectoringAltitude(scenarios.shape.Polygon([(0, 10), (15, 0), (15, 28.7), (0, 17)]), 8000),
        model.MinimumVectoringAltitude(scenarios.shape.Polygon([(0, 17), (15, 28.7), (15, 40), (0, 32)]), 6500)]
    rwy = model.Runway(20, 20, 0, 180)
    asp = model.Airspace(mvas, rwy)
    faf_mva = asp.get_mva_height(rwy.corridor.faf[0][0], rwy.corridor.faf[1][0])
    kat['model_test'] = {
        'runway': [20, 20, 0, 180],
        'get_mva_height_34_1': int(asp.get_mva_height(34, 1)),
        'faf_mva': int(faf_mva),
        'inside_corridor': [
            [19, 10, faf_mva + 300, 30, bool(rwy.inside_corridor(19, 10, faf_mva + 300, 30))],
            [19, 10, faf_mva, 330, bool(rwy.inside_corridor(19, 10, faf_mva, 330))]],
        'inside_corridor_angle': [
            [21, 10, 30, bool(rwy.corridor._inside_corridor_angle(21, 10, 30))],
            [19, 10, 340, bool(rwy.corridor._inside_corridor_angle(19, 10, 340))],
            [19, 10, 190, bool(rwy.corridor._inside_corridor_angle(19, 10, 190))],
            [21, 10, 340, bool(rwy.corridor._inside_corridor_angle(21, 10, 340))]],
        'bbox': [float(v) for v in asp.get_bounding_box()],
    }
    loww = scenarios.LOWW()
    env = atc_gym.AtcGym(scenario=loww)
    kat['LOWW'] = scenario_constants(loww, env)
    kat['LOWW_random_entrypoints'] = scenario_constants(scenarios.LOWW(random_entrypoints=True), env)['entrypoints']
    simple = scenarios.SimpleScenario()
    env_s = atc_gym.AtcGym(scenario=simple)
    kat['SimpleScenario'] = scenario_constants(simple, env_s)
    kat['K0_reset_obs'] = [float(v) for v in env.reset()]

    # K7 corridor gates at the corridor centroids
    c = loww.runway.corridor
    cen1 = np.mean(np.asarray(c.corridor1_list)[:3], axis=0)
    cen2 = np.mean(np.asarray(c.corridor2_list)[:3], axis=0)
    rows = []
    for cen in (cen1, cen2):
        for phi in (340, 340.1, 350, 20, 339.9, 330, 25.1, 160, 300, 294.9, 295.0, 25.0, 0, 360, 700, -20):
            for h in (3000, 6000, 3579, 3580):
                rows.append([float(cen[0]), float(cen[1]), float(h), float(phi),
                             bool(loww.runway.inside_corridor(float(cen[0]), float(cen[1]), h, phi))])
    kat['K7_corridor'] = rows

    # K10 seeded spawns
    env_r = atc_gym.AtcGym(scenario=scenarios.LOWW(random_entrypoints=True))
    env_r.seed(0)
    sp = []
    for _ in range(4):
        env_r.reset()
        sp.append([env_r._airplane.x, env_r._airplane.y, env_r._airplane.h, env_r._airplane.phi])
    kat['K10_seeded_spawns'] = sp

    # K8 discrete action space
    env_d = atc_gym.AtcGym(sim_parameters=model.SimParameters(1, discrete_action_space=True))
    env_d.reset()
    s, r, d, info = env_d.step(np.array([10, 100, 180]))
    kat['K8_discrete'] = {'nvec': [int(v) for v in env_d.action_space.nvec], 'reward': float(r),
                          'raw_obs': [float(v) for v in info['original_state']], 'obs': [float(v) for v in s]}
    with open(os.path.join(OUT, 'kat.json'), 'w') as f:
        json.dump(kat, f, indent=1, sort_keys=True)
    print('kat.json written')


# --------------------------------------------------------------------------------------- geometry
def make_geometry(name, scn, n_random, seed):
    rng = np.random.RandomState(seed)
    asp, rwy = scn.airspace, scn.runway
    bx = asp.get_bounding_box()
    pts = [np.stack([rng.uniform(bx[0] - 2, bx[2] + 2, n_random), rng.uniform(bx[1] - 2, bx[3] + 2, n_random)], 1)]
    # adversarial: vertices, points on vertex y-levels, edge midpoints, +- tiny offsets, bbox corners/edges
    adv = []
    for m in scn.mvas:
        ring = np.asarray(m.area_as_list)
        for i in range(len(ring) - 1):
            p, q = ring[i], ring[i + 1]
            mid = 0.5 * (p + q)
            for base in (p, mid, p + 0.25 * (q - p)):
                for dx in (0.0, 1e-9, -1e-9, 1e-4, -1e-4, 0.3, -0.3):
                    for dy in (0.0, 1e-9, -1e-9):
                        adv.append([base[0] + dx, base[1] + dy])
        b = m.outer_bounds
        for x in (b[0], b[2], 0.5 * (b[0] + b[2])):
            for y in (b[1], b[3], 0.5 * (b[1] + b[3])):
                adv.append([x, y])
    pts.append(np.asarray(adv, dtype=np.float64))
    pts = np.concatenate(pts, 0)
    mva = np.array([mva_or_minus1(asp, float(x), float(y)) for x, y in pts], dtype=np.int32)

    # corridor: dense around the corridor triangle with random h / phi (many inside)
    c = rwy.corridor
    tri = np.asarray(c.corridor_horizontal_list)
    lo, hi = tri.min(0) - 0.5, tri.max(0) + 0.5
    nc = n_random
    cx = rng.uniform(lo[0], hi[0], nc)
    cy = rng.uniform(lo[1], hi[1], nc)
    ch = rng.uniform(0, 8000, nc)
    cphi = np.where(rng.uniform(size=nc) < 0.7,
                    rwy.phi_to_runway + rng.uniform(-60, 60, nc), rng.uniform(-90, 450, nc))
    # exact-boundary headings
    cphi[: nc // 20] = rwy.phi_to_runway + rng.choice([0.0, 45.0, -45.0, 360.0, -360.0, 45.000001, -45.000001], nc // 20)
    # points on the triangle vertices / centreline
    k = 0
    for p in list(tri[:3]) + [0.5 * (tri[0] + tri[1]), 0.5 * (tri[0] + tri[2]), np.array([c.iaf[0][0], c.iaf[1][0]]),
                              0.5 * (tri[0] + np.array([c.iaf[0][0], c.iaf[1][0]]))]:
        for d in (0.0, 1e-9, -1e-9):
            cx[nc // 20 + k] = p[0] + d
            cy[nc // 20 + k] = p[1]
            k += 1
    inside = np.array([bool(rwy.inside_corridor(float(x), float(y), float(h), float(p)))
                       for x, y, h, p in zip(cx, cy, ch, cphi)], dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, 'geometry_%s.npz' % name), pts=pts, mva=mva,
                        corr=np.stack([cx, cy, ch, cphi], 1), inside=inside)
    print('geometry_%s: %d points (%d outside), corridor %d (%d inside)' % (
        name, len(pts), int((mva < 0).sum()), nc, int(inside.sum())))


# --------------------------------------------------------------------------------------- traces
class Recorder(object):
    """Drives E reference envs in lockstep and records everything step() and reset() produce."""

    def __init__(self, envs_, T):
        self.envs = envs_
        E = len(envs_)
        self.T, self.E = T, E
        self.action = np.zeros((T, E, 3), np.float32)
        self.state = np.zeros((T, E, 5), np.float64)
        self.obs = np.zeros((T, E, 10), np.float32)
        self.raw_obs = np.zeros((T, E, 10), np.float32)
        self.reward = np.zeros((T, E), np.float64)
        self.done = np.zeros((T, E), np.uint8)
        self.flags = np.zeros((T, E), np.uint8)
        self.timesteps = np.zeros((T, E), np.int32)
        self.actions_taken = np.zeros((T, E), np.int32)
        self.total_reward = np.zeros((T, E), np.float64)
        self.winning_ratio = np.zeros((T, E), np.float64)
        self.spawn = np.zeros((T, E, 5), np.float64)      # valid where done: state after the reference's reset()
        self.reset_obs = np.zeros((T, E, 10), np.float32)
        self.init_state = np.zeros((E, 5), np.float64)
        self.init_obs = np.zeros((E, 10), np.float32)
        self.init_timesteps = np.zeros((E,), np.int32)

    @staticmethod
    def _st(env):
        a = env._airplane
        return [a.x, a.y, a.h, a.phi, a.v]

    def begin(self, init_obs):
        for e, env in enumerate(self.envs):
            self.init_state[e] = self._st(env)
            self.init_obs[e] = init_obs[e]
            self.init_timesteps[e] = env.timesteps

    def step(self, t, actions, reset_on_done=True):
        for e, env in enumerate(self.envs):
            a = np.asarray(actions[e], dtype=np.float32)
            self.action[t, e] = a
            s, r, d, info = env.step(a)
            self.state[t, e] = self._st(env)
            self.obs[t, e] = s
            self.raw_obs[t, e] = info['original_state']
            self.reward[t, e] = r
            self.done[t, e] = d
            self.flags[t, e] = flags_of(env)
            self.timesteps[t, e] = env.timesteps
            self.actions_taken[t, e] = env.actions_taken
            self.total_reward[t, e] = env.total_reward
            if d and reset_on_done:
                ro = env.reset()
                self.spawn[t, e] = self._st(env)
                self.reset_obs[t, e] = ro
            self.winning_ratio[t, e] = env.winning_ratio

    def save(self, name, meta):
        np.savez_compressed(
            os.path.join(OUT, 'trace_%s.npz' % name), meta=json.dumps(meta),
            action=self.action, state=self.state, obs=self.obs, raw_obs=self.raw_obs, reward=self.reward,
            done=self.done, flags=self.flags, timesteps=self.timesteps, actions_taken=self.actions_taken,
            total_reward=self.total_reward, winning_ratio=self.winning_ratio, spawn=self.spawn,
            reset_obs=self.reset_obs, init_state=self.init_state, init_obs=self.init_obs,
            init_timesteps=self.init_timesteps)
        nd = int(self.done.sum())
        fl = self.flags[self.done > 0]
        print('trace_%s: T=%d E=%d dones=%d below=%d outside=%d win=%d timeout=%d' % (
            name, self.T, self.E, nd, int(((fl & F_BELOW) > 0).sum()), int(((fl & F_OUTSIDE) > 0).sum()),
            int(((fl & F_WIN) > 0).sum()), int(((fl & F_TIMEOUT) > 0).sum())))


def new_envs(E, scenario_name='LOWW', random_entrypoints=False, dt=1, shaping=True, normalize=True, discrete=False):
    out = []
    for _ in range(E):
        scn = getattr(scenarios, scenario_name)(random_entrypoints=random_entrypoints)
        out.append(atc_gym.AtcGym(sim_parameters=model.SimParameters(
            dt, reward_shaping=shaping, normalize_state=normalize, discrete_action_space=discrete), scenario=scn))
    return out


def meta_of(scenario_name, random_entrypoints, dt, shaping, normalize, discrete, note):
    return dict(scenario=scenario_name, random_entrypoints=random_entrypoints, dt=dt, reward_shaping=shaping,
                normalize_state=normalize, discrete=discrete, note=note)


def trace_random(name, E, T, seed, repeat, amp=1.0, h_bias=False, **kw):
    """U(-amp, amp)^3 float32 actions re-sampled every `repeat` steps (atc-gym-demo.py:18-19 cadence is 20)."""
    rng = np.random.RandomState(seed)
    random.seed(seed)
    es = new_envs(E, **kw)
    rec = Recorder(es, T)
    rec.begin([env.reset() for env in es])
    act = None
    for t in range(T):
        if t % repeat == 0:
            if kw.get('discrete'):
                act = np.stack([rng.randint(0, 20, E), rng.randint(0, 380, E), rng.randint(0, 360, E)], 1).astype(np.float32)
            else:
                act = rng.uniform(-amp, amp, (E, 3)).astype(np.float32)
                if h_bias:      # descend hard: reaches the below-MVA branch organically
                    act[:, 1] = rng.uniform(-1.0, -0.7, E).astype(np.float32)
        rec.step(t, act)
    rec.save(name, meta_of(kw.get('scenario_name', 'LOWW'), kw.get('random_entrypoints', False), kw.get('dt', 1),
                           kw.get('shaping', True), kw.get('normalize', True), kw.get('discrete', False),
                           'uniform(-%g,%g) actions, repeat %d, seed %d' % (amp, amp, repeat, seed)))


def guided_action(env, rng, sloppy):
    """A hand-written approach controller, used only to produce action streams that reach the
    capture branch (win) and near-misses.  It reads the reference env; the recorded float32 actions are
    the only thing that enters the golden file."""
    a = env._airplane
    c = env._runway.corridor
    faf = np.array([c.faf[0][0], c.faf[1][0]])
    iaf = np.array([c.iaf[0][0], c.iaf[1][0]])
    axis = (iaf - faf) / np.linalg.norm(iaf - faf)
    p = np.array([a.x, a.y])
    gate = faf + axis * (2.0 + sloppy)          # aim for a point inside the corridor triangle
    far = faf + axis * 14.0
    along = float(np.dot(p - faf, axis))
    lateral = float(np.linalg.norm((p - faf) - along * axis))
    target = gate if (along > 2.0 and lateral < 1.5) or np.linalg.norm(p - far) < 2.0 or env.__dict__.get('_g_phase', 0) else far
    if target is gate:
        env._g_phase = 1
    d = target - p
    hdg = math.degrees(math.atan2(d[0], d[1])) % 360.0      # compass heading towards target
    # avoid the reference's no-wrap long turns: choose the representation of hdg closest to the current phi
    k = round((a.phi - hdg) / 360.0)
    hdg_cmd = hdg + 360.0 * k
    dist = float(np.linalg.norm(faf - p))
    h_cmd = min(max(2800.0 + 250.0 * max(dist - 3.0, 0.0), 2800.0), 15000.0)
    v_cmd = 220.0 if dist > 15 else 180.0
    act = np.array([(v_cmd - 100.0) / 100.0 - 1.0, h_cmd / 19000.0 - 1.0, hdg_cmd / 180.0 - 1.0])
    act += rng.normal(0, 0.002 * (1 + 5 * sloppy), 3)
    return act.astype(np.float32)


def trace_guided(name, E, T, seed, **kw):
    rng = np.random.RandomState(seed)
    random.seed(seed)
    es = new_envs(E, **kw)
    rec = Recorder(es, T)
    rec.begin([env.reset() for env in es])
    sloppy = [0.0 if e % 2 == 0 else 0.8 for e in range(E)]
    for t in range(T):
        acts = []
        for e, env in enumerate(es):
            if env.timesteps == 0:
                env._g_phase = 0
            acts.append(guided_action(env, rng, sloppy[e]))
        rec.step(t, acts)
    rec.save(name, meta_of(kw.get('scenario_name', 'LOWW'), kw.get('random_entrypoints', False), kw.get('dt', 1),
                           kw.get('shaping', True), kw.get('normalize', True), False, 'guided approach controller'))


def trace_scripted(name, T, actions, inject=None, set_timesteps=None, reset_on_done=True, **kw):
    """One constant action per env; optional injected airplane state (K4) / timestep counter (K5)."""
    E = len(actions)
    random.seed(0)
    es = new_envs(E, **kw)
    rec = Recorder(es, T)
    init = [env.reset() for env in es]
    for e, env in enumerate(es):
        if inject is not None and inject[e] is not None:
            x, y, h, phi, v = inject[e]
            env._airplane = model.Airplane(env._sim_parameters, "FLT01", x, y, h, phi, v)
            init[e] = env._get_state(0)
        if set_timesteps is not None and set_timesteps[e] is not None:
            env.timesteps = set_timesteps[e]
    rec.begin(init)
    for t in range(T):
        rec.step(t, actions, reset_on_done=reset_on_done)
    rec.save(name, meta_of(kw.get('scenario_name', 'LOWW'), False, kw.get('dt', 1), kw.get('shaping', True),
                           kw.get('normalize', True), kw.get('discrete', False), 'scripted constant actions'))


def main():
    make_kat()
    make_geometry('LOWW', scenarios.LOWW(), 20000, 1)
    make_geometry('SimpleScenario', scenarios.SimpleScenario(), 8000, 2)

    trace_random('loww_rand20', E=8, T=1500, seed=11, repeat=20)
    trace_random('loww_rand1', E=8, T=600, seed=12, repeat=1)
    trace_random('loww_invalid', E=8, T=600, seed=13, repeat=7, amp=1.3)
    trace_random('loww_entry9', E=8, T=1500, seed=14, repeat=20, random_entrypoints=True)
    trace_random('loww_descend', E=8, T=700, seed=20, repeat=20, random_entrypoints=True, h_bias=True)
    trace_random('loww_dt5_plain', E=8, T=400, seed=15, repeat=10, dt=5.0, shaping=False, normalize=False)
    trace_random('loww_discrete', E=8, T=800, seed=16, repeat=20, discrete=True)
    trace_random('simple_rand', E=4, T=40, seed=17, repeat=5, scenario_name='SimpleScenario')
    trace_guided('loww_guided', E=4, T=2800, seed=18)
    trace_guided('loww_guided_entry9', E=4, T=2800, seed=19, random_entrypoints=True)
    # K2 below MVA, K3 left airspace, K6 invalid, K1 zeros — constant actions, no reset (keeps simulating after done)
    trace_scripted('loww_scripted', T=320, reset_on_done=False,
                   actions=[[0, -1, -0.5], [1, 0, 0.5], [1.5, -1.2, 3.0], [0, 0, 0]])
    # K4 capture on the first step from an injected state, and neighbours that must NOT capture
    cen1 = (47.43529673, 34.08486922)
    a_k4 = [0.5, 3000 / 19000 - 1, 350 / 180 - 1]
    trace_scripted('loww_capture', T=3, reset_on_done=True,
                   actions=[a_k4, a_k4, [0.5, 3000 / 19000 - 1, 330 / 180 - 1], a_k4],
                   inject=[(cen1[0], cen1[1], 3000, 350, 250), (cen1[0], cen1[1], 6000, 350, 250),
                           (cen1[0], cen1[1], 3000, 330, 250), (49.31468197, 34.76890951, 3000, 350, 250)])
    # K5 timeout: start at t=5990
    trace_scripted('loww_timeout', T=14, reset_on_done=False, actions=[[0, 0, -0.5], [0, 0, -0.5]],
                   set_timesteps=[5990, 5999])


if __name__ == '__main__':
    main()
