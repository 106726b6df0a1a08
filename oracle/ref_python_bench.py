#!/usr/bin/env python
"""Times the reference's OWN Python step() on this host (TEST / BENCH INFRASTRUCTURE — never imported by the product).

The unmodified reference package `envs/` is imported from `baseline/_ref/` (a git-ignored copy that
`__graft_entry__.build()` makes from /root/reference when it is present; it travels to the GPU box with the
snapshot) under the import stand-ins of oracle/standins (gym / shapely / pyglet are not installed).  Two loops:

  fixed   the exact loop of /root/reference/learning/atc-gym-compute-performance.py:7-16 — gym.make('AtcEnv-v0'),
          one sampled action repeated, no reset on done — one process;
  random  the demo's policy (/root/reference/learning/atc-gym-demo.py:18-19: a new sampled action every 20 steps)
          with reset on done, in P independent processes — the reference's own scaling model is one env per process
          (SubprocVecEnv, /root/reference/learning/atc-gym-stable-baselines.py:76-78); the aggregate rate is reported.

Each loop is calibrated to run for about --seconds after a JIT warm-up (numba compiles the 12 helpers on first use).
Prints one JSON object.  bench.py runs this file in a subprocess for cpu_baseline.reference_python.
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get('ATC_REFERENCE_ROOT', os.path.join(ROOT, 'baseline', '_ref'))


def _import_reference():
    sys.path.insert(0, os.path.join(HERE, 'standins'))
    sys.path.insert(1, REF)
    import gym
    import envs.atc.atc_gym  # noqa: F401  (registers AtcEnv-v0)
    return gym


def loop_fixed(seconds):
    gym = _import_reference()
    env = gym.make('AtcEnv-v0')
    env.reset()
    nextaction = env.action_space.sample()
    for _ in range(2000):                                   # JIT warm-up
        env.step(nextaction)
    t0 = time.time()
    for _ in range(5000):
        env.step(nextaction)
    rate = 5000 / (time.time() - t0)
    num = max(10000, int(rate * seconds))
    env.reset()
    t0 = time.time()
    for i in range(num):
        state, reward, done, info = env.step(nextaction)
    t1 = time.time()
    return num, t1 - t0


def _random_worker(seconds, seed, q, barrier):
    import random
    import numpy as np
    gym = _import_reference()
    random.seed(seed)
    np.random.seed(seed)
    env = gym.make('AtcEnv-v0')
    env.reset()
    a = env.action_space.sample()
    for _ in range(2000):
        _, _, done, _ = env.step(a)
        if done:
            env.reset()
    barrier.wait()                                          # every process has compiled: time them together
    n, t0 = 0, time.time()
    t_end = t0 + seconds
    while True:
        for _ in range(50):                                 # 50 x 20 steps between clock reads
            a = env.action_space.sample()
            for _ in range(20):
                _, _, done, _ = env.step(a)
                if done:
                    env.reset()
        n += 1000
        if time.time() >= t_end:
            break
    q.put((n, time.time() - t0))


def loop_random(seconds, procs):
    ctx = mp.get_context('fork')
    q = ctx.Queue()
    barrier = ctx.Barrier(procs)
    ps = [ctx.Process(target=_random_worker, args=(seconds, 100 + i, q, barrier)) for i in range(procs)]
    for p in ps:
        p.start()
    res = [q.get() for _ in ps]
    for p in ps:
        p.join()
    return sum(n / s for n, s in res), sum(n for n, _ in res)


def cpu_model():
    try:
        with open('/proc/cpuinfo') as f:
            for line in f:
                if line.startswith('model name'):
                    return line.split(':', 1)[1].strip()
    except OSError:
        pass
    return 'unknown'


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--seconds', type=float, default=4.0)
    ap.add_argument('--procs', type=int, default=0, help='0 = all cores this process may use')
    args = ap.parse_args()
    if not os.path.isdir(os.path.join(REF, 'envs', 'atc')):
        print(json.dumps({'unavailable': 'no reference copy under %s' % REF}))
        return
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    os.environ.setdefault('NUMBA_NUM_THREADS', '1')
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    procs = args.procs or cores
    num, sec = loop_fixed(args.seconds)
    agg, total = loop_random(args.seconds, procs)
    print(json.dumps({
        'steps_per_s_1core': num / sec, 'steps_1core': num,
        'loop_1core': 'learning/atc-gym-compute-performance.py:7-16 verbatim (fixed action, no reset on done)',
        'steps_per_s_allcores': agg, 'steps_allcores': total, 'cores': procs,
        'loop_allcores': '%d independent processes, one AtcGym each, sampled action every 20 steps, reset on done' % procs,
        'cpu_model': cpu_model(), 'aircraft_per_env': 1,
        'reference': 'unmodified envs/atc from fvalka/atc-reinforcement-learning under import stand-ins, numba JIT'}))


if __name__ == '__main__':
    main()
