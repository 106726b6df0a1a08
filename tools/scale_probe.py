#!/usr/bin/env python
"""Step time of the rollout kernel against the batch size (A = 4): tells a latency-bound regime (time per step flat
in N) from an issue-bound one (time per step proportional to N).  Usage: python tools/scale_probe.py [N ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from atc_reinforcement_learning_b200 import BatchedAtcEnv, LOWW, SimParameters

dev = torch.device('cuda', 0)
T = 512
SIZES = [int(a) for a in sys.argv[1:]] or [1024, 2048, 4096, 8192, 12288, 16384, 20480, 32768]
for N in SIZES:
    env = BatchedAtcEnv(N, 4, SimParameters(1), LOWW(random_entrypoints=True), device=dev, seed=0, return_raw_obs=True)
    acts = torch.rand(T, N, 4, 3, device=dev) * 2 - 1
    out = env._alloc_io((T,))
    for _ in range(2):
        env.rollout(acts, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        env.rollout(acts, out=out)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (4 * T)
    print('N %6d  CTAs/SM %5.2f  %.3f us/step  %.2f G env-steps/s' % (N, N * 4 / 32 / 148, us, N / us / 1e3), flush=True)
    del env, acts, out
