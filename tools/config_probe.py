#!/usr/bin/env python
"""Device-resident step rate of the other BASELINE.json configs (parity-test cases, not bench lines) for context:
4096 x 1, 16384 x 4, 16384 x 8 + wind; launches of 1024 steps, obs + original_state written."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from atc_reinforcement_learning_b200 import BatchedAtcEnv, LOWW, SimParameters

dev = torch.device('cuda', 0)
T = 1024
wind = np.random.RandomState(99).uniform(-30, 30, (16, 16, 2)).astype(np.float32)
for name, N, A, w in (('configs[1] 4096 x 1', 4096, 1, None), ('configs[2] 16384 x 4', 16384, 4, None),
                      ('configs[3] 16384 x 8 + wind', 16384, 8, wind)):
    env = BatchedAtcEnv(N, A, SimParameters(1), LOWW(random_entrypoints=True), device=dev, seed=0, wind=w)
    acts = (torch.rand(T // 20 + 1, N, A, 3, device=dev) * 2 - 1).repeat_interleave(20, 0)[:T].contiguous()
    out = env._alloc_io((T,))
    env.rollout(acts, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        env.rollout(acts, out=out)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (4 * T)
    b = (92.0 * A + 8.0) * N / us / 1e3
    print('%-28s %.3f us/step  %.2f G env-steps/s  %.1f G aircraft-steps/s  %.0f GB/s algorithmic'
          % (name, us, N / us / 1e3, N * A / us / 1e3, b), flush=True)
    del env, acts, out
