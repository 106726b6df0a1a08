import os, sys, torch
sys.path.insert(0, '/root/repo')
from atc_reinforcement_learning_b200 import BatchedAtcEnv, LOWW, SimParameters
dev = torch.device('cuda', 0)
N = 16384
env = BatchedAtcEnv(N, 4, SimParameters(1), LOWW(random_entrypoints=True), device=dev, seed=0, return_raw_obs=True)
for T in (4, 8, 16, 32, 64, 128, 256):
    acts = torch.rand(T, N, 4, 3, device=dev) * 2 - 1
    out = env._alloc_io((T,))
    for _ in range(3):
        env.rollout(acts, out=out)
    torch.cuda.synchronize()
    reps = max(4, 2048 // T)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        env.rollout(acts, out=out)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print('T %4d  %8.1f us/launch  %.3f us/step' % (T, us, us / T), flush=True)
