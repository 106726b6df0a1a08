#!/usr/bin/env python
"""Diagnostic: the same 16384 x 4 rollout on every visible GPU, one after the other, from one process — tells a slow
GPU from a slow rank.  Prints ms per 1024-step launch and the SM clock NVML reports during the run."""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from atc_reinforcement_learning_b200 import BatchedAtcEnv, LOWW, SimParameters


def clocks(idx, stop, out):
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(idx)
    while not stop.is_set():
        out.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                    pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                    pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
        time.sleep(0.01)


for d in range(torch.cuda.device_count()):
    dev = torch.device('cuda', d)
    torch.cuda.set_device(dev)
    N, A, T = 16384, 4, 1024
    env = BatchedAtcEnv(N, A, SimParameters(1), LOWW(random_entrypoints=True), device=dev, seed=0, return_raw_obs=True,
                        grid_cell=0.0625)
    acts = torch.rand(T, N, A, 3, device=dev) * 2 - 1
    out = env._alloc_io((T,))
    for _ in range(2):
        env.rollout(acts, out=out)
    torch.cuda.synchronize(dev)
    stop, samples = threading.Event(), []
    th = threading.Thread(target=clocks, args=(d, stop, samples), daemon=True)
    th.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(17)]
    ev[0].record()
    for k in range(16):
        env.rollout(acts, out=out)
        ev[k + 1].record()
    torch.cuda.synchronize(dev)
    stop.set()
    th.join()
    ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(16)]
    sm = sorted(s[0] for s in samples)
    print('gpu %d %s: ms/launch min %.3f med %.3f max %.3f | SM MHz med %d min %d | power W max %.0f | reasons %s'
          % (d, torch.cuda.get_device_name(d), min(ms), sorted(ms)[8], max(ms), sm[len(sm) // 2] if sm else -1,
             sm[0] if sm else -1, max(s[1] for s in samples) if samples else -1,
             sorted(set(hex(s[2]) for s in samples))), flush=True)
    del env, acts, out
