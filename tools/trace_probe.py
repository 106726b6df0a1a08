#!/usr/bin/env python
"""Where a rollout launch spends its time, from %globaltimer stamps of the first warp pair of every CTA (development
build with -DATC_TRACE, see csrc/atc_kernels.cu `g_trace`): launch gap between two back-to-back launches, staging of the
compact grid, time per step along the launch, spread of the CTA end times (the tail).  Usage (GPU box):
    bash tools/build_variant.sh trace -DATC_TRACE && cp variants/trace.so atc_reinforcement_learning_b200/csrc/libatc_b200.so
    python tools/trace_probe.py [T ...]"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from atc_reinforcement_learning_b200 import BatchedAtcEnv, LOWW, SimParameters
from atc_reinforcement_learning_b200 import _native as nat

ROWS, COLS = 2176, 1028
dev = torch.device('cuda', 0)
N, A = 16384, 4
env = BatchedAtcEnv(N, A, SimParameters(1), LOWW(random_entrypoints=True), device=dev, seed=0, return_raw_obs=True)
lib = nat.lib()
lib.atc_debug_trace.argtypes = [C.c_void_p, C.c_int]
res = {}
for T in [int(x) for x in sys.argv[1:]] or [20, 128, 1024]:
    T0 = T - (T & 1)                                   # even: slot 0; T0 + 1: slot 1
    acts = (torch.rand((T0 + 20) // 20 + 1, N, A, 3, device=dev) * 2 - 1).repeat_interleave(20, 0)[:T0 + 1].contiguous()
    out = env._alloc_io((T0 + 1,))
    o0 = {k: v[:T0] for k, v in out.items()}
    for _ in range(3):
        env.rollout(acts[:T0], out=o0)
        env.rollout(acts, out=out)
    torch.cuda.synchronize()
    lib.atc_debug_trace(None, 1)
    for _ in range(2):                                 # the pair looked at: the last (T0, T0 + 1) launches
        env.rollout(acts[:T0], out=o0)
        env.rollout(acts, out=out)
    raw = np.zeros(2 * ROWS * COLS * 12, dtype=np.uint8)          # uint64 stamps, then uint32 path masks
    assert lib.atc_debug_trace(raw.ctypes.data_as(C.c_void_p), 0) == 0
    buf = raw[:2 * ROWS * COLS * 8].view(np.uint64).reshape(2, ROWS, COLS)
    mask = raw[2 * ROWS * COLS * 8:].view(np.uint32).reshape(2, ROWS, COLS)
    ll = env.last_launch
    g = ll['grid']
    a, b = buf[0, :g].astype(np.int64), buf[1, :g].astype(np.int64)   # launch k (T0 steps), launch k + 1 (T0 + 1 steps)
    t0 = a[:, 0].min()
    us = lambda x: float(x) / 1e3
    steps_a = a[:, 2:2 + T0]
    d = np.diff(np.concatenate([a[:, 1:2], steps_a], 1), axis=1)      # per-step time of pair 0 of every CTA [g, T0]
    prof = {}
    edges = [0, 1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024]
    for lo, hi in zip(edges[:-1], edges[1:]):
        if lo < T0:
            prof['steps %d-%d' % (lo, min(hi, T0) - 1)] = round(us(d[:, lo:min(hi, T0)].mean()), 3)
    end_a = np.maximum(a[:, COLS - 1], a[:, COLS - 2])
    res[T0] = {
        'kernel': ll['name'], 'grid': g,
        'launch_k_us': {'entry spread (last CTA entry - first)': us(a[:, 0].max() - t0),
                        'staging (entry -> staged), mean / max': [us((a[:, 1] - a[:, 0]).mean()), us((a[:, 1] - a[:, 0]).max())],
                        'first mover step done after staging, mean': us((a[:, 2] - a[:, 1]).mean()),
                        'CTA end (pair 0) relative to first entry: min / mean / max': [us(end_a.min() - t0), us(end_a.mean() - t0), us(end_a.max() - t0)],
                        'observer done after mover done (pair 0), mean': us((a[:, COLS - 2] - a[:, COLS - 1]).mean())},
        'gap_us: last pair-0 end of launch k -> first CTA entry of launch k+1': us(b[:, 0].min() - end_a.max()),
        'launch_k+1 first entry - launch k first entry (launch period)': us(b[:, 0].min() - t0),
        'us_per_step_by_phase (mean over CTAs, pair 0)': prof,
        'steady us/step x T': round(us(d[:, T0 // 2:].mean()) * T0, 1),
    }
    np.savez_compressed(os.path.join(os.environ.get('TRACE_OUT', 'gpurun_out'), 'trace_%s_T%d.npz' % (os.environ.get('TRACE_TAG', 'default'), T0)),
                        a=a, b=b, mask=mask[0, :g], done=out['done'][:T0].cpu().numpy(), term=out['term'][:T0].cpu().numpy())
    del acts, out
print(json.dumps(res, indent=1))
