#!/usr/bin/env python
"""What the host side of the box can move: every rank copies a 1 GiB pinned buffer device -> host (and host -> device)
at the same time; prints per-rank and aggregate GB/s.  Run alone and under torchrun with N ranks: if N ranks together
do not reach N x the single-rank rate, the host fabric (PCIe root complex / host memory of the VM), not the env, is what
limits the end-to-end leg of bench.py at N GPUs.
    python tools/host_fabric_probe.py ; torchrun --nproc-per-node 8 tools/host_fabric_probe.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from atc_reinforcement_learning_b200.dist import init_from_env

rank, world, local = init_from_env('nccl')
dev = torch.device('cuda', local)
torch.cuda.set_device(dev)
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device=dev)
res = {}
for name, (src, dst) in (('d2h', (d, h)), ('h2d', (h, d))):
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    gbs = 8 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9
    t = torch.tensor([gbs], device=dev, dtype=torch.float64)
    allr = [torch.zeros_like(t) for _ in range(world)]
    if world > 1:
        dist.all_gather(allr, t)
    else:
        allr = [t]
    res[name] = {'per_rank_GBps': [round(float(x.item()), 1) for x in allr], 'aggregate_GBps': round(sum(float(x.item()) for x in allr), 1)}
if rank == 0:
    print(json.dumps({'ranks': world, 'bytes_per_copy': n, **res}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
