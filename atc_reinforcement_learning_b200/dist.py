"""Multi-GPU sharding of independent envs: one process per GPU, each rank owns a contiguous block of
the global env index space, no data-path collective.  The only communication is the gather of the per-env episode
returns for logging — the role the reference gives to SubprocVecEnv pipes + Monitor
(/root/reference/learning/atc-gym-stable-baselines.py:31-49, 73-78).  Works with NCCL (cuda tensors) and gloo (cpu)."""
import os

import torch
import torch.distributed as dist


def shard_envs(num_envs_global, rank, world_size):
    """Contiguous block partition: returns (n_local, env_index_base).  The first `rem` ranks get one extra env."""
    if num_envs_global < world_size:
        raise ValueError("fewer envs (%d) than ranks (%d)" % (num_envs_global, world_size))
    q, rem = divmod(int(num_envs_global), int(world_size))
    n_local = q + (1 if rank < rem else 0)
    base = rank * q + min(rank, rem)
    return n_local, base


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*).
    Returns (rank, world_size, local_rank).  A single process without the env vars is world_size 1, uninitialised."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local_rank


class ReturnGather(object):
    """Gathers last_ep_return[N_local] (float32 copy) from every rank.  `gather()` returns the [world * N_local] tensor
    on every rank (all_gather).  overlap=True (default on CUDA): copy + collective run on a side stream, so the step
    stream never waits for them (2 x B200: 18.31 G env-steps/s against 18.14 G with the collective enqueued in order
    on the step stream, overlap=False)."""

    def __init__(self, n_local, device, overlap=None):
        self.device = torch.device(device)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.n_local = int(n_local)
        self.send = torch.zeros(self.n_local, dtype=torch.float32, device=self.device)
        self.recv = torch.zeros(self.world * self.n_local, dtype=torch.float32, device=self.device)
        if overlap is None:
            overlap = True
        self.stream = torch.cuda.Stream(self.device) if (self.device.type == 'cuda' and overlap) else None
        self.calls = 0

    def _collect(self, last_ep_return):
        self.send.copy_(last_ep_return)
        if self.world > 1:
            if self.device.type == 'cuda':
                dist.all_gather_into_tensor(self.recv, self.send)
            else:
                parts = [torch.empty_like(self.send) for _ in range(self.world)]
                dist.all_gather(parts, self.send)
                self.recv.copy_(torch.cat(parts))
        else:
            self.recv.copy_(self.send)

    def gather(self, last_ep_return):
        self.calls += 1
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.stream):
                self._collect(last_ep_return)
        else:
            self._collect(last_ep_return)
        return self.recv

    def wait(self):
        if self.stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
        return self.recv
