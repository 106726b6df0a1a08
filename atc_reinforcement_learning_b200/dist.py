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
        opts = None
        if backend == 'nccl':
            torch.cuda.set_device(local_rank)
            # The only collectives of this path are tiny (the 64 KiB episode-return log, a few scalars): one CTA each.
            # The rollout kernel keeps one CTA resident on (almost) every SM for the whole launch, so a multi-CTA
            # NCCL kernel cannot run beside it — it would wait for the launch to end and then delay the next one
            # (measured on 2 x B200: ~70 us per launch, 4.4 % of the step rate); a single CTA fits on a spare SM.
            try:
                opts = dist.ProcessGroupNCCL.Options()
                opts.config.max_ctas = 1
                opts.config.min_ctas = 1
            except Exception:
                opts = None
        if opts is not None:
            dist.init_process_group(backend=backend, rank=rank, world_size=world, pg_options=opts)
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local_rank


class ReturnGather(object):
    """Gathers last_ep_return[N_local] (float32 copy) from every rank.  `gather()` returns the [sum of N_local] tensor
    (global env order) on every rank (all_gather).  Ranks may own different numbers of envs (shard_envs gives the first
    ranks one more when the global count does not divide): the per-rank counts are exchanged once at construction,
    every rank sends a block padded to the largest count, and the padding is dropped on receipt.

    overlap=True (default on CUDA): the 4-byte-per-env snapshot of the log is taken in order on the step stream (so it
    is a consistent cut between two launches), the collective runs on a side stream and the step stream never waits
    for it.  Two send buffers alternate, so a snapshot is never overwritten while its collective is still in flight
    (gather() makes the step stream wait for the collective issued two calls earlier, which has long finished)."""

    def __init__(self, n_local, device, overlap=None):
        self.device = torch.device(device)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.n_local = int(n_local)
        cuda = self.device.type == 'cuda'
        if self.world > 1:
            mine = torch.tensor([self.n_local], dtype=torch.int64, device=self.device)
            counts = [torch.zeros_like(mine) for _ in range(self.world)]
            dist.all_gather(counts, mine)
            self.counts = [int(c.item()) for c in counts]
        else:
            self.counts = [self.n_local]
        self.n_max = max(self.counts)
        self.equal = all(c == self.n_max for c in self.counts)
        self.send = [torch.zeros(self.n_max, dtype=torch.float32, device=self.device) for _ in range(2)]
        self.padded = torch.zeros(self.world * self.n_max, dtype=torch.float32, device=self.device)
        self.recv = self.padded if self.equal else torch.zeros(sum(self.counts), dtype=torch.float32, device=self.device)
        if overlap is None:
            overlap = True
        self.stream = torch.cuda.Stream(self.device) if (cuda and overlap) else None
        self.done = [torch.cuda.Event(), torch.cuda.Event()] if self.stream is not None else None
        self.calls = 0

    def _collect(self, send):
        if self.world > 1:
            if self.device.type == 'cuda':
                dist.all_gather_into_tensor(self.padded, send)
            else:
                parts = [torch.empty_like(send) for _ in range(self.world)]
                dist.all_gather(parts, send)
                self.padded.copy_(torch.cat(parts))
        else:
            self.padded.copy_(send)
        if not self.equal:
            off = 0
            for r, c in enumerate(self.counts):
                self.recv[off:off + c].copy_(self.padded[r * self.n_max:r * self.n_max + c])
                off += c

    def gather(self, last_ep_return):
        k = self.calls & 1
        self.calls += 1
        send = self.send[k]
        if self.stream is not None:
            cur = torch.cuda.current_stream(self.device)
            if self.calls > 2:
                cur.wait_event(self.done[k])                 # the collective that last read send[k]
            send[:self.n_local].copy_(last_ep_return)        # snapshot, in order on the step stream
            self.stream.wait_stream(cur)
            with torch.cuda.stream(self.stream):
                self._collect(send)
                self.done[k].record(self.stream)
        else:
            send[:self.n_local].copy_(last_ep_return)
            self._collect(send)
        return self.recv

    def wait(self):
        if self.stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
        return self.recv


def _cpulist(text):
    cpus = []
    for part in text.strip().split(','):
        if not part:
            continue
        lo, _, hi = part.partition('-')
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_local_cpus(cuda_index):
    """(numa_node, [cpu ids]) of the host cores next to a GPU, from sysfs through the GPU's PCI address (falls back
    to NVML's ideal-affinity mask).  (None, []) when the platform does not say."""
    import torch
    try:
        p = torch.cuda.get_device_properties(cuda_index)
        bdf = '%04x:%02x:%02x.0' % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        base = '/sys/bus/pci/devices/' + bdf
        with open(base + '/local_cpulist') as f:
            cpus = _cpulist(f.read())
        node = None
        try:
            with open(base + '/numa_node') as f:
                node = int(f.read())
        except (OSError, ValueError):
            pass
        if cpus:
            return node, cpus
    except Exception:
        pass
    try:
        import pynvml
        pynvml.nvmlInit()
        p = torch.cuda.get_device_properties(cuda_index)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(('%08x:%02x:%02x.0' % (p.pci_domain_id, p.pci_bus_id,
                                                                        p.pci_device_id)).encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1]
        return None, cpus
    except Exception:
        return None, []


def pin_rank_to_gpu_numa(local_rank, local_world):
    """Binds the calling process to host cores next to ITS GPU, so that the pinned staging buffers it allocates
    afterwards are placed on that NUMA node (first touch) and its copy-engine submissions do not cross the socket
    interconnect.  The ranks whose GPUs share a node split that node's cores evenly (by local rank).  Returns a dict
    describing the binding (`original` holds the previous affinity for callers that need all cores again), or None if
    the topology is unknown — in which case nothing is changed."""
    import torch
    try:
        original = sorted(os.sched_getaffinity(0))
    except AttributeError:
        return None
    n_gpu = torch.cuda.device_count()
    mine_node, mine = gpu_local_cpus(local_rank)
    mine = [c for c in mine if c in set(original)]
    if not mine:
        return None
    # which local ranks share these cores
    sharers = []
    for r in range(min(local_world, n_gpu)):
        _, cpus = gpu_local_cpus(r)
        if set(cpus) & set(mine):
            sharers.append(r)
    if local_rank not in sharers:
        sharers.append(local_rank)
    sharers.sort()
    k, n = sharers.index(local_rank), len(sharers)
    per = max(1, len(mine) // n)
    chosen = mine[k * per:(k + 1) * per] if k < n - 1 else mine[k * per:]
    if not chosen:
        chosen = mine
    try:
        os.sched_setaffinity(0, chosen)
    except OSError:
        return None
    return {'numa_node': mine_node, 'cpus': '%d-%d (%d cores)' % (chosen[0], chosen[-1], len(chosen)),
            'ranks_sharing_node': n, 'original': original}


def restore_affinity(binding):
    if binding and binding.get('original'):
        try:
            os.sched_setaffinity(0, binding['original'])
        except OSError:
            pass
