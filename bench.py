#!/usr/bin/env python
"""bench.py — env-steps/s of the batched ATC approach-control step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C]          # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]      # the CPU arm (oracle port, all host threads)
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...          # one rank per GPU (weak scaling)

One bench STEP = one `env.rollout()` call = ONE kernel launch advancing every env of the batch by --rollout (1024)
env-steps — the unit a PPO2 runner of the reference collects per update (n_steps = 1024,
/root/reference/learning/atc-gym-stable-baselines.py:109-122).  `value` is env-steps/s = envs x K x rollout / time.
The K-step block is enqueued back to back `timed_blocks` times (as many as make the timed region >= ~150 ms; CUDA
events between the blocks, no host synchronisation inside), each block is timed on the device, the maximum over the
ranks is taken per block and the MEDIAN block is reported; warm-up = W steps.  Workloads (--config):
    16384x4       BASELINE.json configs[2] — the configuration `metric` is quoted on (default)
    4096x1        configs[1]
    16384x8_wind  configs[3]: 8 aircraft, 16 x 16 wind grid U(-30, 30) kt
All: LOWW 12-polygon sector, 9 entry points, all-pairs 3 nm / 1000 ft separation, auto-reset, U(-1,1) actions
re-sampled every 20 steps (the reference demo's cadence), every step writing obs + info["original_state"] + reward +
done + term.  After the timing the same action stream is replayed through a fresh env in the timed launch shape and a
sample of its rows is checked against the CPU oracle (`parity_in_run`).  One JSON line is printed by rank 0.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC, UNIT = 'env-steps/sec', 'env-steps/s'
ACTION_REPEAT = 20
CONFIGS = {
    '16384x4': dict(n_envs=16384, n_aircraft=4, wind=False, baseline='BASELINE.json configs[2]'),
    '4096x1': dict(n_envs=4096, n_aircraft=1, wind=False, baseline='BASELINE.json configs[1]'),
    '16384x8_wind': dict(n_envs=16384, n_aircraft=8, wind=True, baseline='BASELINE.json configs[3]'),
}
L2_MB = 126.0


def host_threads():
    """All host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which would hide them)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def bytes_rollout(A, T, raw_obs=False):
    """algorithmic bytes per env-step of the fused T-step rollout (float32 SoA accounting): actions in,
    observation / reward / done out every step, state once per launch; + 40 B per aircraft when the step also writes
    info["original_state"] (the raw observation the reference returns with every step, atc_gym.py:192)."""
    return 52.0 * A + 8.0 + (40.0 * A + 8.0) / T + (40.0 * A if raw_obs else 0.0)


def bytes_single_step(A, raw_obs=False):
    return 92.0 * A + 16.0 + (40.0 * A if raw_obs else 0.0)


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def wind_grid():
    """configs[3]: 16 x 16 nodes, U(-30, 30) kt per component, fixed seed."""
    import numpy as np
    return np.random.RandomState(4).uniform(-30, 30, (16, 16, 2)).astype(np.float32)


def workload_config(args):
    """Pure description of the workload — identical for both arms (--impl b200 / reference)."""
    c = CONFIGS[args.config]
    N, A, T = c['n_envs'], c['n_aircraft'], args.rollout
    out_mb = T * N * (A * 40 * (2 if args.raw_obs else 1) + 9) / 1e6
    in_mb = T * N * A * 12 / 1e6
    return {
        'workload': '%d envs x %d aircraft per GPU, LOWW 12-polygon MVA map, 9 entry points, all-pairs 3nm/1000ft '
                    'separation%s, auto-reset, U(-1,1) actions re-sampled every %d steps (%s)'
                    % (N, A, ', 16x16 wind grid U(-30,30) kt' if c['wind'] else '', ACTION_REPEAT, c['baseline']),
        'name': args.config, 'envs_per_gpu': N, 'aircraft_per_env': A,
        'step_definition': 'one bench step = one rollout() call = one kernel launch of %d env-steps of every env' % T,
        'env_steps_per_step': T,
        'step_outputs': 'obs + info[original_state] + reward + done + term' if args.raw_obs else
                        'obs + reward + done + term (no info[original_state])',
        'l2': 'inputs larger than L2: every step streams %.0f MB of actions in and %.0f MB of outputs out (L2 = %.0f MB), '
              'each byte touched once; no flush' % (in_mb, out_mb, L2_MB),
    }


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index, period=0.004):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        self.t_begin, self.t_end = 0.0, float('inf')
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        self.names = {
            getattr(nv, 'nvmlClocksThrottleReasonHwSlowdown', 0x8): 'hw_slowdown',
            getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40): 'hw_thermal_slowdown',
            getattr(nv, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20): 'sw_thermal_slowdown',
            getattr(nv, 'nvmlClocksThrottleReasonSwPowerCap', 0x4): 'sw_power_cap',
        }
        while not self._stop_evt.is_set():
            try:
                mhz = int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                try:
                    mem = int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_MEM))
                    pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                except Exception:
                    mem, pw = None, None
                self.samples.append((time.perf_counter(), mhz, r, mem, pw))
            except Exception:
                pass
            time.sleep(self.period)

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def mark_end(self):
        self.t_end = time.perf_counter()

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.ok:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        kept = [x for x in self.samples if self.t_begin <= x[0] <= self.t_end] or self.samples[-1:]
        s = sorted(x[1] for x in kept)
        for x in kept:
            for bit, name in self.names.items():
                if x[2] & bit:
                    self.reasons.add(name)
        mem = sorted(x[3] for x in kept if x[3] is not None)
        pw = [x[4] for x in kept if x[4] is not None]
        return {'sm_mhz': s[len(s) // 2] if s else None, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'samples': len(s), 'mem_mhz': mem[len(mem) // 2] if mem else None,
                'power_w_first_last': [round(pw[0]), round(pw[-1])] if pw else None}


# ------------------------------------------------------------------------------------------------ CPU arm
def make_oracle(args, n_env, env_index_base=0, seed=0):
    from oracle import oracle as O
    c = CONFIGS[args.config]
    return O.Oracle('LOWW', True, n_env=n_env, n_ac=c['n_aircraft'], seed=seed, env_index_base=env_index_base,
                    wind=wind_grid() if c['wind'] else None)


def host_actions(rng, T, n_env, A):
    import numpy as np
    return np.repeat(rng.uniform(-1, 1, ((T + ACTION_REPEAT - 1) // ACTION_REPEAT, n_env, A, 3)).astype(np.float32),
                     ACTION_REPEAT, 0)[:T]


def cpu_oracle_rate(args, n_envs, target_seconds):
    """Times the CPU oracle (kind='port': the C restatement of the reference step, OpenMP over envs) on a bounded
    sample of the bench workload.  Returns (env_steps_per_s, cores, sample description, seconds)."""
    import numpy as np
    from oracle import oracle as O
    O.set_num_threads(host_threads())
    cores = O.num_threads()
    A = CONFIGS[args.config]['n_aircraft']
    rng = np.random.RandomState(1234)
    ora = make_oracle(args, n_envs)
    ora.reset()
    a0 = host_actions(rng, 20, n_envs, A)
    ora.rollout(a0)                                   # warm-up (page faults, thread pool)
    t = time.perf_counter()
    ora.rollout(a0)
    per_step = (time.perf_counter() - t) / 20
    chunk = 200
    acts = host_actions(rng, chunk, n_envs, A)
    n_chunks = max(1, int(target_seconds / max(per_step * chunk, 1e-9)))
    T = n_chunks * chunk
    t = time.perf_counter()
    for _ in range(n_chunks):
        ora.rollout(acts)
    sec = time.perf_counter() - t
    sample = '%d envs x %d aircraft x %d env-steps of the bench workload, %d OpenMP threads' % (n_envs, A, T, cores)
    return n_envs * T / sec, cores, sample, sec


def reference_python_rate(seconds):
    """The reference's OWN Python step() on this host: oracle/ref_python_bench.py in a subprocess —
    unmodified envs/atc from baseline/_ref under the import stand-ins; one core and all cores."""
    try:
        env = dict(os.environ)
        env.pop('OMP_NUM_THREADS', None)
        out = subprocess.run([sys.executable, os.path.join(ROOT, 'oracle', 'ref_python_bench.py'), '--seconds',
                              str(seconds)], capture_output=True, text=True, timeout=300, env=env)
        line = [ln for ln in out.stdout.splitlines() if ln.startswith('{')]
        if not line:
            return {'unavailable': (out.stderr or 'no output').strip().splitlines()[-1][:200]}
        return json.loads(line[-1])
    except Exception as e:                                   # pragma: no cover - reported, not fatal
        return {'unavailable': repr(e)[:200]}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path, timed on this box's host cores.  The
    reference is pure Python (nothing to compile into oracle/_ref), so this is the oracle port (oracle/atc_oracle.c —
    bit-identical to the live reference on every golden trace) with all host threads, in the SAME unit as the GPU
    arm: one step = one rollout of --rollout env-steps of the batch.  If the full batch would take too long the step
    runs on a contiguous sample of the envs (stated in cpu_baseline.sample; value is env-steps/s either way)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import numpy as np
    from oracle import oracle as O
    O.set_num_threads(host_threads())
    cores = O.num_threads()
    c = CONFIGS[args.config]
    N, A, TR = c['n_envs'], c['n_aircraft'], args.rollout
    rate, _, _, _ = cpu_oracle_rate(args, 1024, 1.0)
    total_steps = args.steps + args.warmup
    budget = args.reference_seconds
    n_s = int(min(N, max(64, rate * budget / (total_steps * TR))))
    if n_s >= N * 0.9:
        n_s = N
    rng = np.random.RandomState(1234)
    ora = make_oracle(args, n_s)
    ora.reset()
    acts = host_actions(rng, TR, n_s, A)                      # like the GPU arm: one action buffer, re-used every step
    for _ in range(args.warmup):
        ora.rollout(acts)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ora.rollout(acts)
    sec = time.perf_counter() - t0
    value = n_s * args.steps * TR / sec
    sample = ('%d of the %d envs x %d aircraft, %d steps of %d env-steps each, oracle port (C, float64, OpenMP %d '
              'threads)' % (n_s, N, A, args.steps, TR, cores))
    cpu = {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample, 'seconds': sec}
    if not args.skip_extras:
        cpu['reference_python'] = reference_python_rate(args.refpy_seconds)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * sec / args.steps * (N / n_s), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args), 'cpu_baseline': cpu,
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0, 'timing': {'timed_region_ms': sec * 1e3, 'sampled_envs_per_step': n_s},
    }
    emit(line)


# ------------------------------------------------------------------------------------------------ GPU arm
def parity_in_run(args, env_factory, acts, out, rank, timed_launch):
    """the action stream that was timed, replayed from reset through a FRESH env in the timed launch
    shape (same batch, same launch length, same kernel layout), and a sample of the rows checked against the CPU
    oracle: 4 blocks of 128 consecutive envs spread over the batch, every step, every output."""
    import numpy as np
    import torch
    c = CONFIGS[args.config]
    N, A = c['n_envs'], c['n_aircraft']
    env = env_factory()
    env.rollout(acts, out=out)
    torch.cuda.synchronize()
    ll = env.last_launch
    same = all(ll[k] == timed_launch[k] for k in ('kernel', 'n_steps', 'grid', 'block', 'pairs_per_cta', 'raw_obs'))
    blk = min(128, N)
    bases = sorted(set(int(b) for b in np.linspace(0, N - blk, 4)))
    res = {'max_obs_err': 0.0, 'max_raw_err': 0.0, 'max_rew_err': 0.0, 'max_err_over_tolerance': 0.0, 'flags_equal': True, 'within_tolerance': True,
           'rows_checked': 0, 'episodes_finished': 0, 'same_launch_shape_as_timed': bool(same),
           'tolerance': '1e-5 + 1e-5 |ref| on obs / original_state / reward; done and term bit-exact'}
    for b0 in bases:
        ora = make_oracle(args, blk, env_index_base=rank * N + b0, seed=0)
        ora.reset()
        a = acts[:, b0:b0 + blk].cpu().numpy()
        ref = ora.rollout(a, raw=True)
        o_obs, o_raw, o_rew, o_done, o_term = ref
        g_done = out['done'][:, b0:b0 + blk].cpu().numpy()
        g_term = out['term'][:, b0:b0 + blk].cpu().numpy()
        res['flags_equal'] &= bool((g_done == o_done).all() and (g_term == o_term).all())
        pairs = [('max_obs_err', out['obs'], o_obs), ('max_rew_err', out['reward'], o_rew)]
        if out.get('raw_obs') is not None:
            pairs.append(('max_raw_err', out['raw_obs'], o_raw))
        for key, g, r in pairs:
            gv = g[:, b0:b0 + blk].cpu().numpy().astype(np.float64)
            err = np.abs(gv - r)
            res[key] = max(res[key], float(err.max()))
            res['max_err_over_tolerance'] = max(res['max_err_over_tolerance'],
                                                float((err / (1e-5 + 1e-5 * np.abs(r))).max()))
            res['within_tolerance'] &= bool((err <= 1e-5 + 1e-5 * np.abs(r)).all())
        res['rows_checked'] += int(o_done.size)
        res['episodes_finished'] += int(o_done.sum())
    env.close()
    return res


def quick_config(args, name, dev, peak, launches=10):
    """Short leg of another BASELINE config: `launches` back-to-back rollout launches of --rollout env-steps timed on
    the device after 3 warm-up launches, the roofline of what actually launched, and the parity check of that launch
    shape against the oracle."""
    import copy
    import torch
    from atc_reinforcement_learning_b200 import BatchedAtcEnv, LOWW, SimParameters
    a2 = copy.copy(args)
    a2.config = name
    c = CONFIGS[name]
    N, A, TR, RAW = c['n_envs'], c['n_aircraft'], args.rollout, bool(args.raw_obs)
    wind = wind_grid() if c['wind'] else None

    def make():
        return BatchedAtcEnv(N, A, SimParameters(1), LOWW(random_entrypoints=True), device=dev, seed=0,
                             return_raw_obs=RAW, grid_cell=args.grid_cell, wind=wind)
    env = make()
    g = torch.Generator(device=dev).manual_seed(1234)
    acts = (torch.rand((TR + ACTION_REPEAT - 1) // ACTION_REPEAT, N, A, 3, device=dev, generator=g) * 2 - 1)
    acts = acts.repeat_interleave(ACTION_REPEAT, 0)[:TR].contiguous()
    out = env._alloc_io((TR,))
    for _ in range(3):
        env.rollout(acts, out=out)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(launches):
        env.rollout(acts, out=out)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / launches
    ll = env.last_launch
    bpe = bytes_rollout(A, ll['n_steps'], RAW)
    achieved = bpe * N * ll['n_steps'] / (ms * 1e-3) / 1e9
    res = {'workload': workload_config(a2)['workload'], 'value': N * TR / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms,
           'steps': launches, 'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                                           'frac': achieved / peak, 'kernel': '%s (%s)' % (ll['name'], ll['layout']),
                                           'algorithmic_bytes_per_env_step': bpe}}
    if not args.skip_parity:
        res['parity_in_run'] = parity_in_run(a2, make, acts, out, 0, ll)
    env.close()
    del env, acts, out
    torch.cuda.empty_cache()
    return res


def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from atc_reinforcement_learning_b200 import BatchedAtcEnv, LOWW, SimParameters
    from atc_reinforcement_learning_b200.dist import ReturnGather, init_from_env, pin_rank_to_gpu_numa, restore_affinity

    rank, world, local_rank = init_from_env('nccl')
    if world != args.gpus and rank == 0:
        print('warning: --gpus %d but WORLD_SIZE %d' % (args.gpus, world), file=sys.stderr)
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    affinity = pin_rank_to_gpu_numa(local_rank, world) if not args.no_pin else None
    c = CONFIGS[args.config]
    N, A, TR, K = c['n_envs'], c['n_aircraft'], args.rollout, args.steps
    RAW = bool(args.raw_obs)
    wind = wind_grid() if c['wind'] else None

    def make(seed=0):
        return BatchedAtcEnv(N, A, SimParameters(1), LOWW(random_entrypoints=True), device=dev, seed=seed,
                             env_index_base=rank * N, return_raw_obs=RAW, grid_cell=args.grid_cell, wind=wind)

    env = make()
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    acts = (torch.rand((TR + ACTION_REPEAT - 1) // ACTION_REPEAT, N, A, 3, device=dev, generator=g) * 2 - 1)
    acts = acts.repeat_interleave(ACTION_REPEAT, 0)[:TR].contiguous()
    out = env._alloc_io((TR,))
    gather = ReturnGather(N, dev, overlap=None if args.gather_overlap < 0 else bool(args.gather_overlap))
    stream = torch.cuda.current_stream(dev)

    def run_steps(n):
        for _ in range(n):
            env.rollout(acts, out=out)
            gather.gather(env.last_ep_return)        # NCCL all_gather of the episode-return log, side stream
        return n

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)
    sampler.start()
    # ---- warm-up: W steps, timed only to size the timed region
    ev_w0, ev_w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    run_steps(1)
    ev_w0.record(stream)
    run_steps(max(args.warmup - 1, 1))
    ev_w1.record(stream)
    gather.wait()
    barrier()
    est_ms = ev_w0.elapsed_time(ev_w1) / max(args.warmup - 1, 1)
    blocks = args.blocks if args.blocks > 0 else int(min(64, max(3, math.ceil(args.min_region_ms / max(est_ms * K, 1e-3)))))
    if world > 1:                                   # every rank must run the same number of blocks
        tb = torch.tensor([blocks], device=dev, dtype=torch.int64)
        dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        blocks = int(tb.item())
    # ---- timed region: `blocks` x K steps back to back, an event between blocks, nothing else on the stream
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(blocks + 1)]
    barrier()
    sampler.mark_begin()
    l0 = env.launch_count
    evs[0].record(stream)
    for b in range(blocks):
        run_steps(K)
        evs[b + 1].record(stream)
    gather.wait()
    barrier()
    sampler.mark_end()
    clocks = sampler.stop()
    gpu_launches_total = env.launch_count - l0
    timed_launch = env.last_launch
    block_ms = torch.tensor([evs[b].elapsed_time(evs[b + 1]) for b in range(blocks)], device=dev, dtype=torch.float64)
    per_rank = [block_ms.clone()]
    if world > 1:
        per_rank = [torch.zeros_like(block_ms) for _ in range(world)]
        dist.all_gather(per_rank, block_ms)
    block_max = torch.stack(per_rank).max(0).values                 # per block: max over ranks
    ms_block = float(block_max.median().item())
    ms_ranks = [float(p.median().item()) for p in per_rank]
    clocks_ranks = [clocks]
    if world > 1:
        clocks_ranks = [None] * world
        dist.all_gather_object(clocks_ranks, clocks)
    value = N * world * K * TR / (ms_block * 1e-3)

    # ---- single-step-per-launch mode (the gym step() call), for context: eager launches and a CUDA graph of them
    step_ms = graph_ms = None
    step_launch = None
    if not args.skip_extras:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ks = 2048
        a1 = acts[0]
        o1 = {k: v[0] for k, v in out.items()}
        for _ in range(20):
            env.step(a1, out=o1)
        torch.cuda.synchronize(dev)
        ev0.record(stream)
        for i in range(ks):
            env.step(acts[i % TR], out=o1)
        ev1.record(stream)
        torch.cuda.synchronize(dev)
        step_ms = ev0.elapsed_time(ev1) / ks
        step_launch = env.last_launch
        GT = min(TR, 128)                              # steps captured into one CUDA graph
        try:                                           # the library never allocates -> step() is graph-capturable
            gs = torch.cuda.Stream(dev)
            gs.wait_stream(stream)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.stream(gs):
                env.step(a1, out=o1)
                with torch.cuda.graph(graph, stream=gs):
                    for i in range(GT):
                        env.step(acts[i], out={k: v[i] for k, v in out.items()})
            stream.wait_stream(gs)
            torch.cuda.synchronize(dev)
            reps = max(1, ks // GT)
            graph.replay()
            torch.cuda.synchronize(dev)
            ev0.record(stream)
            for _ in range(reps):
                graph.replay()
            ev1.record(stream)
            torch.cuda.synchronize(dev)
            graph_ms = ev0.elapsed_time(ev1) / (reps * GT)
        except Exception as e:                         # pragma: no cover - reported, not fatal
            print('cuda graph mode failed: %r' % (e,), file=sys.stderr)

    # ---- end to end through the host-buffer C-ABI entry point (pinned host memory both ways)
    e2e = None
    if not args.skip_extras:
        te = TR if args.e2e_rollout <= 0 else min(args.e2e_rollout, TR)   # default: the same step as the timed leg
        h_act, h_all = env.alloc_pinned_io(te)
        h_act.copy_(acts[:te].cpu())

        def e2e_leg(h_out, calls):
            env.rollout_pinned(h_act, h_out)
            barrier()
            t0 = time.perf_counter()
            for _ in range(calls):
                env.rollout_pinned(h_act, h_out)      # returns after the D2H copies completed
            torch.cuda.synchronize(dev)
            sec = time.perf_counter() - t0
            t = torch.tensor([sec], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return N * world * calls * te / float(t.item())

        calls = max(1, args.e2e_steps)
        v_all = e2e_leg(h_all, calls)                  # everything step() returns, info["original_state"] included
        d2h = N * A * 10 * 4 * (2 if RAW else 1) + N * (4 + 1 + 4)
        e2e = {'value': v_all, 'unit': UNIT, 'h2d_bytes_per_step': N * A * 3 * 4 * te, 'd2h_bytes_per_step': d2h * te,
               'h2d_bytes_per_env_step_batch': N * A * 3 * 4, 'd2h_bytes_per_env_step_batch': d2h,
               'env_steps_per_call': te, 'calls': calls,
               'api': 'BatchedAtcEnv.rollout_pinned -> atc_rollout_host: pinned H2D of the actions, kernel, D2H of '
                      'EVERY output the device-timed step writes (%s), stream sync; bytes are per call of %d env-steps'
                      % (workload_config(args)['step_outputs'], te)}
        if RAW:                                        # for context: the trainer-facing subset (no original_state)
            h_sub = {k: v for k, v in h_all.items() if k != 'raw_obs'}
            e2e['without_original_state'] = {'value': e2e_leg(h_sub, calls), 'unit': UNIT,
                                             'd2h_bytes_per_step': (N * A * 40 + N * 9) * te}

    # ---- parity of the timed launch shape against the oracle, same action stream (every rank its own shard)
    parity = None
    if not args.skip_parity:
        parity = parity_in_run(args, make, acts, out, rank, timed_launch)
        if world > 1:
            allp = [None] * world
            dist.all_gather_object(allp, parity)
            if rank == 0:
                for p in allp[1:]:
                    for k in ('max_obs_err', 'max_raw_err', 'max_rew_err', 'max_err_over_tolerance'):
                        parity[k] = max(parity[k], p[k])
                    for k in ('flags_equal', 'within_tolerance', 'same_launch_shape_as_timed'):
                        parity[k] = bool(parity[k] and p[k])
                    for k in ('rows_checked', 'episodes_finished'):
                        parity[k] += p[k]

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peaks()
    # ---- the other BASELINE configs, short legs (1-GPU runs only): device-timed value, roofline, parity in run
    others = None
    if world == 1 and not args.skip_extras and not args.skip_other_configs:
        del out, acts, env
        torch.cuda.empty_cache()
        others = {}
        for name in sorted(CONFIGS):
            if name != args.config:
                others[name] = quick_config(args, name, dev, peak)
    # dominant kernel: the rollout launch (one per bench step); the region also holds, per step, the 64 KB snapshot copy
    # of the return log (a torch kernel, < 3 us) — included in the time, not in the bytes
    avg_launch_s = ms_block * 1e-3 / K
    n_launch_steps = timed_launch['n_steps']
    bpe = bytes_rollout(A, n_launch_steps, RAW)
    bytes_per_launch = bpe * N * n_launch_steps
    achieved = bytes_per_launch / avg_launch_s / 1e9
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': None, 'peak_source': peak_src,
                'kernel': '%s (%s; %d env-steps per launch, grid %d x %d threads, %d B dynamic smem)'
                          % (timed_launch['name'], timed_launch['layout'], n_launch_steps, timed_launch['grid'],
                             timed_launch['block'], timed_launch['dyn_smem_bytes']),
                'launch': {k: timed_launch[k] for k in ('kernel', 'n_steps', 'grid', 'block', 'pairs_per_cta',
                                                         'lanes_per_env', 'wind', 'raw_obs', 'dyn_smem_bytes')},
                'algorithmic_bytes_per_env_step': bpe, 'algorithmic_bytes_per_launch': bytes_per_launch,
                'avg_launch_ms': avg_launch_s * 1e3}
    tp = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tp):                              # ncu --set full capture of ONE launch; only if it is this shape
        try:
            with open(tp) as f:
                tr = json.load(f)
            shape = tr.get('launch', {})
            if all(shape.get(k) == v for k, v in (('n_envs', N), ('n_aircraft', A), ('n_steps', n_launch_steps),
                                                    ('raw_obs', int(RAW)), ('kernel', timed_launch['kernel']))):
                roofline['traffic'] = tr.get('dram_bytes_per_launch')
                roofline['traffic_source'] = tr.get('source')
        except Exception:
            pass
    cpu = None
    if not args.skip_extras:
        restore_affinity(affinity)                      # the CPU baselines use every host core again
        v, cores, sample, sec = cpu_oracle_rate(args, 2048, args.cpu_seconds)
        cpu = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample, 'seconds': sec,
               'reference_python': reference_python_rate(args.refpy_seconds)}
    peak_step = bytes_single_step(A, RAW) * N / 1e9 / peak
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': args.warmup,
        'ms_per_step': ms_block / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic', 'config': workload_config(args),
        'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': int(gpu_launches_total // blocks),
        'clocks': clocks, 'parity_in_run': parity,
        'timing': {'timed_blocks': blocks, 'block_ms': [float(x) for x in block_max.tolist()],
                   'timed_region_ms': float(block_max.sum().item()), 'reported': 'median block, max over ranks',
                   'gpu_launches_in_region': int(gpu_launches_total), 'env_steps_per_block': K * TR,
                   'cpu_affinity': None if not affinity else {k: v for k, v in affinity.items() if k != 'original'}},
        'single_step_launch': None if step_ms is None else {
            'kernel': step_launch['name'], 'ms_per_step': step_ms, 'value': N / (step_ms * 1e-3), 'unit': UNIT,
            'roofline_frac': peak_step / (step_ms * 1e-3),
            'cuda_graph_ms_per_step': graph_ms,
            'cuda_graph_value': None if graph_ms is None else N / (graph_ms * 1e-3),
            'cuda_graph_roofline_frac': None if graph_ms is None else peak_step / (graph_ms * 1e-3)},
        'other_configs': others, 'nccl_gathers': gather.calls, 'ms_per_rank': ms_ranks,
        'clocks_per_rank': clocks_ranks if world > 1 else None,
    }
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_JSON_OUT = None


def emit(line):
    """The one JSON line goes to the real stdout; everything else libraries print to fd 1 (e.g. NCCL's version banner)
    has been rerouted to stderr by main()."""
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + '\n')
    out.flush()


def main():
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20, help='bench steps per timed block (one step = one rollout launch)')
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='16384x4', choices=sorted(CONFIGS))
    ap.add_argument('--rollout', type=int, default=1024,
                    help='env steps fused per kernel launch = per bench step (1024 = the n_steps of the reference\'s '
                         'PPO2 runner, learning/atc-gym-stable-baselines.py:109-122)')
    ap.add_argument('--blocks', type=int, default=0, help='timed K-step blocks (0 = as many as fill --min-region-ms, >= 3)')
    ap.add_argument('--min-region-ms', type=float, default=150.0)
    ap.add_argument('--e2e-rollout', type=int, default=0, help='env-steps per end-to-end call (0 = --rollout, the bench step)')
    ap.add_argument('--e2e-steps', type=int, default=4, help='end-to-end calls timed')
    ap.add_argument('--cpu-seconds', type=float, default=8.0)
    ap.add_argument('--refpy-seconds', type=float, default=4.0)
    ap.add_argument('--reference-seconds', type=float, default=60.0,
                    help='--impl reference: budget for the K + W steps; the env sample is sized to it')
    ap.add_argument('--grid-cell', type=float, default=0.0625, help='MVA lookup grid cell size in nm')
    ap.add_argument('--raw-obs', type=int, default=1,
                    help='1: every step also writes info["original_state"] (the raw observation the reference '
                         'returns, atc_gym.py:192); 0: normalised observation only')
    ap.add_argument('--gather-overlap', type=int, default=-1,
                    help='episode-return gather on a side stream (1) or in order on the step stream (0); -1 = default')
    ap.add_argument('--no-pin', action='store_true', help='do not bind the rank to the cores of its GPU\'s NUMA node')
    ap.add_argument('--skip-extras', action='store_true', help='only the device-resident timing (used under ncu)')
    ap.add_argument('--skip-parity', action='store_true')
    ap.add_argument('--skip-other-configs', action='store_true',
                    help='do not append the short legs of the other BASELINE configs (4096x1, 16384x8_wind)')
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
