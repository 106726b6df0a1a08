#!/usr/bin/env python
"""bench.py — env-steps/s of the batched ATC approach-control step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]   # the CPU arm (oracle port, all host threads)
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...       # one rank per GPU (weak scaling)

A "step" is one environment step of every env in the batch.  Workload = BASELINE.json configs[2]:
16384 envs x 4 aircraft per GPU, LOWW sector with 9 entry points, all-pairs separation on, uniform random actions
re-sampled every 20 steps (the reference demo's cadence), auto-reset on.  Steps are executed as fused rollout
launches of --rollout steps each (state stays in registers); the last launch is shortened so that EXACTLY K steps
are timed.  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_ENVS, N_AIRCRAFT = 16384, 4
METRIC, UNIT = 'env-steps/sec', 'env-steps/s'
ACTION_REPEAT = 20


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


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index, period=0.005):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
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
        names = {
            getattr(nv, 'nvmlClocksThrottleReasonHwSlowdown', 0x8): 'hw_slowdown',
            getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40): 'hw_thermal_slowdown',
            getattr(nv, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20): 'sw_thermal_slowdown',
            getattr(nv, 'nvmlClocksThrottleReasonSwPowerCap', 0x4): 'sw_power_cap',
        }
        self.names = names
        while not self._stop_evt.is_set():
            try:
                mhz = int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((time.perf_counter(), mhz, r))
            except Exception:
                pass
            time.sleep(self.period)

    def mark_begin(self):
        """Start of the timed region: earlier samples (warm-up) are dropped by stop()."""
        self.t_begin = time.perf_counter()

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.ok:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        kept = [x for x in self.samples if x[0] >= getattr(self, 't_begin', 0.0)] or self.samples[-1:]
        s = sorted(x[1] for x in kept)
        for _, _, r in kept:
            for bit, name in self.names.items():
                if r & bit:
                    self.reasons.add(name)
        return {'sm_mhz': s[len(s) // 2] if s else None, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'samples': len(s)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_oracle_rate(n_envs, n_aircraft, target_seconds, steps_hint=None, threads=None):
    """Times the CPU oracle (kind='port': the C restatement of the reference step, OpenMP over envs) on a bounded
    sample of the bench workload.  Returns (env_steps_per_s, cores, sample description, seconds)."""
    import numpy as np
    from oracle import oracle as O
    O.set_num_threads(threads or host_threads())
    cores = O.num_threads()
    rng = np.random.RandomState(1234)
    ora = O.Oracle('LOWW', True, n_env=n_envs, n_ac=n_aircraft, seed=0)
    ora.reset()
    T0 = 20
    a0 = np.repeat(rng.uniform(-1, 1, (1, n_envs, n_aircraft, 3)).astype(np.float32), T0, 0)
    ora.rollout(a0)                                   # warm-up (page faults, thread pool)
    t = time.perf_counter()
    ora.rollout(a0)
    per_step = (time.perf_counter() - t) / T0
    chunk = 200
    acts = np.repeat(rng.uniform(-1, 1, (chunk // ACTION_REPEAT, n_envs, n_aircraft, 3)).astype(np.float32),
                     ACTION_REPEAT, 0)
    n_chunks = steps_hint // chunk if steps_hint else max(1, int(target_seconds / max(per_step * chunk, 1e-9)))
    T = n_chunks * chunk
    t = time.perf_counter()
    for _ in range(n_chunks):
        ora.rollout(acts)
    sec = time.perf_counter() - t
    sample = '%d envs x %d aircraft x %d steps of the bench workload, %d OpenMP threads' % (n_envs, n_aircraft, T, cores)
    return n_envs * T / sec, cores, sample, sec


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path, timed on this box's host cores.  The
    reference is pure Python and cannot travel to the GPU box (no /root/reference there), so this is the oracle port
    (oracle/atc_oracle.c — bit-identical to the live reference on every golden trace) with all host threads."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import numpy as np
    from oracle import oracle as O
    O.set_num_threads(host_threads())
    cores = O.num_threads()
    # bounded sample per step: n_s envs of the 16384, sized so K + W steps take about 20 s
    n_probe = 1024
    rate, _, _, _ = cpu_oracle_rate(n_probe, N_AIRCRAFT, 1.0)
    total_steps = args.steps + args.warmup
    n_s = int(max(64, min(N_ENVS, rate * 20.0 / max(total_steps, 1))))
    rng = np.random.RandomState(1234)
    ora = O.Oracle('LOWW', True, n_env=n_s, n_ac=N_AIRCRAFT, seed=0)
    ora.reset()

    def run(T):
        done = 0
        while done < T:
            c = min(256, T - done)
            acts = np.repeat(rng.uniform(-1, 1, ((c + ACTION_REPEAT - 1) // ACTION_REPEAT, n_s, N_AIRCRAFT, 3))
                             .astype(np.float32), ACTION_REPEAT, 0)[:c]
            t0 = time.perf_counter()
            ora.rollout(acts)
            run.sec += time.perf_counter() - t0
            done += c
    run.sec = 0.0
    run(args.warmup)
    run.sec = 0.0
    run(args.steps)
    sec = run.sec
    value = n_s * args.steps / sec
    sample = ('%d of the %d envs x %d aircraft per step, %d steps, oracle port (C, float64, OpenMP %d threads)'
              % (n_s, N_ENVS, N_AIRCRAFT, args.steps, cores))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * sec / args.steps * (N_ENVS / n_s), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args, extra={'sampled_envs_per_step': n_s}),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


def workload_config(args, extra=None):
    c = {'workload': '%d envs x %d aircraft per GPU, LOWW 12-polygon MVA map, 9 entry points, all-pairs 3nm/1000ft '
                     'separation, auto-reset, U(-1,1) actions re-sampled every %d steps (BASELINE.json configs[2])'
                     % (N_ENVS, N_AIRCRAFT, ACTION_REPEAT),
         'envs_per_gpu': N_ENVS, 'aircraft_per_env': N_AIRCRAFT, 'rollout_steps_per_launch': args.rollout,
         'step_outputs': 'obs + info[original_state] + reward + done + term' if args.raw_obs else
                         'obs + reward + done + term (no info[original_state])',
         'l2': 'per-launch working set (actions %.0f MB in + observations %.0f MB out) exceeds the 126 MB L2; no flush'
               % (args.rollout * N_ENVS * N_AIRCRAFT * 12 / 1e6,
                  args.rollout * N_ENVS * N_AIRCRAFT * 40 * (2 if args.raw_obs else 1) / 1e6)}
    if extra:
        c.update(extra)
    return c


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from atc_reinforcement_learning_b200 import BatchedAtcEnv, LOWW, SimParameters
    from atc_reinforcement_learning_b200.dist import ReturnGather, init_from_env

    rank, world, local_rank = init_from_env('nccl')
    if world != args.gpus and rank == 0:
        print('warning: --gpus %d but WORLD_SIZE %d' % (args.gpus, world), file=sys.stderr)
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    N, A, TR = N_ENVS, N_AIRCRAFT, args.rollout
    RAW = bool(args.raw_obs)
    env = BatchedAtcEnv(N, A, SimParameters(1), LOWW(random_entrypoints=True), device=dev, seed=0,
                        env_index_base=rank * N, return_raw_obs=RAW, grid_cell=args.grid_cell)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    acts = (torch.rand((TR + ACTION_REPEAT - 1) // ACTION_REPEAT, N, A, 3, device=dev, generator=g) * 2 - 1)
    acts = acts.repeat_interleave(ACTION_REPEAT, 0)[:TR].contiguous()
    out = env._alloc_io((TR,))
    gather = ReturnGather(N, dev, overlap=None if args.gather_overlap < 0 else bool(args.gather_overlap))
    stream = torch.cuda.current_stream(dev)

    def run_steps(n):
        launches = 0
        done = 0
        while done < n:
            c = min(TR, n - done)
            env.rollout(acts[:c], out={k: v[:c] for k, v in out.items()})
            gather.gather(env.last_ep_return)            # NCCL all_gather of the episode-return log, side stream
            done += c
            launches += 1
        return launches

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # every rank samples its own GPU's clocks; the sampler (NVML init takes milliseconds) is up before the barrier so
    # that all ranks enter the timed region together, and only samples taken inside the region are kept
    sampler = ClockSampler(local_rank)
    sampler.start()
    run_steps(args.warmup)
    gather.wait()
    barrier()
    sampler.mark_begin()
    l0 = env.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    launches = run_steps(args.steps)
    ev1.record(stream)
    gather.wait()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    gpu_launches = env.launch_count - l0
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    ms_ranks = [ms]
    if world > 1:
        allms = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allms, t)
        ms_ranks = [float(x.item()) for x in allms]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    clocks_ranks = [clocks]
    if world > 1:
        clocks_ranks = [None] * world
        dist.all_gather_object(clocks_ranks, clocks)
    value = N * world * args.steps / (ms_max * 1e-3)

    # ---- single-step-per-launch mode (the gym step() call), for context: eager launches and a CUDA graph of them
    step_ms = graph_ms = None
    if not args.skip_extras:
        ks = min(args.steps, 2048)
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
        te = min(args.e2e_rollout, TR)
        h_act, h_out = env.alloc_pinned_io(te)
        # what travels back per step: obs, reward, done, term.  info["original_state"] is still written by the kernel,
        # it stays in HBM (the reference hands it out by reference, too: atc_gym.py:192)
        h_out = {k: v for k, v in h_out.items() if k != 'raw_obs'}
        h_act.copy_(acts[:te].cpu())
        ke = max(te, (min(args.steps, args.e2e_steps) // te) * te)
        env.rollout_pinned(h_act, h_out)
        barrier()
        t0 = time.perf_counter()
        for _ in range(ke // te):
            env.rollout_pinned(h_act, h_out)
        torch.cuda.synchronize(dev)
        sec = time.perf_counter() - t0
        te_t = torch.tensor([sec], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te_t, op=dist.ReduceOp.MAX)
        e2e = {'value': N * world * ke / float(te_t.item()), 'unit': UNIT,
               'h2d_bytes_per_step': N * A * 3 * 4, 'd2h_bytes_per_step': N * A * 10 * 4 + N * (4 + 1 + 4),
               'steps': ke, 'rollout_steps_per_call': te,
               'api': 'BatchedAtcEnv.rollout_pinned -> atc_rollout_host (pinned H2D of the actions, kernel, D2H of obs + '
                      'reward + done + term, sync); info[original_state] stays on the device'}

    if rank != 0:
        return
    peak, peak_src = measured_peaks()
    # dominant kernel = atc_step_kernel<4,...> in rollout mode; the timed region contains nothing else on its stream
    avg_launch_s = ms * 1e-3 / launches
    t_launch = args.steps / launches                    # mean steps per launch (the last launch may be shorter)
    bytes_per_launch = bytes_rollout(A, t_launch, RAW) * N * t_launch
    achieved = bytes_per_launch / avg_launch_s / 1e9
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': None, 'peak_source': peak_src, 'kernel': 'atc_rollout_pipe_kernel<4,false,false,false,14> (one CTA per SM, MVA grid in shared memory; rollout, T=%d)' % TR,
                'algorithmic_bytes_per_env_step': bytes_rollout(A, t_launch, RAW), 'avg_launch_ms': avg_launch_s * 1e3}
    tp = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tp):
        try:
            with open(tp) as f:
                roofline['traffic'] = json.load(f).get('dram_bytes_per_launch')
        except Exception:
            pass
    cpu = None
    if not args.skip_extras:
        v, cores, sample, sec = cpu_oracle_rate(2048, A, args.cpu_seconds)
        cpu = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample, 'seconds': sec}
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_max / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic', 'config': workload_config(args),
        'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': int(gpu_launches), 'clocks': clocks,
        'single_step_launch': None if step_ms is None else {
            'ms_per_step': step_ms, 'value': N / (step_ms * 1e-3), 'unit': UNIT,
            'roofline_frac': bytes_single_step(A, RAW) * N / (step_ms * 1e-3) / 1e9 / peak,
            'cuda_graph_ms_per_step': graph_ms,
            'cuda_graph_value': None if graph_ms is None else N / (graph_ms * 1e-3),
            'cuda_graph_roofline_frac': None if graph_ms is None else
            bytes_single_step(A, RAW) * N / (graph_ms * 1e-3) / 1e9 / peak},
        'nccl_gathers': gather.calls, 'ms_per_rank': ms_ranks,
        'clocks_per_rank': clocks_ranks if world > 1 else None,
    }
    emit(line)
    if world > 1:
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
    ap.add_argument('--steps', type=int, default=65536)
    ap.add_argument('--warmup', type=int, default=1024)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--rollout', type=int, default=1024,
                    help='env steps fused per kernel launch (1024 = the n_steps of the reference\'s PPO2 runner, '
                         'learning/atc-gym-stable-baselines.py:109-122)')
    ap.add_argument('--e2e-rollout', type=int, default=128)
    ap.add_argument('--e2e-steps', type=int, default=2048)
    ap.add_argument('--cpu-seconds', type=float, default=10.0)
    ap.add_argument('--grid-cell', type=float, default=0.0625, help='MVA lookup grid cell size in nm')
    ap.add_argument('--raw-obs', type=int, default=1,
                    help='1: every step also writes info["original_state"] (the raw observation the reference '
                         'returns, atc_gym.py:192); 0: normalised observation only')
    ap.add_argument('--gather-overlap', type=int, default=-1,
                    help='episode-return gather on a side stream (1) or in order on the step stream (0); -1 = default')
    ap.add_argument('--skip-extras', action='store_true', help='only the device-resident timing (used under ncu)')
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
