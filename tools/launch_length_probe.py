"""Step time against the launch length T (16384 envs x 4 aircraft, every output written), for the default layout
selection and with one layout forced — the per-launch fixed cost (launch, staging the MVA grid per SM, slowest-pair tail)
is what separates short launches from long ones.  Writes JSON to stdout.  Usage: python tools/launch_length_probe.py"""
import json, os, subprocess, sys
if len(sys.argv) > 1 and sys.argv[1] == 'child':
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from atc_reinforcement_learning_b200 import BatchedAtcEnv, LOWW, SimParameters
    dev = torch.device('cuda', 0)
    N = 16384
    env = BatchedAtcEnv(N, 4, SimParameters(1), LOWW(random_entrypoints=True), device=dev, seed=0, return_raw_obs=True)
    res = {}
    for T in (4, 8, 16, 20, 32, 64, 128, 256, 512, 1024):
        acts = (torch.rand((T + 19) // 20, N, 4, 3, device=dev) * 2 - 1).repeat_interleave(20, 0)[:T].contiguous()
        out = env._alloc_io((T,))
        for _ in range(3):
            env.rollout(acts, out=out)
        torch.cuda.synchronize()
        reps = max(8, 8192 // T)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            env.rollout(acts, out=out)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        res[T] = {'us_per_launch': us, 'us_per_step': us / T, 'kernel': env.last_launch['name']}
        del acts, out
    print(json.dumps(res))
    sys.exit(0)
out = {}
for name, envs in (('default', {}), ('one_cta_per_sm_forced', {'ATC_B200_BIG_MIN_STEPS': '1'}),
                   ('pair_per_cta_forced', {'ATC_B200_NO_SMEM_GRID': '1'})):
    e = dict(os.environ); e.update(envs)
    r = subprocess.run([sys.executable, os.path.abspath(__file__), 'child'], capture_output=True, text=True, env=e)
    line = [l for l in r.stdout.splitlines() if l.startswith('{')]
    out[name] = json.loads(line[-1]) if line else {'error': r.stderr[-400:]}
print(json.dumps(out, indent=1))
