import os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
N, A, T = 16384, 4, 128
dbg = torch.zeros(2048 * 16, dtype=torch.int64, device='cuda')
os.environ['ATC_B200_DBG_PTR'] = str(dbg.data_ptr())
from atc_reinforcement_learning_b200 import BatchedAtcEnv, LOWW, SimParameters
env = BatchedAtcEnv(N, A, SimParameters(1), LOWW(random_entrypoints=True), seed=0, return_raw_obs=False)
g = torch.Generator(device='cuda').manual_seed(1234)
acts = (torch.rand(7, N, A, 3, device='cuda', generator=g) * 2 - 1).repeat_interleave(20, 0)[:T].contiguous()
out = env._alloc_io((T,))
for i in range(12):
    env.rollout(acts, out=out)
torch.cuda.synchronize()
ts = dbg.cpu().numpy().reshape(2048, 16).astype(np.float64)
t0 = ts[:, 0].min()
ts = (ts - t0) / 1e3     # us
print('start spread (us): min %.1f max %.1f' % (ts[:, 0].min(), ts[:, 0].max()))
seg = np.diff(ts[:, :8], axis=1)     # time per 16 steps, segments 0..6 (steps 0-112)
print('mean us per 16 steps by segment:', np.round(seg.mean(0), 1))
print('p50/p95/max of step-112 timestamp (us):', np.percentile(ts[:, 7], [50, 95, 100]).round(1))
print('per-CTA total(0->112) us: mean %.1f std %.1f min %.1f max %.1f' % ((ts[:, 7]-ts[:, 0]).mean(), (ts[:, 7]-ts[:, 0]).std(), (ts[:, 7]-ts[:, 0]).min(), (ts[:, 7]-ts[:, 0]).max()))
raw = dbg.cpu().numpy().reshape(2048, 16)
smid = (raw[:, 15] >> 32).astype(int); wid = (raw[:, 15] & 0xFFFFFFFF).astype(int)
tot = ts[:, 7] - ts[:, 0]
import collections
per_sm = collections.Counter(smid.tolist())
print('CTAs per SM histogram:', collections.Counter(per_sm.values()))
key = smid * 4 + (wid % 4)
cnt = collections.Counter(key.tolist())
movers_on_smsp = np.array([cnt[k] for k in key])
for m in sorted(set(movers_on_smsp)):
    sel = movers_on_smsp == m
    print('movers on the same SMSP = %d: %4d CTAs, mean total %.1f us (std %.1f)' % (m, sel.sum(), tot[sel].mean(), tot[sel].std()))
print('mover warpid%4 histogram:', collections.Counter((wid % 4).tolist()))
print('mover warpid histogram (first 16):', sorted(collections.Counter(wid.tolist()).items())[:32])
nsm = np.array([per_sm[s] for s in smid])
for m in sorted(set(nsm)):
    print('CTAs on SM = %d: mean total %.1f' % (m, tot[nsm == m].mean()))
