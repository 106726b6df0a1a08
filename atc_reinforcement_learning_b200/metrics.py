"""Episode-metric log of a batched env — the scalars the reference's training callback writes to
TensorBoard (/root/reference/learning/atc-gym-stable-baselines.py:31-49): `simulation/mean_actions`,
`simulation/winning_ratio`, `simulation/fps`, `simulation/mean_episode_length`, plus the mean episode return the
Monitor wrapper records (:73).  All of them are reductions over per-env DEVICE counters the step kernels maintain
(atc_gym.py:29-41, 194-197, 354-363): one small reduction per log call, summed over ranks with ONE all_reduce of six
float64 values (NCCL on cuda tensors, gloo on cpu tensors); rank 0 appends a JSON line per call.  TensorBoard is not
assumed: the file is `{"step": ..., "simulation/...": ...}` per line, which any scalar logger can ingest.

Also here: `write_evaluation_csv`, the reference's evaluation trace (atc-gym-stable-baselines.py:95-106): x, y, h, phi
of info["original_state"] per step, "%.2f, %.2f, %.0f, %.1f"."""
import json
import time

import torch
import torch.distributed as dist

TAGS = ('simulation/mean_actions', 'simulation/winning_ratio', 'simulation/mean_episode_length',
        'simulation/mean_episode_return', 'simulation/fps', 'simulation/episodes')


def local_sums(timesteps, actions_taken, win_ring, last_ep_len, last_ep_return):
    """The six additive statistics of one shard, float64 [6] on the tensors' device:
    sum(actions_per_timestep), sum(winning_ratio), sum(last_ep_len | finished), sum(last_ep_return | finished),
    number of envs that have finished an episode, number of envs.  (atc_gym.py:197: actions_taken / timesteps;
    :359-363: wins among the last 9 episodes x 0.1; envs that never finished an episode do not vote on length/return.)"""
    f64 = torch.float64
    n = timesteps.numel()
    if actions_taken is not None:
        apt = actions_taken.to(f64) / timesteps.clamp(min=1).to(f64)
    else:
        apt = torch.zeros(n, dtype=f64, device=timesteps.device)
    w = win_ring.to(torch.int64) & 0x1FF
    wins = torch.zeros_like(w)
    for k in range(9):
        wins += (w >> k) & 1
    fin = (last_ep_len > 0).to(f64)                    # an env votes on length / return once it has finished an episode
    return torch.stack([apt.sum(), (wins.to(f64) * 0.1).sum(), (last_ep_len.to(f64) * fin).sum(),
                        (last_ep_return.to(f64) * fin).sum(), fin.sum(),
                        torch.tensor(float(n), dtype=f64, device=timesteps.device)])


def reduce_scalars(sums, steps_done, seconds):
    """[6] local sums -> the scalar dict, summed over the process group if there is one."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        sums = sums.clone()
        dist.all_reduce(sums)
    s = [float(v) for v in sums.cpu()]
    n_env, n_fin = max(s[5], 1.0), s[4]
    return {
        'simulation/mean_actions': s[0] / n_env,
        'simulation/winning_ratio': s[1] / n_env,
        'simulation/mean_episode_length': s[2] / n_fin if n_fin > 0 else float('nan'),
        'simulation/mean_episode_return': s[3] / n_fin if n_fin > 0 else float('nan'),
        'simulation/fps': s[5] * steps_done / seconds if seconds > 0 else float('nan'),
        'simulation/episodes': n_fin,
    }


class EpisodeMetricsLog(object):
    """log(env, num_timesteps) -> dict of the TAGS; rank 0 appends it as a JSON line to `path` (if given)."""

    def __init__(self, path=None):
        self.path = path
        self.rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        self._t_last = time.perf_counter()
        self._steps_last = 0

    def log(self, env, num_timesteps):
        """`env`: anything with the BatchedAtcEnv counters (timesteps, actions_taken (or None), win_ring, last_ep_len,
        last_ep_return).  `num_timesteps`: env steps taken so far by THIS rank's envs (per env)."""
        sums = local_sums(env.timesteps, getattr(env, 'actions_taken', None), env.win_ring, env.last_ep_len,
                          env.last_ep_return)
        now = time.perf_counter()
        out = reduce_scalars(sums, num_timesteps - self._steps_last, now - self._t_last)
        self._t_last, self._steps_last = now, num_timesteps
        out['step'] = int(num_timesteps)
        if self.path and self.rank == 0:
            with open(self.path, 'a') as f:
                f.write(json.dumps(out) + '\n')
        return out


def write_evaluation_csv(path, original_state, env_index=0, aircraft=0):
    """original_state [T, N, A, 10] (info['original_state'] of a rollout) -> the reference's evaluation.csv lines."""
    rows = original_state[:, env_index, aircraft, :4].detach().cpu().tolist()
    with open(path, 'a+') as f:
        for x, y, h, phi in rows:
            f.write("%.2f, %.2f, %.0f, %.1f\n" % (x, y, h, phi))
    return len(rows)
