"""Episode-metric log: the reference callback's scalars as reductions over the per-env counters,
on CPU tensors here (the same code runs on the env's cuda tensors), single process and world size 2 over gloo."""
import json
import os
import socket
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _counters(n, seed):
    rng = np.random.RandomState(seed)
    ts = rng.randint(1, 500, n).astype(np.int32)
    at = (ts * rng.uniform(0, 1.5, n)).astype(np.int32)
    ring = rng.randint(0, 1 << 16, n).astype(np.int32)
    ll = np.where(rng.uniform(size=n) < 0.7, rng.randint(1, 900, n), 0).astype(np.int32)
    lr = np.where(ll > 0, rng.uniform(-300, 300, n), 0.0)
    return types.SimpleNamespace(timesteps=torch.from_numpy(ts), actions_taken=torch.from_numpy(at),
                                 win_ring=torch.from_numpy(ring), last_ep_len=torch.from_numpy(ll),
                                 last_ep_return=torch.from_numpy(lr))


def _expected(c):
    ts, at = c.timesteps.numpy().astype(np.float64), c.actions_taken.numpy().astype(np.float64)
    ring, ll, lr = c.win_ring.numpy(), c.last_ep_len.numpy(), c.last_ep_return.numpy()
    wins = np.array([bin(int(r) & 0x1FF).count('1') for r in ring]) * 0.1          # atc_gym.py:359-363 (9-deep window)
    fin = ll > 0
    return {'simulation/mean_actions': float(np.mean(at / ts)),                    # atc_gym.py:197
            'simulation/winning_ratio': float(np.mean(wins)),
            'simulation/mean_episode_length': float(ll[fin].mean()),
            'simulation/mean_episode_return': float(lr[fin].mean()),
            'simulation/episodes': float(fin.sum())}


def test_scalars_match_a_numpy_restatement(tmp_path):
    from atc_reinforcement_learning_b200.metrics import EpisodeMetricsLog, TAGS
    c = _counters(1000, 3)
    path = str(tmp_path / 'scalars.jsonl')
    log = EpisodeMetricsLog(path)
    out = log.log(c, 2048)
    exp = _expected(c)
    for k, v in exp.items():
        assert out[k] == pytest.approx(v, rel=1e-12), k
    assert out['simulation/fps'] > 0 and out['step'] == 2048 and set(TAGS) <= set(out)
    log.log(c, 4096)
    lines = [json.loads(l) for l in open(path)]
    assert [l['step'] for l in lines] == [2048, 4096]


def test_no_finished_episode_gives_nan_not_zero():
    from atc_reinforcement_learning_b200.metrics import EpisodeMetricsLog
    c = _counters(8, 1)
    c.last_ep_len.zero_()
    out = EpisodeMetricsLog().log(c, 10)
    assert np.isnan(out['simulation/mean_episode_length']) and out['simulation/episodes'] == 0


def test_evaluation_csv_format(tmp_path):
    from atc_reinforcement_learning_b200.metrics import write_evaluation_csv
    raw = torch.zeros(3, 2, 1, 10)
    raw[:, 1, 0, :4] = torch.tensor([[10.126, 20.5, 9000.4, 123.45], [10.2, 20.6, 8959.0, 126.4], [0, 0, 0, 0]])
    p = str(tmp_path / 'evaluation.csv')
    assert write_evaluation_csv(p, raw, env_index=1) == 3
    assert open(p).read().splitlines()[:2] == ['10.13, 20.50, 9000, 123.4', '10.20, 20.60, 8959, 126.4']


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_global, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from atc_reinforcement_learning_b200.dist import init_from_env, shard_envs
    from atc_reinforcement_learning_b200.metrics import EpisodeMetricsLog
    init_from_env('gloo')
    n_local, base = shard_envs(n_global, rank, world)
    full = _counters(n_global, 11)
    shard = types.SimpleNamespace(**{k: v[base:base + n_local] for k, v in vars(full).items()})
    out = EpisodeMetricsLog(os.path.join(out_dir, 'scalars.jsonl')).log(shard, 100)
    with open(os.path.join(out_dir, 'rank%d.json' % rank), 'w') as f:
        json.dump(out, f)
    dist.destroy_process_group()


def test_two_ranks_report_the_global_means(tmp_path):
    n_global, world = 101, 2
    mp.spawn(_worker, args=(world, _free_port(), n_global, str(tmp_path)), nprocs=world, join=True)
    exp = _expected(_counters(n_global, 11))
    for rank in range(world):
        got = json.load(open(os.path.join(str(tmp_path), 'rank%d.json' % rank)))
        for k, v in exp.items():
            assert got[k] == pytest.approx(v, rel=1e-12), (rank, k)
    assert len(open(os.path.join(str(tmp_path), 'scalars.jsonl')).read().splitlines()) == 1      # rank 0 only


@pytest.mark.gpu
def test_env_counters_on_the_device_give_the_oracles_scalars():
    """The log over a BatchedAtcEnv's cuda counters after a rollout == the same reductions over the CPU oracle's."""
    from atc_reinforcement_learning_b200 import BatchedAtcEnv, LOWW, SimParameters
    from atc_reinforcement_learning_b200.metrics import EpisodeMetricsLog
    from oracle.oracle import Oracle
    N, A, T = 512, 4, 400
    env = BatchedAtcEnv(N, A, SimParameters(1), LOWW(random_entrypoints=True), seed=12, track_actions=True)
    ora = Oracle('LOWW', True, n_env=N, n_ac=A, seed=12)
    ora.reset(); ora.reset(); env.reset()
    rng = np.random.RandomState(5)
    acts = np.repeat(rng.uniform(-1, 1, (T // 20, N, A, 3)).astype(np.float32), 20, 0)
    env.rollout(torch.from_numpy(acts).cuda())
    ora.rollout(acts)
    m, (_, ots) = ora.metrics(), ora.get_state()
    c = types.SimpleNamespace(timesteps=torch.from_numpy(ots), actions_taken=torch.from_numpy(m['actions_taken']),
                              win_ring=torch.from_numpy(m['win_ring']), last_ep_len=torch.from_numpy(m['last_ep_len']),
                              last_ep_return=torch.from_numpy(m['last_ep_return']))
    got, exp = EpisodeMetricsLog().log(env, T), EpisodeMetricsLog().log(c, T)
    assert got['simulation/episodes'] == exp['simulation/episodes'] > 50
    for k in ('simulation/mean_actions', 'simulation/winning_ratio', 'simulation/mean_episode_length'):
        assert got[k] == exp[k], k
    assert got['simulation/mean_episode_return'] == pytest.approx(exp['simulation/mean_episode_return'], rel=1e-5)


@pytest.mark.gpu
def test_vecenv_surface_step_async_env_method_attrs():
    """The VecEnv calls the reference's scripts make through stable-baselines wrappers
    (learning/atc-gym-stable-baselines.py:34-36,76-85): step_async / step_wait == step, get_attr / set_attr per env,
    env_method on the batch.  (SB 2.8.0 itself is not under /root/reference: unpinned.)"""
    from atc_reinforcement_learning_b200 import BatchedAtcEnv, LOWW, SimParameters
    N, A = 64, 2
    mk = lambda: BatchedAtcEnv(N, A, SimParameters(1), LOWW(random_entrypoints=True), seed=3, track_actions=True)
    e1, e2 = mk(), mk()
    a = torch.rand(N, A, 3, device='cuda') * 2 - 1
    for _ in range(5):
        o1, r1, d1, i1 = e1.step(a)
        e2.step_async(a)
        with pytest.raises(RuntimeError):
            e2.step_async(a)
        o2, r2, d2, i2 = e2.step_wait()
        assert torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(d1, d2)
        assert torch.equal(i1['original_state'], i2['original_state'])
    with pytest.raises(RuntimeError):
        e2.step_wait()
    wr = e1.get_attr('winning_ratio')
    assert len(wr) == N and all(w == 0.0 for w in wr)
    assert e1.get_attr('timesteps', indices=[0, 5]) == [5, 5]
    assert e1.get_attr('timestep_limit', indices=3) == [6000]
    apt = e1.get_attr('actions_per_timestep')
    assert len(apt) == N and all(0.0 <= v <= 3.0 * A for v in apt)
    e1.set_attr('timesteps', 17, indices=[2])
    assert e1.get_attr('timesteps', indices=[1, 2]) == [5, 17]
    with pytest.raises(ValueError):
        e1.set_attr('timestep_limit', 10, indices=[0])
    st = e1.env_method('get_state', indices=[0])
    assert len(st) == 1 and st[0][0].shape == (N, A, 5)
    assert e1.env_method('seed', 9) == [[9]] * N
    e1.close(); e2.close()
