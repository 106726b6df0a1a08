"""world_size-2 gloo test (CPU) of the multi-GPU host logic: contiguous env sharding with per-rank env_index_base, and the
one collective on the path — the all_gather of the per-env episode-return log (dist.ReturnGather).  The oracle stands
in for the kernel as the per-rank env (it is the checker: a rank's shard must equal that slice of the one-rank job)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_global, n_ac, T, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from atc_reinforcement_learning_b200.dist import ReturnGather, init_from_env, shard_envs
    from oracle.oracle import Oracle, set_num_threads
    set_num_threads(1)
    r, w, _ = init_from_env('gloo')
    assert (r, w) == (rank, world)
    n_local, base = shard_envs(n_global, rank, world)
    rng = np.random.RandomState(0)
    acts = np.repeat(rng.uniform(-1, 1, (T // 20 + 1, n_global, n_ac, 3)).astype(np.float32), 20, 0)[:T]
    env = Oracle('LOWW', True, n_env=n_local, n_ac=n_ac, seed=5, env_index_base=base)
    env.reset()
    obs, rew, done, term = env.rollout(acts[:, base:base + n_local])
    gather = ReturnGather(n_local, 'cpu')
    got = gather.gather(torch.from_numpy(env.metrics()['last_ep_return']).float())
    gather.wait()
    tot = torch.tensor([float(done.sum())])
    dist.all_reduce(tot)
    np.savez(os.path.join(out_dir, 'rank%d.npz' % rank), rew=rew, done=done, gathered=got.numpy(), base=base,
             n_local=n_local, total_done=tot.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize('n_global', [96, 97])
def test_two_rank_sharding_and_return_gather(tmp_path, n_global):
    """97 envs on 2 ranks: unequal shards (49 + 48) — the gather pads to the largest shard and drops the padding."""
    from atc_reinforcement_learning_b200.dist import shard_envs
    from oracle.oracle import Oracle
    n_ac, T, world = 4, 300, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_global, n_ac, T, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.RandomState(0)
    acts = np.repeat(rng.uniform(-1, 1, (T // 20 + 1, n_global, n_ac, 3)).astype(np.float32), 20, 0)[:T]
    full = Oracle('LOWW', True, n_env=n_global, n_ac=n_ac, seed=5)
    full.reset()
    obs, rew, done, term = full.rollout(acts)
    ref_returns = full.metrics()['last_ep_return'].astype(np.float32)
    assert done.sum() > 0
    for rank in range(world):
        z = np.load(os.path.join(str(tmp_path), 'rank%d.npz' % rank))
        n_local, base = shard_envs(n_global, rank, world)
        assert (int(z['base']), int(z['n_local'])) == (base, n_local)
        np.testing.assert_array_equal(z['rew'], rew[:, base:base + n_local])
        np.testing.assert_array_equal(z['done'], done[:, base:base + n_local])
        np.testing.assert_array_equal(z['gathered'], ref_returns)       # rank order == global env order
        assert float(z['total_done'][0]) == float(done.sum())


def test_shard_envs_partitions():
    from atc_reinforcement_learning_b200.dist import shard_envs
    for n, w in ((131072, 8), (16384, 1), (10, 3), (7, 7), (100, 8)):
        parts = [shard_envs(n, r, w) for r in range(w)]
        assert sum(p[0] for p in parts) == n
        assert parts[0][1] == 0 and all(parts[i][1] + parts[i][0] == parts[i + 1][1] for i in range(w - 1))
        assert max(p[0] for p in parts) - min(p[0] for p in parts) <= 1
    with pytest.raises(ValueError):
        shard_envs(3, 0, 4)
