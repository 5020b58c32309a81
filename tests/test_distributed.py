"""World-size-2 gloo test (CPU) of the multi-GPU plumbing: env sharding + return statistics.
Each rank steps its shard with the same seed and its own env_offset; the all-reduced statistics
must equal those of the unsharded batch (results are independent of the number of ranks because
Philox counters use global env ids)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pyrddlgym_b200 import parallel


def _rollout_returns(n_envs, env_offset, steps=4):
    from oracle.oracle_sim import OracleSimulator
    from pyrddlgym_b200 import workloads, lowering
    p = workloads.load_program(workloads.wildfire(32))
    tab = lowering.lower(p)
    sim = OracleSimulator(p, n_envs=n_envs, seed=5, env_offset=env_offset, bitsliced_nodes=tab.plane_nodes)
    ret = np.zeros(n_envs)
    for t in range(steps):
        rng = np.random.default_rng(100 + t)                       # same global action draw on every rank
        put = rng.random((6, 32, 32)) < 0.05
        cut = rng.random((6, 32, 32)) < 0.05
        sl = slice(env_offset, env_offset + n_envs)
        _, r, _ = sim.step({'put-out': put[sl], 'cut-out': cut[sl]})
        ret += r * (0.9 ** t)
    return ret


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    n_local, off = parallel.shard_envs(6, world, rank)
    ret = _rollout_returns(n_local, off)
    stats = parallel.reduce_return_stats(torch.from_numpy(ret))
    stats['median'] = parallel.gather_median(torch.from_numpy(ret))
    if rank == 0:
        torch.save(stats, out)
    dist.destroy_process_group()


def test_shard_envs_partitions_contiguously():
    for total, world in [(4096, 8), (10, 4), (3, 8), (6, 2)]:
        blocks = [parallel.shard_envs(total, world, r) for r in range(world)]
        assert sum(n for n, _ in blocks) == total
        nxt = 0
        for n, off in blocks:
            assert off == nxt
            nxt += n


def test_return_statistics_are_shard_invariant(tmp_path):
    out = str(tmp_path / 'stats.pt')
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    full = _rollout_returns(6, 0)
    assert got['count'] == 6
    np.testing.assert_allclose(got['mean'], full.mean(), rtol=1e-12)
    np.testing.assert_allclose(got['std'], full.std(), rtol=1e-9, atol=1e-9)
    assert got['min'] == full.min() and got['max'] == full.max()
    np.testing.assert_allclose(got['median'], np.median(full), rtol=1e-12)
    single = parallel.reduce_return_stats(torch.from_numpy(full))
    np.testing.assert_allclose(single['mean'], got['mean'], rtol=1e-12)
