"""Host-side logic of the multi-GPU path (robo_b200/distributed.py) with world_size 2 on CPU
(gloo): shard bounds, the 16-byte {value, index} exchange and the numpy.argmax merge."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from robo_b200.distributed import merge_best, pack_pair, shard_bounds


def test_shard_bounds_cover_everything():
    for m in (0, 1, 7, 500, 2 ** 20 + 3):
        for world in (1, 2, 3, 8):
            got = [shard_bounds(m, r, world) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == m
            assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
            sizes = [hi - lo for lo, hi in got]
            assert max(sizes) - min(sizes) <= 1


def test_merge_matches_numpy_argmax():
    rng = np.random.RandomState(0)
    for trial in range(200):
        vals = rng.randint(0, 4, size=37).astype(float)
        if trial % 3 == 0:
            vals[rng.randint(0, 37, size=2)] = np.nan
        world = rng.randint(1, 6)
        pairs_v, pairs_i = [], []
        for r in range(world):
            lo, hi = shard_bounds(len(vals), r, world)
            if hi > lo:
                k = int(np.argmax(vals[lo:hi]))
                pairs_v.append(vals[lo + k])
                pairs_i.append(lo + k)
            else:
                pairs_v.append(0.0)
                pairs_i.append(-1)
        order = rng.permutation(world)            # gather order must not matter
        _, idx = merge_best([pairs_v[o] for o in order], [pairs_i[o] for o in order])
        assert idx == int(np.argmax(vals))


def _worker_thetas(rank, world, port, thetas, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from robo_b200.distributed import sharded_loglik
        vals = sharded_loglik(lambda th: -np.sum(th ** 2, axis=1), thetas, rank, world)
        out[rank] = vals.tolist()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sharded_theta_evaluation_gloo_world2():
    thetas = np.random.RandomState(2).randn(7, 3)          # 7 thetas over 2 ranks: uneven split
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_thetas, args=(2, _free_port(), thetas, out), nprocs=2, join=True)
    ref = -np.sum(thetas ** 2, axis=1)
    np.testing.assert_allclose(out[0], ref)
    np.testing.assert_allclose(out[1], ref)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, vals, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from robo_b200.distributed import allgather_best
        lo, hi = shard_bounds(len(vals), rank, world)
        k = int(np.argmax(vals[lo:hi]))
        v, i = allgather_best(pack_pair(vals[lo + k], lo + k))
        out[rank] = i
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_allgather_best_gloo_world2():
    vals = np.random.RandomState(1).rand(1001)
    vals[[100, 900]] = vals.max() + 1.0          # tie across the two shards: lowest index wins
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), vals, out), nprocs=world, join=True)
    assert out[0] == out[1] == 100 == int(np.argmax(vals))
