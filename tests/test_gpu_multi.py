"""Multi-GPU answer check through the C ABI (needs >= 2 GPUs; skipped otherwise): one process per GPU, NCCL bound by
libgpk.so itself (gpk_comm_init), no torch.distributed anywhere.  Every rank fits the same model (replicated fit
state), scores its contiguous shard of a COMMON candidate list, and the merged (value, global index) must equal the
single-GPU arg-max over the full list — robo/maximizers/random_sampling.py:48-50 semantics (first maximum wins)."""
import os
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _problem():
    rng = np.random.RandomState(21)
    N, D, M = 700, 6, 50000 + 3                       # uneven shards
    X = rng.rand(N, D)
    y = np.sinc(X * 10 - 5).sum(axis=1) + 0.01 * rng.randn(N)
    Xs = rng.rand(M, D)
    theta = np.concatenate(([0.0], np.full(D, np.log(D / 4.0))))
    return X, y, Xs, theta


def _worker(rank, world, id_file, out_dir):
    import torch
    from robo_b200 import _lib
    from robo_b200 import kernels as K
    from robo_b200.distributed import init_comm
    torch.cuda.set_device(rank)
    X, y, Xs, theta = _problem()
    D = X.shape[1]
    h = _lib.Handle(rank)
    h.set_data(X, y)
    f = K.Product(K.ConstantKernel(theta[0], ndim=D), K.Matern52Kernel(np.exp(theta[1:]), ndim=D)).flatten()
    h.set_kernel(f["family"], f["log_amp"], f["axis"], f["group"], f["log_metric"])
    h.fit(float(np.sqrt(np.float64(np.sqrt(1e-3)) ** 2 + 1.25e-12) ** 2), float(np.mean(y)))
    # plant an exact tie of the maximum across the two shards (every rank builds the same list): the LOWER index wins
    eta = float(np.min(y))
    r = h.acq(Xs, _lib.ACQ_EI, eta, 0.0, want_values=False)
    i_star = r["best_idx"]
    j = len(Xs) - 5 if i_star < len(Xs) // 2 else 7
    Xs[j] = Xs[i_star]
    expect_tie_winner = min(i_star, j)
    init_comm(h, rank, world, id_file=id_file)
    info = h.comm_info()
    assert info["rank"] == rank and info["world"] == world and info["nccl_version"] > 0
    res = {"tie": expect_tie_winner}
    for kind in (_lib.ACQ_EI, _lib.ACQ_LCB):
        res["host_%d" % kind] = h.acq_argmax_sharded(Xs, kind, eta, 0.0 if kind == _lib.ACQ_EI else 1.0)
    # device-resident shard, asynchronous variant
    lo, hi = _lib.shard_bounds(len(Xs), rank, world)
    d_X = torch.from_numpy(np.ascontiguousarray(Xs[lo:hi])).cuda()
    d_best = torch.zeros(2, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    h.acq_argmax_sharded_dev(d_X.data_ptr(), hi - lo, lo, _lib.ACQ_EI, eta, 0.0, d_best.data_ptr())
    h.synchronize()
    pair = d_best.cpu()
    res["dev"] = (float(pair[0]), int(pair[1:].view(torch.int64)[0]))
    # empty shard on the last rank
    m_small = world - 1
    res["tiny"] = h.acq_argmax_sharded(Xs[:m_small], _lib.ACQ_EI, eta, 0.0)
    # device-generated candidates
    lower, upper, inc = np.zeros(D), np.ones(D), X[np.argmin(y)]
    bx, bv, bi = h.maximize_random_sharded(777, 30001, 21000, lower, upper, inc, 0.1, _lib.ACQ_EI, eta, 0.0)
    res["rand"] = (bx.tolist(), bv, bi)
    if rank == 0:
        # single-GPU answers over the FULL list on the same device, no communicator involved
        g = _lib.Handle(rank)
        g.set_data(X, y)
        g.set_kernel(f["family"], f["log_amp"], f["axis"], f["group"], f["log_metric"])
        g.fit(float(np.sqrt(np.float64(np.sqrt(1e-3)) ** 2 + 1.25e-12) ** 2), float(np.mean(y)))
        for kind in (_lib.ACQ_EI, _lib.ACQ_LCB):
            r = g.acq(Xs, kind, eta, 0.0 if kind == _lib.ACQ_EI else 1.0, want_values=True)
            assert r["best_idx"] == int(np.argmax(r["values"]))
            res["single_%d" % kind] = (r["best_val"], r["best_idx"])
        r = g.acq(Xs[:m_small], _lib.ACQ_EI, eta, 0.0, want_values=False)
        res["single_tiny"] = (r["best_val"], r["best_idx"])
        x1, v1, i1 = g.maximize_random(777, 0, 30001, 21000, lower, upper, inc, 0.1, _lib.ACQ_EI, eta, 0.0)
        res["single_rand"] = (x1.tolist(), v1, i1)
        g.close()
    h.comm_destroy()
    h.close()
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), np.array([res], dtype=object), allow_pickle=True)


@pytest.mark.timeout(600)
def test_sharded_argmax_equals_single_gpu_argmax_nccl_world2():
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_worker, args=(world, os.path.join(tmp, "nccl_id"), tmp), nprocs=world, join=True)
        res = [np.load(os.path.join(tmp, "rank%d.npy" % r), allow_pickle=True)[0] for r in range(world)]
    r0, r1 = res
    for kind in (1, 4):
        assert r0["host_%d" % kind] == r1["host_%d" % kind] == r0["single_%d" % kind]
    assert r0["single_1"][1] == r0["tie"] == r1["tie"]  # the planted tie resolves to the lower index
    assert r0["dev"] == r1["dev"] == r0["single_1"]
    assert r0["tiny"] == r1["tiny"] == r0["single_tiny"]
    assert r0["rand"] == r1["rand"] == r0["single_rand"]
