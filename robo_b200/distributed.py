"""Multi-GPU scoring: candidates shard across ranks, the fit state is replicated, and the only
exchange on the path is the arg-max.

NCCL has no MAXLOC, so each rank contributes its 16-byte {value, global index} pair to one
all_gather (torch.distributed: NCCL on GPUs, gloo on CPU for the host-logic tests) and every
rank applies the same deterministic merge: numpy.argmax ordering — NaN first, then the larger
value, then the LOWEST global index (robo/maximizers/random_sampling.py:50 takes the first
maximum).  Payload is 16 B per rank, so the exchange is latency-bound and candidate throughput
scales with the number of GPUs.
"""
import numpy as np


def shard_bounds(m, rank, world):
    """Contiguous slice [lo, hi) of m candidates owned by ``rank`` (sizes differ by <= 1)."""
    base, rem = divmod(int(m), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def better(va, ia, vb, ib):
    """True if (va, ia) beats (vb, ib) under numpy.argmax ordering; index < 0 means 'empty'."""
    if ib < 0:
        return ia >= 0
    if ia < 0:
        return False
    na, nb = np.isnan(va), np.isnan(vb)
    if na or nb:
        if na and nb:
            return ia < ib
        return bool(na)
    if va > vb:
        return True
    if va < vb:
        return False
    return ia < ib


def merge_best(values, indices):
    """Deterministic reduction of per-rank (value, global index) pairs."""
    bv, bi = 0.0, -1
    for v, i in zip(values, indices):
        if better(float(v), int(i), bv, bi):
            bv, bi = float(v), int(i)
    return bv, bi


def allgather_best(pair, group=None):
    """pair: torch tensor of 2 float64 {value, index bit-cast to float64} on this rank's device
    (the 16-byte struct gpk_acq_dev writes).  Returns (value, global index) after the exchange."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = torch.empty(world * 2, dtype=torch.float64, device=pair.device)
    dist.all_gather_into_tensor(out, pair.contiguous(), group=group)
    host = out.cpu().view(world, 2)
    vals = host[:, 0].numpy()
    idxs = host[:, 1].contiguous().view(torch.int64).numpy()
    return merge_best(vals, idxs)


def pack_pair(value, index, device="cpu"):
    """Host helper mirroring the device struct layout (used by the gloo tests)."""
    import torch
    t = torch.empty(2, dtype=torch.float64)
    t[0] = value
    t[1:].view(torch.int64)[0] = int(index)
    return t.to(device)


def sharded_argmax(model, acq_kind, X_all, rank, world, eta=None, par=0.0, group=None):
    """Score this rank's contiguous shard of X_all (host array, identical on all ranks) with the
    fused GPU path and agree on the global arg-max.  Returns (best_value, best_global_index)."""
    import torch
    lo, hi = shard_bounds(len(X_all), rank, world)
    if hi > lo:
        r = model.score(np.ascontiguousarray(X_all[lo:hi]), acq_kind, eta=eta, par=par, want_values=False)
        val, idx = r["best_val"], (r["best_idx"] + lo if r["best_idx"] >= 0 else -1)
    else:
        val, idx = 0.0, -1
    dev = "cuda:%d" % torch.cuda.current_device() if torch.cuda.is_available() else "cpu"
    return allgather_best(pack_pair(val, idx, dev), group)


def sharded_loglik(eval_fn, thetas, rank, world, group=None):
    """Hyper-parameter vectors are independent too (SURVEY.md section 8e "other shardable axes": MCMC walkers,
    theta sweeps): rank r evaluates the contiguous slice of ``thetas`` it owns with ``eval_fn(theta_block) ->
    values`` and one all_gather of the values (8 bytes per theta) gives every rank the full vector."""
    import torch
    import torch.distributed as dist
    thetas = np.asarray(thetas, dtype=np.float64)
    n = len(thetas)
    per = (n + world - 1) // world                     # equal-sized slots so that all_gather_into_tensor applies
    lo, hi = min(rank * per, n), min((rank + 1) * per, n)
    mine = np.full(per, np.nan)
    if hi > lo:
        mine[:hi - lo] = np.asarray(eval_fn(thetas[lo:hi]), dtype=np.float64)
    dev = "cuda:%d" % torch.cuda.current_device() if torch.cuda.is_available() and dist.get_backend(group) == "nccl" else "cpu"
    out = torch.empty(per * world, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, torch.from_numpy(mine).to(dev), group=group)
    host = out.cpu().numpy()
    return np.concatenate([host[r * per:r * per + max(0, min(per, n - r * per))] for r in range(world)])
