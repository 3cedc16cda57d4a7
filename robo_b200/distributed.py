"""Multi-GPU scoring: candidates shard across ranks, the fit state is replicated, and the only
exchange on the path is the arg-max.

NCCL has no MAXLOC, so each rank contributes its 16-byte {value, global index} pair to one
all-gather and every rank applies the same deterministic merge: numpy.argmax ordering — NaN first,
then the larger value, then the LOWEST global index (robo/maximizers/random_sampling.py:50 takes the
first maximum).  Payload is 16 B per rank, so the exchange is latency-bound and candidate throughput
scales with the number of GPUs.

On GPUs the exchange lives behind the C ABI (include/gpk.h: gpk_comm_init, gpk_acq_argmax_sharded,
gpk_acq_argmax_sharded_dev, gpk_maximize_random_sharded): ncclAllGather on the handle's stream, merge
kernel, one 16-byte D2H of the winner — no torch op and no host synchronisation between scoring and
exchange.  ``init_comm`` only has to get rank 0's 128-byte NCCL id to the other ranks; any launcher
channel will do (a torch.distributed store under torchrun, a shared file otherwise).  The
torch.distributed functions below (``allgather_best`` ...) are the same host logic over gloo for the
CPU tests and for models that are not device GPs.
"""
import os
import time

import numpy as np


def init_comm(handle, rank, world, group=None, id_file=None, timeout_s=120.0):
    """gpk_comm_init on ``handle`` (robo_b200._lib.Handle).  The 128-byte id made on rank 0 travels through
    torch.distributed (object broadcast, when a process group is initialised) or through ``id_file`` (rank 0 writes
    it atomically, the others poll)."""
    from robo_b200 import _lib
    if world <= 1:
        handle.comm_init(0, 1, None)
        return
    uid = _lib.comm_unique_id() if rank == 0 else None
    if id_file is not None:
        if rank == 0:
            tmp = "%s.%d.tmp" % (id_file, os.getpid())
            with open(tmp, "wb") as f:
                f.write(uid)
            os.replace(tmp, id_file)
        else:
            t0 = time.time()
            while not os.path.exists(id_file):
                if time.time() - t0 > timeout_s:
                    raise RuntimeError("init_comm: rank 0 never wrote %s" % id_file)
                time.sleep(0.01)
            with open(id_file, "rb") as f:
                uid = f.read()
    else:
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("init_comm: pass id_file=... or initialise torch.distributed (used only to ship the id)")
        box = [uid]
        dist.broadcast_object_list(box, src=0, group=group)
        uid = box[0]
    handle.comm_init(rank, world, uid)


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
    fused GPU path and agree on the global arg-max.  Returns (best_value, best_global_index).
    A device GP whose handle carries a communicator (init_comm) does all of it in one C-ABI call."""
    gp = getattr(model, "gp", None)
    handle = getattr(gp, "_handle", None)
    if handle is not None and hasattr(handle, "comm_info") and handle.comm_info()["world"] == world and world > 1:
        from robo_b200 import _lib
        if eta is None:
            eta = 0.0 if acq_kind == "lcb" else model.get_incumbent()[1]
        gp._restore()
        gp._push_cfg()
        return handle.acq_argmax_sharded(X_all, _lib.ACQ_KIND[acq_kind], float(eta), float(par))
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
