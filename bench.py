#!/usr/bin/env python
"""bench.py — EI evaluations / second + GP fit time on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--m M]

Workload (BASELINE.json configs[1] / SURVEY.md section 8d "C2"): GP posterior with N = 4096
training points, D = 16, amp * ARD Matern-5/2, fp64, synthetic seeded data; one *step* scores a
batch of M candidates per GPU (predict -> EI -> arg-max) against the fitted model.

Our arm (default): `value` = candidates scored per second with the candidates resident in HBM
(gpk_acq_dev), timed with CUDA events on the launching stream, max over ranks; `e2e` = the same
through the host-buffer C-ABI call a RoBO user makes (gpk_acq: pinned-host candidates copied
H2D and the arg-max read back D2H inside the timed region).  Multi-GPU: one process per GPU
(torchrun), candidates sharded (weak scaling: M per GPU), fit state replicated, one 16-byte
all_gather per step for the arg-max (robo_b200/distributed.py).

Reference arm (--impl reference): the reference's CPU implementation of the same path.  george
is an un-vendored third-party dependency that cannot be installed here,
so the arm runs the oracle port (oracle/robo_oracle.py: reference-faithful predict with the full
M x M covariance, then EI — gaussian_process.py:280-294 + ei.py:65-78) on the host cores with all
BLAS threads, in the reference's own batch size of 500 candidates (random_sampling.py:9).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_TRAIN, DIM = 4096, 16
METRIC = "EI evals/sec (GP N=4096, D=16, fp64)"
FP64_PEAK_TFLOPS = 37.0      # SURVEY.md section 8d "P64": B200 FP64 / FP64-tensor datasheet (296 TF per 8-GPU HGX)


def problem(m, seed_cand=4321):
    """SURVEY.md section 8d synthetic inputs (restated here so the product arm does not import oracle/)."""
    rng = np.random.RandomState(1234)
    X = rng.rand(N_TRAIN, DIM)
    y = np.sinc(X * 10 - 5).sum(axis=1) + 0.01 * rng.randn(N_TRAIN)
    Xs = np.random.RandomState(seed_cand).rand(m, DIM)
    theta = np.concatenate(([0.0], np.full(DIM, np.log(DIM / 4.0))))
    return X, y, Xs, theta, 1e-3


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.lines, self.proc, self.gpu, self.t_timed = [], None, gpu_index, None

    def mark_timed_region(self):
        self.t_timed = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        timed = [ln for t, ln in self.lines if self.t_timed is None or t >= self.t_timed]
        for ln in (timed if timed else [ln for _, ln in self.lines]):
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_rate(n_batches, batch=500, return_state=False):
    """Reference path on the host cores via the oracle port: predict (full cov) + EI per batch of 500."""
    from oracle import robo_oracle as O
    X, y, Xs, theta, noise = problem(batch * max(n_batches, 1))
    t0 = time.perf_counter()
    st = O.gp_fit(O.make_kernel("matern52", DIM, theta), X, y, noise=noise, normalize_input=False)
    fit_s = time.perf_counter() - t0
    eta = O.gp_get_incumbent(st)[1]
    times = []
    best = None
    for b in range(n_batches):
        t0 = time.perf_counter()
        m, v = O.gp_predict(st, Xs[b * batch:(b + 1) * batch])
        ei = O.acq_ei(m, v, eta)
        best = int(np.argmax(ei))
        times.append(time.perf_counter() - t0)
    return fit_s, times, best


def run_reference(args, rank):
    if rank != 0:
        return
    import threadpoolctl
    cores = os.cpu_count()
    # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host thread it can
    threadpoolctl.threadpool_limits(limits=cores)
    fit_s, times, _ = cpu_reference_rate(args.warmup + args.steps)
    timed = times[args.warmup:]
    total = float(np.sum(timed))
    value = 500 * len(timed) / total
    try:
        blas_threads = max(p.get("num_threads", 1) for p in threadpoolctl.threadpool_info())
    except Exception:
        blas_threads = cores
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "EI evals/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(timed),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "GP posterior N=4096 D=16 Matern52 fp64: predict + EI + argmax (configs[1])",
                   "n_train": N_TRAIN, "dim": DIM, "candidates_per_step": 500,
                   "note": "reference CPU path (oracle port of gaussian_process.py:280-294 + ei.py:65-78, full "
                           "M x M covariance per 500-candidate batch as random_sampling.py:9); george itself is "
                           "not installable here"},
        "cpu_baseline": {"value": value, "unit": "EI evals/s", "cores": blas_threads, "kind": "port",
                         "sample": "%d batches of 500 candidates, N=4096 D=16; fit (K build + Cholesky) %.2f s not "
                                   "included" % (len(timed), fit_s)},
        "e2e": {"value": value, "unit": "EI evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "fit_ms": 1e3 * fit_s, "gpu_launches": 0,
    }
    print(json.dumps(line))


def measure_dgemm_tflops(torch, dev):
    """cuBLAS DGEMM on this GPU as the practical fp64 tensor-pipe reference (reported, not on our path)."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    del a, b
    return best


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from robo_b200 import _lib
    from robo_b200 import kernels as K
    from robo_b200.distributed import allgather_best, pack_pair

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback in the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    M = args.m
    X, y, Xs, theta, noise = problem(M, seed_cand=4321 + rank)

    h = _lib.Handle(local_rank)
    # a real (non-default) torch stream: the handle launches on it, and torch.cuda.Event timing sees it
    # high priority like the handle's own stream: the single-CTA kernels of the Cholesky chain must not queue behind the
    # trailing-update tiles the handle launches on its low-priority side stream
    stream = torch.cuda.Stream(device=dev, priority=-1)
    torch.cuda.set_stream(stream)
    h.set_stream(stream.cuda_stream)
    h.set_option("chunk", args.chunk)
    h.set_data(X, y)
    f = K.Product(K.ConstantKernel(theta[0], ndim=DIM), K.Matern52Kernel(np.exp(theta[1:]), ndim=DIM)).flatten()
    h.set_kernel(f["family"], f["log_amp"], f["axis"], f["group"], f["log_metric"])
    diag_add = float(np.sqrt(np.float64(np.sqrt(noise)) ** 2 + 1.25e-12) ** 2)
    mean = float(np.mean(y))
    eta = float(np.min(y))

    # ---- GP fit (K build + Cholesky + forward solve + log-det), then L^-1 for scoring ----
    fit_ms = []
    for _ in range(4):
        logdet, ll = h.fit(diag_add, mean)
        fit_ms.append(h.timings()["fit_ms"])
    d_X = torch.from_numpy(Xs).to(dev)
    d_best = torch.zeros(2, dtype=torch.float64, device=dev)
    h.acq_dev(d_X.data_ptr(), min(M, 1024), _lib.ACQ_EI, eta, 0.0, 0, 0, 0, d_best.data_ptr())   # builds L^-1
    torch.cuda.synchronize()
    t_fit = h.timings()

    def step_dev():
        h.acq_dev(d_X.data_ptr(), M, _lib.ACQ_EI, eta, 0.0, 0, 0, 0, d_best.data_ptr())
        if world > 1:
            return allgather_best(d_best)
        return None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()            # started before the warm-up so that samples exist even for short timed regions
    for _ in range(args.warmup):
        step_dev()
    launches0 = h.timings()["launches_total"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_timed_region()
    e0.record(stream)
    for _ in range(args.steps):
        step_dev()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    tim = h.timings()
    launches = tim["launches_total"] - launches0
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * M * args.steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel (variance GEMM), measured live with CUDA events ----
    # algorithmic flops per launch: every candidate row of the chunk contracts with the lower
    # triangle of L^-1 (N^2/2 FMA = N^2 flop) plus the mean reduction (2N)   [SURVEY.md 8d: F_ei ~ N^2]
    rows = min(args.chunk, ((M + 127) // 128) * 128)
    last_rows = rows if M > rows else M          # timings() averages the full-size chunk launches of the last step
    gemm_ms = tim["vargemm_ms"]
    flops = float(last_rows) * (N_TRAIN ** 2 + 2 * N_TRAIN)
    achieved = flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0

    # ---- end to end through the host-buffer C ABI (pinned host candidates, H2D + D2H timed) ----
    Xs_pinned = torch.from_numpy(Xs).pin_memory()
    Xs_host = Xs_pinned.numpy()
    for _ in range(max(1, args.warmup // 2)):
        h.acq(Xs_host, _lib.ACQ_EI, eta, 0.0, want_values=False)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        r = h.acq(Xs_host, _lib.ACQ_EI, eta, 0.0, want_values=False)
        if world > 1:
            allgather_best(pack_pair(r["best_val"], r["best_idx"], dev))
    e1.record(stream)
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e_value = world * M * args.steps / (ms_e2e * 1e-3)

    if rank != 0:
        return
    dgemm = measure_dgemm_tflops(torch, dev)
    dmma_peak, dfma_peak = h.measure_fp64_peaks()
    # ---- CPU baseline on this box's host cores (bounded sample) ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=os.cpu_count())
        fit_s, times, _ = cpu_reference_rate(6)
        timed = times[1:]
        try:
            blas_threads = max(p.get("num_threads", 1) for p in threadpoolctl.threadpool_info())
        except Exception:
            blas_threads = os.cpu_count()
        cpu = {"value": 500 * len(timed) / float(np.sum(timed)), "unit": "EI evals/s", "cores": blas_threads,
               "kind": "port", "fit_ms": 1e3 * fit_s,
               "sample": "%d batches of 500 candidates (reference batch size, full M x M covariance), "
                         "N=4096 D=16, oracle port of gaussian_process.py:280-294 + ei.py:65-78" % len(timed)}
    # incremental refit (gpk_fit_append): the last 8 rows appended to a model fitted on N - 8 rows (extra info, untimed
    # with respect to the headline; rank 0 only)
    append_ms = None
    if rank == 0:
        try:
            for rep in range(2):                                 # the first pass loads the kernels of this path
                h2 = _lib.Handle(dev.index or 0)
                h2.set_data(X[:N_TRAIN - 8], y[:N_TRAIN - 8])
                h2.set_kernel(f["family"], f["log_amp"], f["axis"], f["group"], f["log_metric"])
                h2.fit(diag_add, float(np.mean(y[:N_TRAIN - 8])))
                h2.predict(Xs[:128])
                res = h2.fit_append(X, y, diag_add, mean)
                if res is not None:
                    append_ms = h2.timings()["fit_ms"]
                    assert abs(res[1] - ll) <= 1e-10 * abs(ll), "incremental refit disagrees with the full fit"
                h2.close()
        except Exception as e:                                   # noqa: BLE001
            print("fit_append timing skipped: %r" % (e,), file=sys.stderr)
    line = {
        "metric": METRIC, "value": value, "unit": "EI evals/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "GP posterior N=4096 D=16 Matern52 fp64: predict + EI + argmax (configs[1])",
                   "n_train": N_TRAIN, "dim": DIM, "candidates_per_step_per_gpu": M, "chunk": args.chunk,
                   "parallelism": "candidate shards x%d, fit state replicated, 16 B all_gather per step" % world,
                   "l2": "working set per step (L^-1 134 MB + K* chunk %d MB) exceeds the 126 MB L2; no flush needed"
                         % (rows * N_TRAIN * 8 // 2 ** 20)},
        "fit_ms": float(np.median(fit_ms[1:])), "fit_breakdown_ms": {k: t_fit[k] for k in ("kbuild_ms", "potrf_ms", "linv_ms")},
        "loglik": ll, "fit_append_8rows_ms": append_ms,
        "e2e": {"value": e2e_value, "unit": "EI evals/s", "h2d_bytes_per_step": int(M * DIM * 8),
                "d2h_bytes_per_step": 24, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "gpk_gemm_ws_kernel<EPI_COLREDUCE> (L^-1 K*^T contraction, fp64 DMMA, warp-specialised TMA)",
                     "achieved": achieved, "peak": dmma_peak, "unit": "TFLOP/s", "frac": achieved / dmma_peak,
                     "peak_source": "measured live on this GPU: register-resident DMMA m8n8k4 issue rate "
                                    "(gpk_measure_fp64_peaks); MEASURED_PEAKS.json has no fp64 figure; datasheet "
                                    "FP64-tensor = %.0f TF/s (SURVEY 8d P64)" % FP64_PEAK_TFLOPS,
                     "frac_of_datasheet": achieved / FP64_PEAK_TFLOPS, "dfma_vector_peak_tflops": dfma_peak,
                     "dgemm_cublas_tflops": dgemm, "frac_of_cublas_dgemm": achieved / dgemm if dgemm > 0 else None,
                     "launch_ms": gemm_ms, "launch_candidates": int(last_rows),
                     "launches_averaged": int(max(1, (M + rows - 1) // rows - 1)) if M > rows else 1,
                     "traffic": 1.5977e9 if last_rows == 16384 else None,
                     "traffic_source": "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum of one 16384-candidate "
                                       "launch (profiles/r01b_vargemm_ws_ncu_full_raw.csv: 1.584 GB read + 13.3 MB written); algorithmic minimum 0.60e9 "
                                       "(L^-1 lower triangle 67 MB + K* 537 MB read once)"},
        "kernel_ms_last_chunk": {k: tim[k] for k in ("kstar_ms", "vargemm_ms", "finish_ms")},
        "cpu_baseline": cpu, "clocks": clocks,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--m", type=int, default=131072, help="candidates per step per GPU")
    ap.add_argument("--chunk", type=int, default=16384)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
