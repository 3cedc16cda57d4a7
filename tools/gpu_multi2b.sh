#!/bin/bash
# 2-GPU check of the final state: NCCL arg-max exchange test through the C ABI + torchrun bench line.
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider > $out/r2n_pytest_multi.log 2>&1; echo "pytest[multi] exit $?"; tail -3 $out/r2n_pytest_multi.log
NCCL_DEBUG=WARN timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 10 --warmup 3 > $out/r2_bench_2gpu.json 2> $out/r2_bench_2gpu.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_2gpu.json").read().strip().splitlines()[-1])
    c3 = d["configs"]["c3"]
    print("2 GPUs: value %.4g ms/step %.2f e2e %.4g argmax_check %s c3 host %.2f ms dev %.2f ms check %s" % (
        d["value"], d["ms_per_step"], d["e2e"]["value"], d["argmax_check"], c3["host_pageable"]["wall_ms"], c3["device_philox"]["wall_ms"], c3["argmax_check"]))
except Exception as e:
    print("2-GPU bench failed:", e); print(open("gpurun_out/r2_bench_2gpu.err").read()[-1500:])
PY
