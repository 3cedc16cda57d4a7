#!/bin/bash
# 2-GPU round trip: the NCCL arg-max exchange behind the C ABI (tests/test_gpu_multi.py) and the torchrun bench line.
out=gpurun_out
mkdir -p $out
nvidia-smi -L > $out/r2_multi_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider --timeout 600 > $out/r2_pytest_multi.log 2>&1
echo "pytest[multi] exit $?" >> $out/r2_pytest_multi.log
grep -E "^FAILED|^ERROR|passed|failed|skipped|exit|Error" $out/r2_pytest_multi.log | head -20
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 5 --warmup 3 > $out/r2_bench_2gpu.json 2> $out/r2_bench_2gpu.err
tail -c 1500 $out/r2_bench_2gpu.json; tail -5 $out/r2_bench_2gpu.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_2gpu.json").read().strip().splitlines()[-1])
    c3 = d["configs"]["c3"]
    print("2 GPUs: value %.4g e2e %.4g argmax_check %s c3 host %.2f ms dev %.2f ms check %s" % (
        d["value"], d["e2e"]["value"], d["argmax_check"], c3["host_pageable"]["wall_ms"], c3["device_philox"]["wall_ms"], c3["argmax_check"]))
except Exception as e:
    print("2-GPU bench failed:", e)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $out/r2_bench_2gpu_reference.json 2> $out/r2_bench_2gpu_reference.err
tail -c 400 $out/r2_bench_2gpu_reference.json
