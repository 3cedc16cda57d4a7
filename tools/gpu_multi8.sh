#!/bin/bash
# 8-GPU (then 4-GPU) torchrun bench lines: weak-scaling headline + configs[2] strong-scaling block with argmax checks.
out=gpurun_out
mkdir -p $out
nvidia-smi -L > $out/r2_multi8_smi.txt 2>&1
for n in 8 4; do
  NCCL_DEBUG=WARN timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
      bench.py --gpus $n --steps 10 --warmup 3 > $out/r2_bench_${n}gpu.json 2> $out/r2_bench_${n}gpu.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_${n}gpu.json").read().strip().splitlines()[-1])
    c3 = d["configs"]["c3"]
    print("$n GPUs: value %.4g ms/step %.2f e2e %.4g argmax_check %s c3 host %.2f ms dev %.2f ms check %s clocks %s" % (
        d["value"], d["ms_per_step"], d["e2e"]["value"], d["argmax_check"], c3["host_pageable"]["wall_ms"], c3["device_philox"]["wall_ms"], c3["argmax_check"], d["clocks"]))
except Exception as e:
    print("$n-GPU bench failed:", e); print(open("gpurun_out/r2_bench_${n}gpu.err").read()[-2000:])
PY
done
