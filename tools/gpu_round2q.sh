#!/bin/bash
# Round-2 GPU call Q: does the persistent tile walk pay at N = 1024 (configs[2]), where tiles are short?
out=gpurun_out
mkdir -p $out
for v in "X=1" "GPK_OZPERSIST=1"; do
  tag=$(echo $v | tr ' =' '__')
  env $v timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/r2q_bench_$tag.json 2> $out/r2q_bench_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("$out/r2q_bench_$tag.json").read().strip().splitlines()[-1])
    c = d["configs"]["c3"]
    print("$v", "value", d["value"], "c3 host", c["host_pageable"]["wall_ms"], "dev", c["device_philox"]["wall_ms"], "score", c["device_philox"]["rank0_score_ms_last_call"], "check", c["argmax_check"], d["config"]["l2"][:80])
except Exception as e:
    print("$v failed", e); print(open("$out/r2q_bench_$tag.err").read()[-800:])
PY
done
