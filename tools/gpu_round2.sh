#!/bin/bash
# One GPU round trip of round 2 (run from the repo root on the GPU box).
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/r2_smi.txt 2>&1
nproc >> $out/r2_smi.txt
run_tests() {  # tag, extra env...
  tag=$1; shift
  env "$@" timeout 1500 python -m pytest tests -m gpu -q --durations=10 -p no:cacheprovider --timeout 900 > $out/r2_pytest_$tag.log 2>&1
  rc=$?
  echo "pytest[$tag] exit $rc" >> $out/r2_pytest_$tag.log
  grep -E "^FAILED|^ERROR|passed|failed|exit" $out/r2_pytest_$tag.log | head -30
  return $rc
}
run_tests default X=1
run_tests fp64 GPK_OZAKI=0
timeout 300 python __graft_entry__.py --smoke > $out/r2_smoke.log 2>&1; tail -2 $out/r2_smoke.log
summ() {
  python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    c3 = (d.get("configs") or {}).get("c3") or {}
    print(sys.argv[2], "value %.4g e2e %.4g ms/step %.3f roofline %.3f (%s) fit_ms %.3f argmax_check %s kernel_ms %s c3 host %.2f ms dev %.2f ms check %s" % (
        d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel"][:28], d["fit_ms"], d.get("argmax_check"),
        d["kernel_ms_last_chunk"], (c3.get("host_pageable") or {}).get("wall_ms", float("nan")),
        (c3.get("device_philox") or {}).get("wall_ms", float("nan")), c3.get("argmax_check")))
except Exception as e:
    print(sys.argv[2], "bench failed:", e)
PY
}
timeout 900 python bench.py --steps 5 --warmup 3 > $out/r2_bench.json 2> $out/r2_bench.err; summ $out/r2_bench.json default; tail -3 $out/r2_bench.err
GPK_OZAKI=0 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $out/r2_bench_fp64.json 2> $out/r2_bench_fp64.err; summ $out/r2_bench_fp64.json fp64; tail -3 $out/r2_bench_fp64.err
GPK_OZFUSED=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-c3 > $out/r2_bench_ozfused.json 2> $out/r2_bench_ozfused.err; summ $out/r2_bench_ozfused.json ozfused
timeout 900 python tools/run_configs.py > $out/r2_configs_c3_c4_c5.jsonl 2> $out/r2_configs.err; cut -c1-600 $out/r2_configs_c3_c4_c5.jsonl; tail -3 $out/r2_configs.err
(cd tools/microbench && nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o ozaki_probe.bin ozaki_probe.cu -lcuda \
   && timeout 300 ./ozaki_probe.bin) > $out/r2_ozaki_probe_v3.json 2> $out/r2_ozaki_probe_v3.err; cat $out/r2_ozaki_probe_v3.json; tail -3 $out/r2_ozaki_probe_v3.err
