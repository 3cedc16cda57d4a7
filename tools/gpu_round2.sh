#!/bin/bash
# One GPU round trip of round 2 (run from the repo root on the GPU box): parity suite with the default kernels, per-switch
# reruns if anything fails (to tell WHICH new kernel is at fault), bench, fit comparison with cuSOLVER.
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/r2_smi.txt 2>&1
nproc >> $out/r2_smi.txt
run_tests() {  # tag, extra env...
  tag=$1; shift
  env "$@" timeout 1500 python -m pytest tests -m gpu -q --durations=10 -p no:cacheprovider --timeout 900 > $out/r2_pytest_$tag.log 2>&1
  rc=$?
  echo "pytest[$tag] exit $rc" >> $out/r2_pytest_$tag.log
  tail -25 $out/r2_pytest_$tag.log
  return $rc
}
if ! run_tests default; then
  # which switch? (the exact-config file is skipped in the reruns: minutes of CPU oracle each time)
  for sw in "GPK_COV=1" "GPK_PERSIST=0" "GPK_CHAINSPLIT=0" "GPK_GRAPH=0" "GPK_COV=1 GPK_PERSIST=0 GPK_CHAINSPLIT=0"; do
    tag=$(echo $sw | tr ' =' '__')
    env $sw timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider --timeout 600 > $out/r2_pytest_$tag.log 2>&1
    echo "pytest[$sw] exit $?" >> $out/r2_pytest_$tag.log
    tail -6 $out/r2_pytest_$tag.log
  done
fi
timeout 300 python __graft_entry__.py --smoke > $out/r2_smoke.log 2>&1; tail -2 $out/r2_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > $out/r2_bench.json 2> $out/r2_bench.err; tail -c 3000 $out/r2_bench.json; tail -5 $out/r2_bench.err
for sw in "GPK_PERSIST=0" "GPK_COV=1"; do
  tag=$(echo $sw | tr ' =' '__')
  env $sw timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-c3 > $out/r2_bench_$tag.json 2> $out/r2_bench_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("$out/r2_bench_$tag.json").read().strip().splitlines()[-1])
    print("$sw", "value", d["value"], "ms/step", d["ms_per_step"], "roofline", d["roofline"]["frac"], "kstar", d["kernel_ms_last_chunk"])
except Exception as e:
    print("$sw bench failed", e)
PY
done
timeout 600 python tools/fit_compare.py > $out/r2_fit_compare.jsonl 2> $out/r2_fit_compare.err; cat $out/r2_fit_compare.jsonl; tail -3 $out/r2_fit_compare.err
# int8 tcgen05 probe (VERDICT item 8 groundwork): descriptors checked against a CPU integer GEMM + issue rate
(cd tools/microbench && nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o i8_umma_probe.bin i8_umma_probe.cu -lcuda \
   && timeout 120 ./i8_umma_probe.bin) > $out/r2_i8_probe.json 2> $out/r2_i8_probe.err; cat $out/r2_i8_probe.json; tail -3 $out/r2_i8_probe.err
# Ozaki prototype: fp64 variance contraction on the int8 tensor pipe (accuracy vs an 80-bit CPU reference + C2-chunk timing)
(cd tools/microbench && nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o ozaki_probe.bin ozaki_probe.cu -lcuda \
   && timeout 300 ./ozaki_probe.bin) > $out/r2_ozaki_probe.json 2> $out/r2_ozaki_probe.err; cat $out/r2_ozaki_probe.json; tail -3 $out/r2_ozaki_probe.err
# the other BASELINE configs (C3 host-fed on one GPU, C4 likelihood pool + fused marginalised EI, C5 nll + gradient)
timeout 900 python tools/run_configs.py > $out/r2_configs_c3_c4_c5.jsonl 2> $out/r2_configs.err; cat $out/r2_configs_c3_c4_c5.jsonl; tail -3 $out/r2_configs.err
