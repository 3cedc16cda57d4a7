#!/bin/bash
# Round-2 GPU call H: single-stream PDL pipeline (resident builder + dependent contraction): parity of all variants, bench per mode.
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ozaki" > $out/r2h_pytest_ozaki.log 2>&1; echo "pytest[ozaki variants] exit $?"; tail -5 $out/r2h_pytest_ozaki.log
for v in "X=1" "GPK_OZPDL=0" "GPK_COVCTAS=1" "GPK_COVCTAS=3" "GPK_OZPAIR=1 GPK_OZTILE=128" "GPK_OZPAIR=1 GPK_OZTILE=128 GPK_OZPDL=0" "GPK_OZPAIR=1 GPK_OZTILE=128 GPK_OZPERSIST=1"; do
  tag=$(echo $v | tr ' =' '__')
  env $v timeout 300 python bench.py --steps 10 --warmup 3 --no-c3 --no-cpu-baseline > $out/r2h_bench_$tag.json 2> $out/r2h_bench_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("$out/r2h_bench_$tag.json").read().strip().splitlines()[-1])
    print("$v", "value", d["value"], "ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], "kernel_ms", d.get("kernel_ms_last_chunk"), "e2e", d["e2e"]["value"], "argmax", d["argmax_check"])
except Exception as e:
    print("$v bench failed", e); print(open("$out/r2h_bench_$tag.err").read()[-1500:])
PY
done
timeout 900 python -m pytest tests -x -q -m gpu > $out/r2h_pytest_default.log 2>&1; echo "pytest[default] exit $?"; tail -3 $out/r2h_pytest_default.log
