#!/bin/bash
# Round-2 GPU call F: persistent / CTA-pair int8 contraction in the library: parity test of all variants, bench per variant.
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ozaki" > $out/r2f_pytest_ozaki.log 2>&1; echo "pytest[ozaki variants] exit $?"; tail -5 $out/r2f_pytest_ozaki.log
for v in "X=1" "GPK_OZPAIR=1" "GPK_OZPERSIST=0" "GPK_OZPAIR=1 GPK_OZPERSIST=0"; do
  tag=$(echo $v | tr ' =' '__')
  env $v timeout 300 python bench.py --steps 10 --warmup 3 --no-c3 --no-cpu-baseline > $out/r2f_bench_$tag.json 2> $out/r2f_bench_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("$out/r2f_bench_$tag.json").read().strip().splitlines()[-1])
    print("$v", "value", d["value"], "ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], "kernel_ms", d.get("kernel_ms_last_chunk"), "e2e", d["e2e"]["value"], "argmax", d["argmax_check"])
except Exception as e:
    print("$v bench failed", e); print(open("$out/r2f_bench_$tag.err").read()[-1500:])
PY
done
timeout 900 python -m pytest tests -x -q -m gpu > $out/r2f_pytest_default.log 2>&1; echo "pytest[default] exit $?"; tail -3 $out/r2f_pytest_default.log
NCU="ncu --clock-control none"
for k in "oz_persist gpk_oz_persist_kernelILb0E X=1" "oz_nonpersist gpk_oz_vargemm_kernel GPK_OZPERSIST=0"; do
  set -- $k
  env $3 timeout 600 $NCU --set full --import-source on --kernel-name-base mangled -k "regex:$2" -s 2 -c 1 -f -o $out/r02_$1 \
      python tools/profile_driver.py 4096 16 32768 > $out/r02_$1.log 2>&1
  tail -2 $out/r02_$1.log
done
