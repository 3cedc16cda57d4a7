#!/bin/bash
# Round-2 GPU call G: role-level cycle profile of the persistent int8 contraction; pair two-pass variant in the library.
out=gpurun_out
mkdir -p $out
timeout 300 python tools/oz_profile.py > $out/r2_oz_profile.json 2> $out/r2_oz_profile.err; cat $out/r2_oz_profile.json; tail -3 $out/r2_oz_profile.err
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ozaki" > $out/r2g_pytest_ozaki.log 2>&1; echo "pytest[ozaki variants] exit $?"; tail -5 $out/r2g_pytest_ozaki.log
for v in "GPK_OZPAIR=1 GPK_OZTILE=128 GPK_OZPERSIST=0" "GPK_OZPAIR=1 GPK_OZTILE=128" "GPK_OZPERSIST=2"; do
  tag=$(echo $v | tr ' =' '__')
  env $v timeout 300 python bench.py --steps 10 --warmup 3 --no-c3 --no-cpu-baseline > $out/r2g_bench_$tag.json 2> $out/r2g_bench_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("$out/r2g_bench_$tag.json").read().strip().splitlines()[-1])
    print("$v", "value", d["value"], "ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], "kernel_ms", d.get("kernel_ms_last_chunk"), "e2e", d["e2e"]["value"], "argmax", d["argmax_check"])
except Exception as e:
    print("$v bench failed", e); print(open("$out/r2g_bench_$tag.err").read()[-1500:])
PY
done
