#!/bin/bash
# Round-2 GPU call P: evidence refresh on the final code: sanitizers over every kernel variant, ncu of the default scoring kernel
# and the builder, launch list of the bench.
out=gpurun_out
mkdir -p $out
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool python tools/sanitize_small.py > $out/r02_sanitizer_$tool.txt 2>&1
  tail -2 $out/r02_sanitizer_$tool.txt
done
bash tools/profile_r2.sh r02 "oz_pair2 cov_oz" 2>&1 | tail -6
timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum -c 3000 --csv --log-file $out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-c3 --no-cpu-baseline > $out/r2p_bench_under_ncu.log 2>&1
python tools/launch_shares.py $out/r02_launches_bench.csv 2>/dev/null | head -8
