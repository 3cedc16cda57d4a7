#!/bin/bash
# Round-2 GPU call D: balanced base-256 digits (7 slices, 28 pairs) in probe + library: tests, bench (both tiles).
out=gpurun_out
mkdir -p $out
(cd tools/microbench && nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o ozaki_probe.bin ozaki_probe.cu -lcuda \
   && timeout 300 ./ozaki_probe.bin) > $out/r2_ozaki_probe_v6.json 2> $out/r2_ozaki_probe_v6.err; cat $out/r2_ozaki_probe_v6.json; tail -3 $out/r2_ozaki_probe_v6.err
timeout 900 python -m pytest tests -x -q -m gpu > $out/r2d_pytest_default.log 2>&1; echo "pytest[default] exit $?"; tail -3 $out/r2d_pytest_default.log
timeout 600 python bench.py --steps 10 --warmup 3 > $out/r2d_bench.json 2> $out/r2d_bench.err; tail -1 $out/r2d_bench.json | cut -c1-1200
GPK_OZTILE=128 timeout 600 python bench.py --steps 10 --warmup 3 --no-c3 > $out/r2d_bench_oztile128.json 2> $out/r2d_bench_oztile128.err; tail -1 $out/r2d_bench_oztile128.json | cut -c1-400
