#!/bin/bash
# Round-2 GPU call C: grouped tile order of the int8 contraction (probe + library), tests, bench, missing ncu captures.
out=gpurun_out
mkdir -p $out
(cd tools/microbench && nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o ozaki_probe.bin ozaki_probe.cu -lcuda \
   && timeout 300 ./ozaki_probe.bin) > $out/r2_ozaki_probe_v5.json 2> $out/r2_ozaki_probe_v5.err; cat $out/r2_ozaki_probe_v5.json; tail -3 $out/r2_ozaki_probe_v5.err
timeout 900 python -m pytest tests -x -q -m gpu > $out/r2c_pytest_default.log 2>&1; echo "pytest[default] exit $?"; tail -3 $out/r2c_pytest_default.log
timeout 600 python bench.py --steps 10 --warmup 3 > $out/r2c_bench.json 2> $out/r2c_bench.err; tail -1 $out/r2c_bench.json | cut -c1-1500
GPK_OZTILE=128 timeout 600 python bench.py --steps 10 --warmup 3 --no-c3 > $out/r2c_bench_oztile128.json 2> $out/r2c_bench_oztile128.err; tail -1 $out/r2c_bench_oztile128.json | cut -c1-600
bash tools/profile_r2.sh r02 "oz_vargemm cov_oz cov_kbuild cov_kstar_fp64 vargemm_fp64 gemm_trailing_ws gemm_chain32" 2>&1 | tail -30
