#!/bin/bash
# Round-2 GPU call I: integer digit extraction, new defaults (CTA pair x two passes, builder not co-run): tests, bench, launch list, ncu.
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -x -q -m gpu > $out/r2i_pytest_default.log 2>&1; echo "pytest[default] exit $?"; tail -3 $out/r2i_pytest_default.log
timeout 600 python bench.py --steps 10 --warmup 3 > $out/r2i_bench.json 2> $out/r2i_bench.err; tail -1 $out/r2i_bench.json | cut -c1-1000; tail -2 $out/r2i_bench.err
GPK_OZAKI=0 timeout 600 python -m pytest tests -x -q -m gpu > $out/r2i_pytest_fp64.log 2>&1; echo "pytest[GPK_OZAKI=0] exit $?"; tail -2 $out/r2i_pytest_fp64.log
bash tools/profile_r2.sh r02 "oz_pair2 cov_oz" 2>&1 | tail -8
