#!/bin/bash
# Round-2 GPU call L: final validation of the committed state: GPU test suite, smoke, bench line.
out=gpurun_out
mkdir -p $out
python -c "import __graft_entry__ as g; g.smoke()" > $out/r2l_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $out/r2l_smoke.log
timeout 900 python -m pytest tests -x -q -m gpu > $out/r2l_pytest_default.log 2>&1; echo "pytest[default] exit $?"; tail -3 $out/r2l_pytest_default.log
timeout 600 python bench.py > $out/r2l_bench.json 2> $out/r2l_bench.err; tail -1 $out/r2l_bench.json | cut -c1-600; tail -2 $out/r2l_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/r2l_bench_reference.json 2> $out/r2l_bench_reference.err; tail -1 $out/r2l_bench_reference.json | cut -c1-500
