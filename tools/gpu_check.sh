#!/bin/bash
# GPU round trip: parity tests, smoke, bench, ncu launch list + one full capture of the top kernel.
# usage: bash tools/gpu_check.sh [tests|bench|ncu|all]
what=${1:-all}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
if [[ $what == tests || $what == all ]]; then
  timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest.log 2>&1
  echo "pytest exit $?" >> gpurun_out/pytest.log
  tail -15 gpurun_out/pytest.log
  timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
fi
if [[ $what == bench || $what == all ]]; then
  timeout 900 python bench.py > gpurun_out/bench.log 2>gpurun_out/bench.err; tail -3 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
fi
if [[ $what == ncu || $what == all ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 3 --m 16384 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:gpk_gemm_ws_kernelILi1E -s 2 -c 1 \
      -f -o gpurun_out/prof_vargemm python bench.py --steps 1 --warmup 3 --m 16384 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
  tail -3 gpurun_out/ncu_full.log
fi
if [[ $what == diag || $what == all ]]; then
  # blocked diagonal-block kernel: launch list of three fits and one full capture (block 5 of the second fit)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_fit.csv \
      python tools/fit_only.py > gpurun_out/ncu_fit_list.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gpk_potrf_diag_dmma_kernel -s 37 -c 1 \
      -f -o gpurun_out/prof_diag_dmma python tools/fit_only.py > gpurun_out/ncu_diag_full.log 2>&1
  tail -3 gpurun_out/ncu_diag_full.log
fi
if [[ $what == sanitize ]]; then
  for tool in memcheck racecheck synccheck; do
    timeout 1200 compute-sanitizer --tool $tool python tools/sanitize_small.py > gpurun_out/sanitizer_$tool.txt 2>&1
    tail -4 gpurun_out/sanitizer_$tool.txt
  done
fi
