#!/bin/bash
# Round-2 GPU evidence: launch lists + one `ncu --set full` capture per kernel family the verdict asked for.
# usage (on the GPU box, from the repo root): bash tools/profile_r2.sh [tag] [only: space-separated capture names]
# Kernel filters use MANGLED names: the demangled template arguments print as <(int)0>, which a plain regex misses.
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
NCU="ncu --clock-control none"
# launch lists (device time per launch; cold-cache, serialised: compare shares)
timeout 600 $NCU --metrics gpu__time_duration.sum -c 1500 --csv --log-file $out/${tag}_launches_fit_score_grad.csv \
    python tools/profile_driver.py 4096 16 32768 > $out/${tag}_driver.log 2>&1
GPK_OZAKI=0 timeout 600 $NCU --metrics gpu__time_duration.sum -c 1500 --csv --log-file $out/${tag}_launches_fit_score_grad_fp64.csv \
    python tools/profile_driver.py 4096 16 32768 > $out/${tag}_driver_fp64.log 2>&1
only=" ${2:-} "
cap() {  # name regex skip [env]
  if [ "$only" != "  " ] && [[ "$only" != *" $1 "* ]]; then return; fi
  env $4 timeout 600 $NCU --set full --import-source on --kernel-name-base mangled -k "regex:$2" -s $3 -c 1 -f -o $out/${tag}_$1 \
      python tools/profile_driver.py 4096 16 32768 > $out/${tag}_$1.log 2>&1
  tail -2 $out/${tag}_$1.log
}
cap oz_pair2          'gpk_oz_pair2_kernel'     2      X=1           # int8 variance contraction, CTA pair x two passes (default scoring kernel)
cap oz_vargemm        'gpk_oz_vargemm_kernel'   2      'GPK_OZPAIR=0 GPK_OZTILE=64'   # one-pass single-CTA int8 contraction
cap cov_oz            'gpk_cov_oz_kernelILi4E'  1      X=1           # fused covariance builder + int8 digits (look-ahead chunk)
cap cov_kbuild        'gpk_cov_tma_kernelILi8E' 2      X=1           # third fit's K build (tri = 1)
cap cov_kstar_fp64    'gpk_cov_tma_kernelILi4E' 1      GPK_OZAKI=0   # fp64 K* of a look-ahead chunk
cap vargemm_fp64      'gpk_gemm_ws_kernelILi1E' 2      GPK_OZAKI=0   # fp64 DMMA variance contraction
cap gemm_trailing_ws  'gpk_gemm_ws_kernelILi0E' 70     X=1           # a trailing update of the third fit
cap gemm_chain32      'gpk_gemm_nt_kernelILi0ELi1ELi2E' 70 X=1           # 32-row panel solve / next-panel update
cap diag_dmma         'gpk_potrf_diag_dmma_kernel' 70  X=1
cap finish            'gpk_finish_kernel'       2      X=1
cap grad_trace        'gpk_grad_trace_kernel'   0      X=1
ls -la $out/*.ncu-rep
