#!/bin/bash
# Round-2 GPU evidence: launch lists + one `ncu --set full` capture per kernel family the verdict asked for.
# usage (on the GPU box, from the repo root): bash tools/profile_r2.sh [tag]
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
NCU="ncu --clock-control none"
# launch lists (device time per launch; cold-cache, serialised: compare shares)
timeout 600 $NCU --metrics gpu__time_duration.sum -c 1500 --csv --log-file $out/${tag}_launches_fit_score_grad.csv \
    python tools/profile_driver.py 4096 16 32768 > $out/${tag}_driver.log 2>&1
# full captures
cap() {  # name regex skip
  timeout 600 $NCU --set full --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 -f -o $out/${tag}_$1 \
      python tools/profile_driver.py 4096 16 32768 > $out/${tag}_$1.log 2>&1
  tail -2 $out/${tag}_$1.log
}
cap cov_kbuild        'gpk_cov_tma_kernel<8>'   2      # third fit's K build (tri = 1)
cap cov_kstar         'gpk_cov_tma_kernel<4>'   1      # K* of a look-ahead chunk (128 x 16 tiles, next to the GEMM)
cap gemm_trailing_ws  'gpk_gemm_ws_kernel<0>'   70     # a trailing update of the third fit
cap gemm_chain32      'gpk_gemm_nt_kernel<0, 1, 2>' 70 # 32-row panel solve / next-panel update
cap chain_step        'gpk_chain_step_kernel'   70     # X(k): block row k+1 between two diagonal blocks
cap diag_dmma         'gpk_potrf_diag_dmma_kernel' 70
cap vargemm           'gpk_gemm_ws_kernel<1>'   2
cap finish            'gpk_finish_kernel'       2
cap grad_trace        'gpk_grad_trace_kernel'   0
ls -la $out/*.ncu-rep
