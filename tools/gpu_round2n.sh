#!/bin/bash
# Round-2 GPU call N: new int8 edge-shape tests.
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "int8 or ozaki" > $out/r2n_pytest_int8.log 2>&1; echo "pytest[int8] exit $?"; tail -15 $out/r2n_pytest_int8.log
