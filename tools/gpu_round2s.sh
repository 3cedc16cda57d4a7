#!/bin/bash
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -x -q -m gpu > $out/r2s_pytest_default.log 2>&1; echo "pytest[default] exit $?"; tail -3 $out/r2s_pytest_default.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/r2s_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $out/r2s_smoke.log
