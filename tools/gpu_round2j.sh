#!/bin/bash
# Round-2 GPU call J: validation of the final state: fp64-path suite, smoke, sanitizers over every kernel variant, launch lists.
out=gpurun_out
mkdir -p $out
python -c "import __graft_entry__ as g; g.smoke()" > $out/r2j_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $out/r2j_smoke.log
GPK_OZAKI=0 timeout 900 python -m pytest tests -x -q -m gpu > $out/r2j_pytest_fp64.log 2>&1; echo "pytest[GPK_OZAKI=0] exit $?"; tail -2 $out/r2j_pytest_fp64.log
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool python tools/sanitize_small.py > $out/r02_sanitizer_$tool.txt 2>&1
  tail -3 $out/r02_sanitizer_$tool.txt
done
python tools/sass_summary.py > $out/r02_sass_summary.txt 2>&1; tail -3 $out/r02_sass_summary.txt
