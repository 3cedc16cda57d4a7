#!/bin/bash
out=gpurun_out
mkdir -p $out
(cd tools/microbench && nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o ozaki_probe.bin ozaki_probe.cu -lcuda \
   && timeout 300 ./ozaki_probe.bin) > $out/r2_ozaki_probe_v4.json 2> $out/r2_ozaki_probe_v4.err; cat $out/r2_ozaki_probe_v4.json; tail -3 $out/r2_ozaki_probe_v4.err
bash tools/profile_r2.sh r02 2>&1 | tail -30
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_small.py > $out/r02_sanitizer_$tool.txt 2>&1
  tail -3 $out/r02_sanitizer_$tool.txt
done
