#!/bin/bash
# Round-2 GPU call E: CTA-pair probe (own process, bounded), builder/contraction overlap probe.
out=gpurun_out
mkdir -p $out
(cd tools/microbench && nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o ozaki_probe.bin ozaki_probe.cu -lcuda \
   && timeout 120 ./ozaki_probe.bin pair) > $out/r2_ozaki_probe_pair.json 2> $out/r2_ozaki_probe_pair.err; echo "pair probe exit $?"; cat $out/r2_ozaki_probe_pair.json; tail -3 $out/r2_ozaki_probe_pair.err
(cd tools/microbench && timeout 120 ./ozaki_probe.bin pair2) > $out/r2_ozaki_probe_pair2.json 2> $out/r2_ozaki_probe_pair2.err; echo "pair2 probe exit $?"; cat $out/r2_ozaki_probe_pair2.json; tail -3 $out/r2_ozaki_probe_pair2.err
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
timeout 300 python tools/overlap_probe.py > $out/r2_overlap_probe.json 2> $out/r2_overlap_probe.err; cat $out/r2_overlap_probe.json; tail -3 $out/r2_overlap_probe.err
