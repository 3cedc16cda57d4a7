#!/bin/bash
out=gpurun_out
mkdir -p $out
timeout 300 python tools/persist_threshold.py > $out/r2_persist_threshold.json 2> $out/r2_persist_threshold.err; cat $out/r2_persist_threshold.json | tr -d '\n' | sed 's/  */ /g'; echo; tail -2 $out/r2_persist_threshold.err
