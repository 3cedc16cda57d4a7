#!/bin/bash
# Round-2 GPU call O: sustained int8 issue rate with random operands (roofline denominator); bench line.
out=gpurun_out
mkdir -p $out
python - <<'PY'
import sys
sys.path.insert(0, ".")
from robo_b200 import _lib
h = _lib.Handle(0)
print("burst", h.measure_int8_peak())
for secs in (0.5, 2.0):
    print("sustained %.1fs constant operands" % secs, h.measure_int8_peak_sustained(secs, False))
    print("sustained %.1fs random operands" % secs, h.measure_int8_peak_sustained(secs, True))
h.close()
PY
timeout 600 python bench.py > $out/r2o_bench.json 2> $out/r2o_bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2o_bench.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("value", d["value"], "ms/step", d["ms_per_step"], "achieved", r["achieved"], "peak", r["peak"], "frac", r["frac"], "burst", r["peak_burst"], "const", r["peak_sustained_constant_operands"], d["clocks"])
PY
tail -2 $out/r2o_bench.err
