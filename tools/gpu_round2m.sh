#!/bin/bash
# Round-2 GPU call M: drains of the pair kernel (two levels per TMEM round trip, exact int->fp64 without I2F): parity + bench + role profile.
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ozaki" > $out/r2m_pytest_ozaki.log 2>&1; echo "pytest[ozaki variants] exit $?"; tail -3 $out/r2m_pytest_ozaki.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-c3 --no-cpu-baseline > $out/r2m_bench.json 2> $out/r2m_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2m_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], d["kernel_ms_last_chunk"], d["clocks"])
PY
timeout 300 python tools/oz_profile.py > $out/r2_oz_profile_v3.json 2> $out/r2_oz_profile_v3.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_oz_profile_v3.json"))
for k, v in list(d.items())[:2]:
    print(k, {kk: (round(vv, 1) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("vargemm_ms", "tiles", "issuer_cycles_per_tile", "issuer_wait_operands_per_tile", "issuer_wait_tmem_drain_per_tile")})
PY
