#!/bin/bash
# Round-2 GPU call K: role profile of the default pair kernel; smoke (int8 part); bench launch list under ncu.
out=gpurun_out
mkdir -p $out
python -c "import __graft_entry__ as g; g.smoke()" > $out/r2k_smoke.log 2>&1; echo "smoke exit $?"; tail -3 $out/r2k_smoke.log
timeout 300 python tools/oz_profile.py > $out/r2_oz_profile_v2.json 2> $out/r2_oz_profile_v2.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_oz_profile_v2.json"))
for k, v in d.items():
    print(k, {kk: (round(vv, 1) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("vargemm_ms", "tiles", "issuer_cycles_per_tile", "issuer_wait_operands_per_tile", "issuer_wait_tmem_drain_per_tile", "producer_wait_free_stage_per_tile")})
PY
tail -3 $out/r2_oz_profile_v2.err
timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum -c 3000 --csv --log-file $out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-c3 --no-cpu-baseline > $out/r2k_bench_under_ncu.log 2>&1; tail -1 $out/r2k_bench_under_ncu.log | cut -c1-200
python tools/launch_shares.py $out/r02_launches_bench.csv | head -16 || true
