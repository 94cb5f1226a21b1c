#!/bin/bash
# round 2, GPU call C: merged hop loop, non-blocking descriptor prefetch, ring scan
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/r2c_pytest.txt
B="python bench.py --no-cpu --no-e2e --steps 10"
timeout 300 $B 2>$O/r2c_default.err | tail -1 > $O/r2c_default.json
LZ4K_SCAN_IMPL=thread timeout 300 $B 2>$O/r2c_threadscan.err | tail -1 > $O/r2c_threadscan.json
for v in ring512 ring128 rpt8; do
  LZ4_B200_LIBRARY=$PWD/lz4_b200/build/liblz4_b200_$v.so timeout 300 $B 2>$O/r2c_$v.err | tail -1 > $O/r2c_$v.json
done
timeout 300 $B --proba 0.9 2>$O/r2c_p90.err | tail -1 > $O/r2c_p90.json
timeout 300 $B --proba 0.2 2>$O/r2c_p20.err | tail -1 > $O/r2c_p20.json
for f in default threadscan ring512 ring128 rpt8 p90 p20; do
  python - $f <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.load(open('gpurun_out/r2c_%s.json'%f)); r=d['roofline']
    print(f, d['value'],'GB/s  step',d['ms_per_step'],'ms  scan',r['scan_kernel_ms'],' expand',r['kernel_ms'])
except Exception as e: print(f,'FAILED',e); print(open('gpurun_out/r2c_%s.err'%f).read()[-1500:])
PY
done
LZ4_B200_LIBRARY=$PWD/lz4_b200/build/liblz4_b200_timing.so timeout 200 python tests/perf/phase_timing.py > $O/r2c_phases.txt 2>&1; tail -10 $O/r2c_phases.txt
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r02c.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $O/bench_under_ncu_r02c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"expand_rows|scan_kernel" -s 6 -c 2 -f -o $O/prof_r02c \
    python bench.py --gib 0.5 --steps 2 --warmup 3 --no-cpu --no-e2e > $O/ncu_full_r02c.log 2>&1
ls -la $O | tail -3
