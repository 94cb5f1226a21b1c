#!/bin/bash
# round 2, GPU call D: suspend-hint waits, branch-free scan front loop, new bench line
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/r2d_pytest.txt
B="python bench.py --no-cpu --no-e2e --steps 10"
timeout 300 $B 2>$O/r2d_default.err | tail -1 > $O/r2d_default.json
LZ4K_SCAN_IMPL=thread timeout 300 $B 2>$O/r2d_threadscan.err | tail -1 > $O/r2d_threadscan.json
for v in hint200 hint20k rpt2; do
  LZ4_B200_LIBRARY=$PWD/lz4_b200/build/liblz4_b200_$v.so timeout 300 $B 2>$O/r2d_$v.err | tail -1 > $O/r2d_$v.json
done
timeout 300 $B --proba 0.9 2>$O/r2d_p90.err | tail -1 > $O/r2d_p90.json
for f in default threadscan hint200 hint20k rpt2 p90; do
  python - $f <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.load(open('gpurun_out/r2d_%s.json'%f)); r=d['roofline']
    print(f, d['value'],'GB/s  step',d['ms_per_step'],'ms  scan',r['scan_kernel_ms'],' expand',r['kernel_ms'], 'compress', d['compress']['GBps'])
except Exception as e: print(f,'FAILED',e); print(open('gpurun_out/r2d_%s.err'%f).read()[-1500:])
PY
done
LZ4_B200_LIBRARY=$PWD/lz4_b200/build/liblz4_b200_timing.so timeout 200 python tests/perf/phase_timing.py > $O/r2d_phases.txt 2>&1; tail -10 $O/r2d_phases.txt
timeout 600 python bench.py > $O/r2d_full.json 2>$O/r2d_full.err; tail -c 2500 $O/r2d_full.json; tail -3 $O/r2d_full.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > $O/r2d_ref.json 2>$O/r2d_ref.err; tail -c 1500 $O/r2d_ref.json; tail -3 $O/r2d_ref.err
ncu --set full --clock-control none --import-source on -k regex:"expand_rows|scan_kernel" -s 6 -c 2 -f -o $O/prof_r02d \
    python bench.py --gib 0.5 --steps 2 --warmup 3 --no-cpu --no-e2e > $O/ncu_full_r02d.log 2>&1
ls -la $O | tail -3
