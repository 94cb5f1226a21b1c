#!/bin/bash
# round 2, GPU call A: the rows expand kernel -- parity, bench, A/B against the round-1 kernel, wave sizes, scan v2, fuzzer, ncu
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest -m gpu (default library: rows kernel)"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/r2a_pytest.txt
B="python bench.py --no-cpu --no-e2e --steps 10"
timeout 300 $B 2>$O/r2a_rows.err | tail -1 > $O/r2a_rows.json
LZ4K_EXPAND_IMPL=pieces timeout 300 $B 2>$O/r2a_pieces.err | tail -1 > $O/r2a_pieces.json
for v in scanv2 rpt2 rpt8 rpt16; do
  LZ4_B200_LIBRARY=$PWD/lz4_b200/build/liblz4_b200_$v.so timeout 300 $B 2>$O/r2a_$v.err | tail -1 > $O/r2a_$v.json
done
timeout 300 $B --proba 0.9 2>$O/r2a_rows_p90.err | tail -1 > $O/r2a_rows_p90.json
for f in rows pieces scanv2 rpt2 rpt8 rpt16 rows_p90; do
  python - $f <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.load(open('gpurun_out/r2a_%s.json'%f)); r=d['roofline']
    print(f, d['value'],'GB/s  step',d['ms_per_step'],'ms  scan',r['scan_kernel_ms'],' expand',r['kernel_ms'])
except Exception as e: print(f,'FAILED',e)
PY
done
LZ4_B200_LIBRARY=$PWD/lz4_b200/build/liblz4_b200_timing.so timeout 200 python tests/perf/phase_timing.py > $O/r2a_phases.txt 2>&1; tail -12 $O/r2a_phases.txt
timeout 300 oracle/_ref/fuzzer_b200 -s1 -i5 > $O/r2a_fuzzer.txt 2>&1; echo "fuzzer rc=$?"; tail -3 $O/r2a_fuzzer.txt
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r02a.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $O/bench_under_ncu_r02a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:expand_rows -s 3 -c 1 -f -o $O/prof_r02a \
    python bench.py --gib 0.5 --steps 2 --warmup 3 --no-cpu --no-e2e > $O/ncu_full_r02a.log 2>&1
ls -la $O | tail -5
