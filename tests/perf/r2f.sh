#!/bin/bash
# round 2, GPU call F: full GPU suite, lean ring scan, encoder phase timing + ncu, racecheck variants
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30 | cut -c1-250 | tee $O/r2f_pytest.txt
B="python bench.py --no-cpu --no-e2e --steps 10"
timeout 300 $B 2>$O/r2f_default.err | tail -1 > $O/r2f_default.json
LZ4K_SCAN_IMPL=thread timeout 300 $B 2>$O/r2f_threadscan.err | tail -1 > $O/r2f_threadscan.json
for f in default threadscan; do
  python - $f <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.load(open('gpurun_out/r2f_%s.json'%f)); r=d['roofline']
    print(f, d['value'],'GB/s  step',d['ms_per_step'],'ms  scan',r['scan_kernel_ms'],' expand',r['kernel_ms'], 'compress', d['compress']['GBps'], 'parallel', d['compress_parallel']['GBps'], d['compress_parallel']['ratio'])
except Exception as e: print(f,'FAILED',e); print(open('gpurun_out/r2f_%s.err'%f).read()[-1500:])
PY
done
LZ4_B200_LIBRARY=$PWD/lz4_b200/build/liblz4_b200_timing.so timeout 200 python tests/perf/enc_timing.py > $O/r2f_enc_phases.txt 2>&1; tail -10 $O/r2f_enc_phases.txt
LZ4_B200_LIBRARY=$PWD/lz4_b200/build/liblz4_b200_timing.so PROBA=0.9 timeout 200 python tests/perf/enc_timing.py 2>&1 | tail -9
ncu --set full --clock-control none --import-source on -k regex:"encode_par|scan_kernel" -s 2 -c 2 -f -o $O/prof_r02f \
    python bench.py --gib 0.5 --steps 2 --warmup 3 --no-cpu --no-e2e > $O/ncu_full_r02f.log 2>&1
bash tests/perf/sanitize.sh r02 > /dev/null 2>&1; cat $O/sanitizer_r02.txt
