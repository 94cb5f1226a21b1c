#!/bin/bash
# round 2, GPU call J: split scan (4 merging lanes per block), encoder with cooperative long-match extension
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -12 | cut -c1-250 | tee $O/r2j_pytest.txt
B="python bench.py --no-cpu --no-e2e --steps 10"
timeout 300 $B 2>$O/r2j_default.err | tail -1 > $O/r2j_default.json
LZ4K_SCAN_IMPL=thread timeout 300 $B 2>$O/r2j_threadscan.err | tail -1 > $O/r2j_threadscan.json
timeout 300 $B --proba 0.9 2>$O/r2j_p90.err | tail -1 > $O/r2j_p90.json
for f in default threadscan p90; do
  python - $f <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.load(open('gpurun_out/r2j_%s.json'%f)); r=d['roofline']
    print(f, d['value'],'GB/s  step',d['ms_per_step'],'ms  scan',r['scan_kernel_ms'],' expand',r['kernel_ms'], 'compress', d['compress']['GBps'], 'parallel', d['compress_parallel']['GBps'], d['compress_parallel']['ratio'], d['compress_parallel']['ratio_vs_reference'])
except Exception as e: print(f,'FAILED',e); print(open('gpurun_out/r2j_%s.err'%f).read()[-1500:])
PY
done
for P in 0.5 0.9; do LZ4_B200_LIBRARY=$PWD/lz4_b200/build/liblz4_b200_timing.so PROBA=$P timeout 200 python tests/perf/enc_timing.py 2>&1 | tail -10; done | tee $O/r2j_enc_phases.txt
timeout 200 python -m lz4_b200.lz4bench -b1 -i1 -B4 /dev/null > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"scan_split|encode_par" -s 2 -c 2 -f -o $O/prof_r02j \
    python bench.py --gib 0.5 --steps 2 --warmup 3 --no-cpu --no-e2e > $O/ncu_full_r02j.log 2>&1
ls -la $O | tail -2
