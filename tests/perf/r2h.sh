#!/bin/bash
# round 2, GPU call H: parallel compressor v2 (16 K windows, per-selection extension, all-lane chain rounds)
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30 | cut -c1-250 | tee $O/r2h_pytest.txt
timeout 300 python bench.py --no-cpu --no-e2e --steps 10 2>$O/r2h_default.err | tail -1 > $O/r2h_default.json
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2h_default.json')); r=d['roofline']
    print(d['value'],'GB/s  step',d['ms_per_step'],'ms  scan',r['scan_kernel_ms'],' expand',r['kernel_ms'], 'compress', d['compress']['GBps'], 'parallel', d['compress_parallel']['GBps'], d['compress_parallel']['ratio'], d['compress_parallel']['ratio_vs_reference'])
except Exception as e: print('FAILED',e); print(open('gpurun_out/r2h_default.err').read()[-1500:])
PY
for P in 0.5 0.9 0.2; do LZ4_B200_LIBRARY=$PWD/lz4_b200/build/liblz4_b200_timing.so PROBA=$P timeout 200 python tests/perf/enc_timing.py 2>&1 | tail -9; done | tee $O/r2h_enc_phases.txt
ncu --set full --clock-control none --import-source on -k regex:"encode_par" -s 1 -c 1 -f -o $O/prof_r02h \
    python bench.py --gib 0.5 --steps 2 --warmup 3 --no-cpu --no-e2e > $O/ncu_full_r02h.log 2>&1
ls -la $O | tail -2
