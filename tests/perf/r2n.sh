#!/bin/bash
# round 2, GPU call N: encoder with two CTAs per SM and one two-halves table (native atomicMax)
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
T=${1:-r2n}
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -12 | cut -c1-250 | tee $O/${T}_pytest.txt
run() {  # tag extra
  local tag=$1; shift
  timeout 300 python bench.py --no-cpu --no-e2e --steps 6 "$@" 2>$O/${T}_$tag.err | tail -1 > $O/${T}_$tag.json
  python - $O/${T}_$tag <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.load(open(f+'.json')); r=d['roofline']
    print(f, d['value'],'GB/s  step',d['ms_per_step'],'ms  scan',r['scan_kernel_ms'],' expand',r['kernel_ms'], 'compress', d['compress']['GBps'], 'parallel', d['compress_parallel']['GBps'], d['compress_parallel']['ratio'], d['compress_parallel']['ratio_vs_reference'])
except Exception as e: print(f,'FAILED',e); print(open(f+'.err').read()[-800:])
PY
}
{
run p50
run p90 --proba 0.9
run p20 --proba 0.2
} | tee $O/${T}_bench.txt
for P in 0.5 0.9; do LZ4_B200_LIBRARY=$PWD/lz4_b200/build/liblz4_b200_timing.so PROBA=$P timeout 200 python tests/perf/enc_timing.py 2>&1 | tail -10; done | tee $O/${T}_enc_phases.txt
ncu --set full --clock-control none --import-source on -k regex:"encode_par" -s 1 -c 1 -f -o $O/prof_$T \
    python bench.py --gib 0.5 --steps 2 --warmup 3 --no-cpu --no-e2e > $O/ncu_full_$T.log 2>&1
ls -la $O | tail -2
