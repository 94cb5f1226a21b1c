#!/bin/bash
# round 2, GPU call M: encoder with two CTAs per SM (16-bit tables), scan experiments (cache-global stores, fewer blocks per warp)
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -12 | cut -c1-250 | tee $O/r2m_pytest.txt
run() {  # tag lib impl gib extra
  local tag=$1 lib=$2 impl=$3 gib=$4; shift 4
  LZ4_B200_LIBRARY=$lib LZ4K_SCAN_IMPL=$impl timeout 300 python bench.py --no-cpu --no-e2e --steps 6 --gib $gib "$@" 2>$O/r2m_$tag.err | tail -1 > $O/r2m_$tag.json
  python - $tag <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.load(open('gpurun_out/r2m_%s.json'%f)); r=d['roofline']
    print(f, d['value'],'GB/s  step',d['ms_per_step'],'ms  scan',r['scan_kernel_ms'],' expand',r['kernel_ms'], 'compress', d['compress']['GBps'], 'parallel', d['compress_parallel']['GBps'], d['compress_parallel']['ratio'], d['compress_parallel']['ratio_vs_reference'])
except Exception as e: print(f,'FAILED',e); print(open('gpurun_out/r2m_%s.err'%f).read()[-800:])
PY
}
D=$PWD/lz4_b200/liblz4_b200.so
V=$PWD/lz4_b200/build
{
run base_p50 $D thread 4
run base_p90 $D thread 4 --proba 0.9
run base_p20 $D thread 4 --proba 0.2
for v in cgA ls2 ls4 ls4cg; do run ${v}_thread_4 $V/liblz4_b200_$v.so thread 4; done
run cgA_split_4 $V/liblz4_b200_cgA.so split 4
run cgB_split_4 $V/liblz4_b200_cgB.so split 4
run cgA_split_1 $V/liblz4_b200_cgA.so split 1
} | tee $O/r2m_scan.txt
for P in 0.5 0.9; do LZ4_B200_LIBRARY=$V/liblz4_b200_timing.so PROBA=$P timeout 200 python tests/perf/enc_timing.py 2>&1 | tail -10; done | tee $O/r2m_enc_phases.txt
ncu --set full --clock-control none --import-source on -k regex:"encode_par" -s 1 -c 1 -f -o $O/prof_r02m \
    python bench.py --gib 0.5 --steps 2 --warmup 3 --no-cpu --no-e2e > $O/ncu_full_r02m.log 2>&1
ls -la $O | tail -2
