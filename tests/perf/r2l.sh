#!/bin/bash
# round 2, GPU call L: device-side frame assembly (tests + frame bench), scan prefetch experiments, split/thread crossover
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -12 | cut -c1-250 | tee $O/r2l_pytest.txt
echo "== frame bench"
for id in 7 4; do timeout 300 python tests/perf/frame_bench.py 1 $id 2>&1 | tail -1; done | tee $O/r2l_frame.jsonl
run() {  # tag lib impl gib
  local tag=$1 lib=$2 impl=$3 gib=$4
  LZ4_B200_LIBRARY=$lib LZ4K_SCAN_IMPL=$impl timeout 300 python bench.py --no-cpu --no-e2e --steps 6 --gib $gib 2>$O/r2l_$tag.err | tail -1 > $O/r2l_$tag.json
  python - $tag <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.load(open('gpurun_out/r2l_%s.json'%f)); r=d['roofline']
    print(f, d['value'],'GB/s  step',d['ms_per_step'],'ms  scan',r['scan_kernel_ms'],' expand',r['kernel_ms'])
except Exception as e: print(f,'FAILED',e); print(open('gpurun_out/r2l_%s.err'%f).read()[-800:])
PY
}
echo "== scan experiments"
D=$PWD/lz4_b200/liblz4_b200.so
V=$PWD/lz4_b200/build
{
run base_thread_4 $D thread 4
for g in 0.5 1 2; do run base_thread_$g $D thread $g; run base_split_$g $D split $g; done
for v in pfA pfB pfC pfE pfF; do run ${v}_thread_4 $V/liblz4_b200_$v.so thread 4; done
run pfA_split_4 $V/liblz4_b200_pfA.so split 4
run pfD_split_4 $V/liblz4_b200_pfD.so split 4
run pfA_split_0.5 $V/liblz4_b200_pfA.so split 0.5
} | tee $O/r2l_scan.txt
