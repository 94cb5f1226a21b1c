#!/bin/bash
# One gpurun call that measures every experiment prepared in round 1:
#   default build | LZ4K_PHASEB_V2 | LZ4K_SCAN_V2 | both  ->  GPU parity tests + bench line + small-batch lz4bench,
#   then the scan/expand overlap probe and the reference fuzzer on the drop-in entry points.
# Usage:  gpurun --timeout 1500 -- 'bash tests/perf/round2_experiments.sh'
# Results: gpurun_out/exp_<variant>.{json,txt}, gpurun_out/exp_pipeline.json, gpurun_out/exp_fuzzer.txt
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out lz4_b200/build
python - <<'PY'
import sys; sys.path.insert(0, ".")
from oracle.pyoracle import Oracle
Oracle().datagen_mt(256 << 20, 64 << 20, 0.5, 0).tofile("/tmp/p50_256m.bin")
PY
declare -A DEFS=( [default]="" [pbv2]="-DLZ4K_PHASEB_V2" [scanv2]="-DLZ4K_SCAN_V2" [both]="-DLZ4K_PHASEB_V2 -DLZ4K_SCAN_V2" )
for v in default pbv2 scanv2 both; do
  lib=lz4_b200/build/liblz4_b200_$v.so
  [ -f $lib ] || python -m lz4_b200.build --out $lib ${DEFS[$v]} > gpurun_out/exp_$v.build.txt 2>&1 || { echo "$v: build failed"; continue; }
  export LZ4_B200_LIBRARY=$PWD/$lib
  {
    echo "== $v: GPU parity tests"
    timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_frame.py tests/test_lz4bench.py -m gpu -x -q 2>&1 | tail -3
    echo "== $v: lz4bench -b1 -i1 -B4, 256 MiB (small batch)"
    timeout 120 python -m lz4_b200.lz4bench -b1 -i1 -B4 /tmp/p50_256m.bin 2>&1 | tail -1
  } > gpurun_out/exp_$v.txt 2>&1
  timeout 300 python bench.py --no-cpu --no-e2e --steps 10 2> gpurun_out/exp_$v.bench.err | tail -1 > gpurun_out/exp_$v.json
  unset LZ4_B200_LIBRARY
  echo "$v: $(tail -2 gpurun_out/exp_$v.txt | tr '\n' ' ') $(python -c "import json;d=json.load(open('gpurun_out/exp_$v.json'));print(d['value'],'GB/s scan',d['roofline'].get('scan_kernel_ms'),'ms expand',d['roofline'].get('kernel_ms'),'ms')" 2>/dev/null)"
done
timeout 300 python tests/perf/pipeline_probe.py 4 10 > gpurun_out/exp_pipeline.json 2>&1; tail -1 gpurun_out/exp_pipeline.json
timeout 300 oracle/_ref/fuzzer_b200 -s1 -i5 > gpurun_out/exp_fuzzer.txt 2>&1; echo "fuzzer rc=$?"; tail -2 gpurun_out/exp_fuzzer.txt
