#!/bin/bash
# round 2, GPU call Y: the straight-line cooperative finish (-DLZ4K_COOP2) under the tools, its tests and its speed
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export LZ4_B200_LIBRARY=$PWD/lz4_b200/build/liblz4_b200_coop2.so
echo "== memcheck"; TAG=coop2m timeout 600 compute-sanitizer --tool memcheck python tests/perf/enc_determinism.py 2>&1 | grep -E "run 0 sizes|DOES NOT|DIFFERS|BAD|ERROR SUMMARY" | cut -c1-300
echo "== racecheck"; TAG=coop2r timeout 900 compute-sanitizer --tool racecheck python tests/perf/enc_determinism.py 2>&1 | grep -E "run 0 sizes|DOES NOT|DIFFERS|BAD|RACECHECK SUMMARY" | cut -c1-300
echo "== plain"; TAG=coop2p timeout 300 python tests/perf/enc_determinism.py 2>&1 | grep -E "run 0 sizes|DOES NOT|DIFFERS|BAD" | cut -c1-300
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_parallel_compress.py -q -m gpu 2>&1 | tail -3 | cut -c1-200
for P in 0.5 0.9; do timeout 300 python bench.py --no-cpu --no-e2e --steps 4 --proba $P 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.load(sys.stdin); print('proba', $P, 'parallel', d['compress_parallel']['GBps'], d['compress_parallel']['ratio'], d['compress_parallel']['ratio_vs_reference'])"; done
