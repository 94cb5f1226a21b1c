#!/bin/bash
# round 2, GPU calls T: the parallel compressor under the sanitizers (tests/perf/enc_determinism.py): plain, memcheck, racecheck
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== plain"; TAG=plain timeout 300 python tests/perf/enc_determinism.py --poison 2>&1 | cut -c1-300
echo "== memcheck"; TAG=memcheck timeout 600 compute-sanitizer --tool memcheck python tests/perf/enc_determinism.py 2>&1 | grep -v "^=========$" | cut -c1-300
echo "== racecheck"; TAG=racecheck timeout 900 compute-sanitizer --tool racecheck python tests/perf/enc_determinism.py 2>&1 | grep -v "^=========$" | cut -c1-300
