#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== memcheck"; TAG=memcheck timeout 600 compute-sanitizer --tool memcheck python tests/perf/enc_determinism.py 2>&1 | grep -v "^=========$" | head -14 | cut -c1-400
echo "== plain"; TAG=plain timeout 300 python tests/perf/enc_determinism.py 2>&1 | head -12 | cut -c1-400
echo "== racecheck"; TAG=racecheck timeout 600 compute-sanitizer --tool racecheck python tests/perf/enc_determinism.py 2>&1 | grep -v "^=========$" | head -14 | cut -c1-400
