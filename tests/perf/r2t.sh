#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
echo "== plain"; timeout 300 python tests/perf/enc_determinism.py 2>&1 | tail -8 | cut -c1-400
echo "== memcheck"; timeout 600 compute-sanitizer --tool memcheck python tests/perf/enc_determinism.py 2>&1 | grep -v "^=========$" | tail -12 | cut -c1-400
echo "== synccheck"; timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python tests/perf/enc_determinism.py 2>&1 | tail -30 | cut -c1-300
