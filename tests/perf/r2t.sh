#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
echo "== memcheck"; timeout 600 compute-sanitizer --tool memcheck python tests/perf/enc_determinism.py 2>&1 | grep -v "^=========$" | tail -16 | cut -c1-600
echo "== plain"; timeout 300 python tests/perf/enc_determinism.py 2>&1 | tail -8 | cut -c1-600
