#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
echo "== plain"; timeout 300 python tests/perf/enc_determinism.py 2>&1 | tail -8 | cut -c1-400


