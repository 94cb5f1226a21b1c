#!/bin/bash
# f-2: the GPU `lz4 -b` harness next to the reference tool on the same file (run under gpurun).
#   bash tests/perf/lz4bench_run.sh [MiB] [seconds]
set -u
cd "$(dirname "$0")/../.."
MIB=${1:-256}; SEC=${2:-1}
mkdir -p gpurun_out
python - <<PY
import sys; sys.path.insert(0, ".")
from oracle.pyoracle import Oracle
Oracle().datagen_mt($MIB << 20, 64 << 20, 0.5, 0).tofile("/tmp/p50_${MIB}m.bin")
PY
{
  echo "# reference tool (1 host thread): lz4 -b1 -i$SEC -B4"
  oracle/_ref/lz4 -b1 -i$SEC -B4 /tmp/p50_${MIB}m.bin 2>&1 | tr '\r' '\n' | grep 'MB/s,' | tail -1
  echo "# lz4_b200.lz4bench (1 B200, device resident): -b1 -i$SEC -B4"
  python -m lz4_b200.lz4bench -b1 -i$SEC -B4 /tmp/p50_${MIB}m.bin 2>&1 | tail -2
  echo "# lz4_b200.lz4bench --fast=8"
  python -m lz4_b200.lz4bench --fast=8 -b -i$SEC -B4 /tmp/p50_${MIB}m.bin 2>&1 | tail -2
} | tee gpurun_out/lz4bench_r01.txt
