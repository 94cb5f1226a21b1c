#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for P in 0.5 0.9; do LZ4_B200_LIBRARY=$PWD/lz4_b200/build/liblz4_b200_timing.so PROBA=$P timeout 200 python tests/perf/enc_timing.py 2>&1 | tail -10; done | tee gpurun_out/r2i_enc_phases.txt
