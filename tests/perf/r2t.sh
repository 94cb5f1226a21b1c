#!/bin/bash
# bisect the parallel compressor's sanitizer-only corruption over its history (current tree, historical lz4_encode_par.cuh)
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for c in branchfree nocoop mailbox; do
  echo "== encoder of $c under memcheck"
  LZ4_B200_LIBRARY=$PWD/lz4_b200/build/liblz4_b200_bis_$c.so TAG=bis timeout 400 compute-sanitizer --tool memcheck python tests/perf/enc_determinism.py 2>&1 | grep -E "run 0 sizes|DOES NOT|DIFFERS|ERROR SUMMARY|Error" | head -8 | cut -c1-300
done
