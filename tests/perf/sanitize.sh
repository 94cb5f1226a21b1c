#!/bin/bash
# compute-sanitizer over the parity tests that exercise every kernel: memcheck + racecheck.
# Usage (under gpurun): bash tests/perf/sanitize.sh TAG      -> gpurun_out/sanitizer_TAG.txt
set -u
cd "$(dirname "$0")/../.."
TAG=${1:-r02}
O=gpurun_out/sanitizer_$TAG.txt
mkdir -p gpurun_out
T="tests/test_gpu_parity.py::test_decode_noisy_source_vs_oracle tests/test_gpu_parity.py::test_mixed_batch_fast_and_slow_lists tests/test_gpu_parity.py::test_compress_random_vs_oracle tests/test_gpu_parity.py::test_decode_overlap_and_long_runs"
: > $O
for tool in memcheck racecheck; do
  echo "===== compute-sanitizer --tool $tool  (python -m pytest $T)" >> $O
  SANITIZE_SMALL=1 timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python -m pytest $T -x -q -m gpu 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error|hazard|=========" | tail -40 >> $O
done
cat $O
