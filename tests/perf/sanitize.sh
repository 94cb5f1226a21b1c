#!/bin/bash
# compute-sanitizer over the parity tests that exercise every kernel: memcheck + racecheck.
# Usage (under gpurun): bash tests/perf/sanitize.sh TAG      -> gpurun_out/sanitizer_TAG.txt
set -u
cd "$(dirname "$0")/../.."
TAG=${1:-r02}
O=gpurun_out/sanitizer_$TAG.txt
mkdir -p gpurun_out
T="tests/test_gpu_parity.py::test_decode_noisy_source_vs_oracle tests/test_gpu_parity.py::test_mixed_batch_fast_and_slow_lists tests/test_gpu_parity.py::test_compress_random_vs_oracle tests/test_gpu_parity.py::test_decode_overlap_and_long_runs tests/test_gpu_parallel_compress.py::test_special_shapes_long_runs_and_long_literals tests/test_gpu_parallel_compress.py::test_limited_output_and_never_past_capacity tests/test_frame.py::test_gpu_frames_byte_identical_and_roundtrip"
: > $O
run() {   # run <title> <tool> [env...]
  echo "===== $1: compute-sanitizer --tool $2  (python -m pytest <7 parity tests> -m gpu)" >> $O
  shift; tool=$1; shift
  env "$@" timeout 1500 compute-sanitizer --tool $tool --print-limit 12 python -m pytest $T -x -q -m gpu 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error|hazard" | cut -c1-200 | tail -16 >> $O
}
run "shipped kernels" memcheck X=1
run "shipped kernels" racecheck X=1
# racecheck does not model mbarrier arrive/wait (the wave barrier of the rows kernel) as synchronisation;
# the same kernel with a CTA barrier between the waves must be hazard free:
if [ -f lz4_b200/build/liblz4_b200_barsync.so ]; then
  run "debug build: CTA barrier between waves (-DLZ4K_WAVE_BARSYNC)" racecheck LZ4_B200_LIBRARY=$PWD/lz4_b200/build/liblz4_b200_barsync.so
fi
cat $O
