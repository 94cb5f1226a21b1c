#!/bin/bash
# compute-sanitizer over the parity tests that exercise every kernel: memcheck + racecheck.
# Usage (under gpurun): bash tests/perf/sanitize.sh TAG      -> gpurun_out/sanitizer_TAG.txt
set -u
cd "$(dirname "$0")/../.."
TAG=${1:-r02}
O=gpurun_out/sanitizer_$TAG.txt
mkdir -p gpurun_out
T="tests/test_gpu_parity.py::test_decode_noisy_source_vs_oracle tests/test_gpu_parity.py::test_mixed_batch_fast_and_slow_lists tests/test_gpu_parity.py::test_compress_random_vs_oracle tests/test_gpu_parity.py::test_decode_overlap_and_long_runs tests/test_gpu_parity.py::test_big_blocks_in_tiles_vs_oracle tests/test_gpu_parallel_compress.py::test_special_shapes_long_runs_and_long_literals tests/test_gpu_parallel_compress.py::test_limited_output_and_never_past_capacity tests/test_frame.py::test_gpu_frames_byte_identical_and_roundtrip"
: > $O
run() {   # run <title> <tool> [env...]
  local title=$1 tool=$2; shift 2
  echo "===== $title: compute-sanitizer --tool $tool  (python -m pytest <8 parity tests> -m gpu)" >> $O
  env "$@" timeout 2400 compute-sanitizer --tool $tool --print-limit 10000 python -m pytest $T -x -q -m gpu > gpurun_out/sanitizer_raw.txt 2>&1
  grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/sanitizer_raw.txt | cut -c1-200 >> $O
  python - >> $O <<'PY'
import re, collections
txt = open('gpurun_out/sanitizer_raw.txt').read()
kinds = collections.Counter()
for blk in re.split(r'=========\s*\n', txt):
    m = re.search(r'(Error|Warning): (Race reported between|.*hazard).*', blk)
    if not m: continue
    sev = m.group(1)
    locs = sorted(set(re.findall(r'in (lz4_[a-z_]+\.(?:cu|cuh|h)):\d+', blk)))
    fns = sorted(set(re.findall(r'(lz4_[a-z_]+_kernel|Table<[^>]*>::\w+|sts_u8|lds_u8|lds_u32|lds_u64)', blk)))
    kinds[(sev, ', '.join(fns) or ', '.join(locs) or '?')] += 1
for (sev, what), n in sorted(kinds.items()):
    print("   %-7s x %-4d %s" % (sev, n, what))
PY
}
run "shipped kernels" memcheck X=1
run "shipped kernels" racecheck X=1
# racecheck does not model mbarrier arrive/wait (the wave barrier of the rows and tiles kernels) as synchronisation;
# the same kernels with a CTA barrier between the waves must be hazard free:
if [ -f lz4_b200/build/liblz4_b200_barsync.so ]; then
  run "debug build: CTA barrier between waves (-DLZ4K_WAVE_BARSYNC)" racecheck LZ4_B200_LIBRARY=$PWD/lz4_b200/build/liblz4_b200_barsync.so
fi
rm -f gpurun_out/sanitizer_raw.txt
cat $O
