#!/bin/bash
# round 2, GPU call S: racecheck of the parallel compressor (full hazard list), new tiles parity test
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 60 python -m pytest tests/test_gpu_parallel_compress.py::test_limited_output_and_never_past_capacity -x -q -m gpu > $O/r2s_race_enc.txt 2>&1
grep -E "Race reported|Read access|Write access|hazard|RACECHECK|passed|failed" $O/r2s_race_enc.txt | cut -c1-220 | head -60
echo "== initcheck"
timeout 900 compute-sanitizer --tool initcheck --print-limit 30 python -m pytest tests/test_gpu_parallel_compress.py::test_limited_output_and_never_past_capacity -x -q -m gpu > $O/r2s_init_enc.txt 2>&1
grep -E "Uninitialized|at |ERROR SUMMARY|passed|failed" $O/r2s_init_enc.txt | cut -c1-220 | head -30
echo "== tiles test"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "big_blocks or large_blocks" 2>&1 | tail -8 | cut -c1-300
