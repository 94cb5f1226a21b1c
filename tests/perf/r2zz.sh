#!/bin/bash
# last GPU call of round 2: ncu capture of the shipped compressor, sanitizers on the final tree
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"encode_par" -s 1 -c 1 -f -o $O/prof_r02z \
    python bench.py --gib 0.5 --steps 2 --warmup 3 --no-cpu --no-e2e > $O/ncu_full_r02z.log 2>&1
ls -la $O/prof_r02z.ncu-rep
bash tests/perf/sanitize.sh r02 > /dev/null 2>&1; cat $O/sanitizer_r02.txt
