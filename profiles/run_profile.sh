#!/bin/bash
# Profile recipe (B200_PROFILING.md): launch list of a short bench run + one full capture of the
# dominant kernel.  Usage (under gpurun): bash profiles/run_profile.sh <tag> [kernel-regex]
TAG=${1:-r01}
KRE=${2:-expand_fast}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --gib 0.5 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:${KRE} -s 3 -c 1 -f -o gpurun_out/prof_${TAG} \
    python bench.py --gib 0.5 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out | tail -8
