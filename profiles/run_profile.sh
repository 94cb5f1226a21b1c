#!/bin/bash
# Profile recipe (B200_PROFILING.md).  Usage (under gpurun): bash profiles/run_profile.sh <tag> [kernel-regex]
#  1. launch list of the SAME command the bench numbers come from (default 4 GiB batch), device times per
#     launch (cold cache, serialised: compare shares, not absolutes);
#  2. one `--set full` capture of the dominant kernel on a 0.5 GiB batch (8192 blocks; ncu replays the kernel
#     ~40 times, so the batch is kept small -- per-block behaviour is identical, the kernel is persistent).
TAG=${1:-r01}
KRE=${2:-"lz4_expand_rows_kernel|lz4_scan_kernel|lz4_encode_par_kernel"}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"${KRE}" -s 6 -c 3 -f -o gpurun_out/prof_${TAG} \
    python bench.py --gib 0.5 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out | tail -8
