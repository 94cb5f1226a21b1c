#!/bin/bash
# developer sweep: CTA size of the fast expand kernel (rebuilds on the GPU box; nvcc is in the image)
for T in 512 768 1024; do
  LZ4K_FAST_THREADS=$T LZ4K_PHASE_TIMING=1 python -m lz4_b200.build --force > /dev/null 2>&1
  echo "== threads $T"; timeout 200 python profiles/phase_timing.py 2>&1 | tail -9
done
