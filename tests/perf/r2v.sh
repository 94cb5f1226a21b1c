#!/bin/bash
# round 2, GPU call V: the evidence committed under profiles/ -- bench lines (ours, P90, 4 MB blocks, reference arm),
# ncu launch list + full captures, sanitizers
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python bench.py --ceiling 2>$O/bench_r02_n1.err | tail -1 > $O/bench_r02_n1.json
timeout 400 python bench.py --impl reference --steps 5 --warmup 2 2>$O/bench_r02_reference.err | tail -1 > $O/bench_r02_reference.json
timeout 300 python bench.py --no-cpu --no-e2e --proba 0.9 2>$O/bench_r02_p90.err | tail -1 > $O/bench_r02_p90.json
timeout 400 python bench.py --block-kb 4096 --gib 1 --no-cpu --no-e2e --steps 5 2>$O/bench_r02_4mb.err | tail -1 > $O/bench_r02_4mb.json
python - <<'PY'
import json
for f in ('n1','reference','p90','4mb'):
    try:
        d=json.load(open('gpurun_out/bench_r02_%s.json'%f))
        print(f, {k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, 'e2e', (d.get('e2e') or {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'),
              'par', (d.get('compress_parallel') or {}).get('GBps'), 'warn', d.get('warning'))
    except Exception as e: print(f,'FAILED',e); print(open('gpurun_out/bench_r02_%s.err'%f).read()[-600:])
PY
bash profiles/run_profile.sh r02q > $O/r2v_profile.log 2>&1; tail -2 $O/r2v_profile.log
bash tests/perf/sanitize.sh r02 > /dev/null 2>&1; cat $O/sanitizer_r02.txt
