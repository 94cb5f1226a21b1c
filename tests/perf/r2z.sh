#!/bin/bash
# round 2, GPU call Z: the final tree -- GPU tests, smoke, the compressor under memcheck, bench lines
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | cut -c1-250 | tee $O/r2z_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== compressor under memcheck"; TAG=final timeout 600 compute-sanitizer --tool memcheck python tests/perf/enc_determinism.py 2>&1 | grep -E "run 0 sizes|DOES NOT|DIFFERS|BAD|ERROR SUMMARY" | cut -c1-300
timeout 600 python bench.py --ceiling 2>$O/bench_r02_n1.err | tail -1 > $O/bench_r02_n1.json
timeout 300 python bench.py --no-cpu --no-e2e --proba 0.9 2>$O/bench_r02_p90.err | tail -1 > $O/bench_r02_p90.json
python - <<'PY'
import json
for f in ('n1','p90'):
    try:
        d=json.load(open('gpurun_out/bench_r02_%s.json'%f))
        print(f, {k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, 'e2e', (d.get('e2e') or {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'),
              'exact', d['compress']['GBps'], 'par', d['compress_parallel']['GBps'], d['compress_parallel']['ratio_vs_reference'])
    except Exception as e: print(f,'FAILED',e); print(open('gpurun_out/bench_r02_%s.err'%f).read()[-600:])
PY
