#!/bin/bash
# round 2, GPU call U: final validation -- GPU tests, sanitizers, determinism of the compressor under memcheck, bench
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | cut -c1-250 | tee $O/r2u_pytest.txt
echo "== compressor under memcheck"; TAG=memcheck timeout 600 compute-sanitizer --tool memcheck python tests/perf/enc_determinism.py 2>&1 | grep -E "run 0 sizes|DOES NOT|DIFFERS|ERROR SUMMARY" | head -8 | cut -c1-300
timeout 300 python bench.py --no-cpu --no-e2e --steps 6 2>$O/r2u_p50.err | tail -1 > $O/r2u_p50.json
timeout 300 python bench.py --no-cpu --no-e2e --steps 6 --proba 0.9 2>$O/r2u_p90.err | tail -1 > $O/r2u_p90.json
python - <<'PY'
import json
for f in ('p50','p90'):
    try:
        d=json.load(open('gpurun_out/r2u_%s.json'%f)); r=d['roofline']
        print(f, d['value'],'GB/s  step',d['ms_per_step'],'ms  scan',r['scan_kernel_ms'],' expand',r['kernel_ms'], 'compress', d['compress']['GBps'], 'parallel', d['compress_parallel']['GBps'], d['compress_parallel']['ratio'], d['compress_parallel']['ratio_vs_reference'])
    except Exception as e: print(f,'FAILED',e); print(open('gpurun_out/r2u_%s.err'%f).read()[-800:])
PY
bash tests/perf/sanitize.sh r02 > /dev/null 2>&1; cat $O/sanitizer_r02.txt
