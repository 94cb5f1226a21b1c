#!/bin/bash
# round 2, GPU call R: ceiling kernel, 4 MB blocks (scan / expand split), final profile, sanitizer
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 | cut -c1-250 | tee $O/r2r_pytest.txt
timeout 300 python bench.py --ceiling --no-cpu --no-e2e --steps 6 2>$O/r2r_ceiling.err | tail -1 > $O/r2r_ceiling.json
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2r_ceiling.json')); r=d['roofline']; print('ceiling', r.get('ceiling'), 'step', d['ms_per_step'], 'expand', r['kernel_ms'], 'scan', r['scan_kernel_ms'], 'traffic', r['traffic'], r['step']['traffic'])
except Exception as e: print('FAILED', e); print(open('gpurun_out/r2r_ceiling.err').read()[-1500:])
PY
timeout 400 python bench.py --block-kb 4096 --gib 1 --no-cpu --no-e2e --steps 3 2>$O/r2r_4mb.err | tail -1 > $O/r2r_4mb.json
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2r_4mb.json')); r=d['roofline']; print('4MB blocks: value', d['value'], 'step', d['ms_per_step'], 'expand', r['kernel_ms'], 'scan', r['scan_kernel_ms'], 'compress', d['compress']['GBps'], 'par', d['compress_parallel']['GBps'])
except Exception as e: print('FAILED', e); print(open('gpurun_out/r2r_4mb.err').read()[-1500:])
PY
for id in 7 4; do timeout 300 python tests/perf/frame_bench.py 1 $id 2>&1 | tail -1 | cut -c1-600; done | tee $O/r2r_frame.jsonl
bash profiles/run_profile.sh r02q > $O/r2r_profile.log 2>&1; tail -3 $O/r2r_profile.log
bash tests/perf/sanitize.sh r02 > /dev/null 2>&1; cat $O/sanitizer_r02.txt
