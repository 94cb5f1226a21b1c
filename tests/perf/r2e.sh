#!/bin/bash
# round 2, GPU call E: parallel-parse compressor, e2e chunk sizes, reference arm (bounded sample), sanitizers
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $O/r2e_pytest.txt
timeout 600 python bench.py --no-cpu --e2e-steps 3 > $O/r2e_full.json 2>$O/r2e_full.err; tail -3 $O/r2e_full.err
for mb in 32 64 256; do
  LZ4B200_HOST_CHUNK_MB=$mb timeout 300 python bench.py --no-cpu --steps 3 2>$O/r2e_chunk$mb.err | tail -1 > $O/r2e_chunk$mb.json
done
python - <<'PY'
import json
for f in ("full","chunk32","chunk64","chunk256"):
    try:
        d=json.load(open("gpurun_out/r2e_%s.json"%f))
        print(f,"value",d["value"],"e2e",d["e2e"]["value"],"compress",d["compress"]["GBps"],"parallel",d["compress_parallel"]["GBps"],"ratio",d["compress_parallel"]["ratio"],d["compress_parallel"]["ratio_vs_reference"])
    except Exception as e: print(f,"FAILED",e); print(open("gpurun_out/r2e_%s.err"%f).read()[-2000:])
PY
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > $O/r2e_ref.json 2>$O/r2e_ref.err; python -c "
import json; d=json.load(open('gpurun_out/r2e_ref.json')); print('reference', d['value'], d['timing'], d.get('warning'))"
bash tests/perf/sanitize.sh r02 > /dev/null 2>&1; cat $O/sanitizer_r02.txt | tail -30
