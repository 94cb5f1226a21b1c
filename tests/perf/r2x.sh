#!/bin/bash
# round 2, GPU call X (8 GPUs): BASELINE config 4's shape -- one frame of 4 MB blocks sharded over the GPUs (4 GiB per rank),
# compress + decompress + exchange of the decoded shards + compressed gather
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
N=${1:-8}
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --block-kb 4096 --gib 4 --steps 3 --warmup 3 --no-cpu --no-e2e 2>$O/bench_r02_config4_n$N.err | tail -1 > $O/bench_r02_config4_n$N.json
python - $N <<'PY'
import json,sys
N=sys.argv[1]
try:
    d=json.load(open('gpurun_out/bench_r02_config4_n%s.json'%N)); m=d['multi_gpu']
    print('config 4 on',N,'GPUs: value',d['value'],'GB/s, step',d['ms_per_step'],'ms; codec only',m['codec_only'],'; per rank',m['per_rank_ms_per_step'],'; compress',d['compress']['GBps'],'; compressed gather',m.get('compressed_reassembly'))
except Exception as e: print('FAILED',e); print(open('gpurun_out/bench_r02_config4_n%s.err'%N).read()[-2500:])
PY
