#!/bin/bash
# round 2, GPU call K (2 GPUs): the N > 1 bench path (decode + overlapped exchange in the timed step)
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --e2e-steps 2 > $O/r2k_n$N.json 2> $O/r2k_n$N.err
tail -c 3000 $O/r2k_n$N.json; tail -5 $O/r2k_n$N.err
for c in 1 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --no-e2e --chunks $c 2>$O/r2k_n${N}_c$c.err | tail -1 > $O/r2k_n${N}_c$c.json
python - $N $c <<'PY'
import json,sys
N,c=sys.argv[1],sys.argv[2]
try:
    d=json.load(open('gpurun_out/r2k_n%s_c%s.json'%(N,c))); print('chunks',c,'value',d['value'],'ms',d['ms_per_step'],d['multi_gpu']['codec_only'],d['multi_gpu']['per_rank_ms_per_step'])
except Exception as e: print('chunks',c,'FAILED',e); print(open('gpurun_out/r2k_n%s_c%s.err'%(N,c)).read()[-1500:])
PY
done
