#!/bin/bash
# round 2, GPU call Q (N GPUs): exchange by copy-engine pushes over peer memory against NCCL send/recv, chunk counts
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
N=${1:-2}
run() { # tag args...
  local tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
      bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --no-e2e "$@" 2>$O/r2q_n${N}_$tag.err | tail -1 > $O/r2q_n${N}_$tag.json
  python - $O/r2q_n${N}_$tag <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.load(open(f+'.json')); m=d['multi_gpu']
    print(f,'value',d['value'],'ms',d['ms_per_step'],'codec',m['codec_only'],'ranks',m['per_rank_ms_per_step'],'verified',m['exchange_verified'])
except Exception as e: print(f,'FAILED',e); print(open(f+'.err').read()[-2500:])
PY
}
run nccl_c4_r16 --exchange nccl --chunks 4 --reserve-sms 16
run nccl_c4_r32 --exchange nccl --chunks 4 --reserve-sms 32
run nccl_c8_r24 --exchange nccl --chunks 8 --reserve-sms 24
run nccl_c1 --exchange nccl --chunks 1
