"""Round-2 probe: how much of the scan's latency can be hidden by overlapping it with the expand kernel
of the previous batch?  Uses only the public phased API (phases=1 scan, phases=2 expand) on two
streams with two workspaces: scan(k+1) runs on stream A while expand(k) runs on stream B.

The scan is latency bound (IPC ~0.2, 128-thread CTAs without shared memory), the expand kernel is a
persistent 1024-thread CTA per SM that is issue bound, so the two can share the SMs.
Prints one JSON line: sequential ms/step vs pipelined ms/step, both verified against the source.
Usage (under gpurun): python tests/perf/pipeline_probe.py [GiB] [steps]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lz4_b200 import _lib, batch  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402

BLOCK = 65536
gib = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
n_blocks = int(gib * (1 << 30)) // BLOCK
total = n_blocks * BLOCK
lib = _lib.load()
data = Oracle().datagen_mt(total, 64 << 20, 0.5, 0)
src = torch.from_numpy(data).cuda()
slots, sizes, stride = batch.compress_blocks(src, BLOCK, 1)
packed, offs = batch.pack_blocks(slots, stride, sizes)
offs = offs[:-1].contiguous()
del slots
out = torch.empty_like(src)
rets = torch.empty(n_blocks, dtype=torch.int32, device="cuda")
ws_bytes = int(lib.LZ4B200_decompress_workspace_bytes(n_blocks))
ws = [torch.empty(ws_bytes, dtype=torch.uint8, device="cuda") for _ in range(2)]
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()


def dec(phases, w, stream):
    batch.decompress_blocks(packed, offs, sizes, BLOCK, out=out, out_sizes=rets, workspace=w, stream=stream, phases=phases)


# sequential reference: scan then expand on one stream
for _ in range(3):
    dec(3, ws[0], sa)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(sa)
for _ in range(steps):
    dec(1, ws[0], sa)
    dec(2, ws[0], sa)
e1.record(sa)
torch.cuda.synchronize()
seq_ms = e0.elapsed_time(e1) / steps
assert torch.equal(out, src)

# pipelined: scan(k+1) on stream A overlaps expand(k) on stream B
out.zero_()
scanned = [torch.cuda.Event() for _ in range(steps + 1)]
expanded = [torch.cuda.Event() for _ in range(steps + 1)]
torch.cuda.synchronize()
p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
p0.record(sa)
sb.wait_event(p0)
for k in range(steps):
    w = ws[k & 1]
    if k >= 2:
        sa.wait_event(expanded[k - 2])          # workspace k&1 is free again
    dec(1, w, sa)
    scanned[k].record(sa)
    sb.wait_event(scanned[k])
    dec(2, w, sb)
    expanded[k].record(sb)
p1.record(sb)
torch.cuda.synchronize()
pipe_ms = p0.elapsed_time(p1) / steps
assert torch.equal(out, src)
print(json.dumps({"blocks": n_blocks, "steps": steps, "sequential_ms_per_step": round(seq_ms, 3),
                  "pipelined_ms_per_step": round(pipe_ms, 3),
                  "sequential_GBps": round(total / seq_ms / 1e6, 1), "pipelined_GBps": round(total / pipe_ms / 1e6, 1)}))
