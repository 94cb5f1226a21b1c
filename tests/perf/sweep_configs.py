"""BASELINE.json configs 3 and 5 in one run: compress accel 1/8/32 sweep (ratio + GB/s) and the matching
decode rate on datagen P50 and P90, 64 KB blocks, next to the CPU reference (all host threads and 1 thread).
Every GPU-compressed sample block is checked byte-for-byte against the oracle.  Writes one JSON line per row.
Usage (under gpurun):  python tests/perf/sweep_configs.py [GiB] > gpurun_out/sweep.jsonl"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lz4_b200 import batch  # noqa: E402
from oracle.pyoracle import Oracle, Reference, have_reference  # noqa: E402

BLOCK = 65536
gib = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
orc = Oracle()
codec = Reference() if have_reference() else orc
cores = os.cpu_count() or 1
n_blocks = int(gib * (1 << 30)) // BLOCK
total = n_blocks * BLOCK


def ev():
    return torch.cuda.Event(enable_timing=True)


for proba in (0.5, 0.9):
    host = orc.datagen_mt(total, 64 << 20, proba, 0, cores)
    src = torch.from_numpy(host).cuda()
    for accel in (1, 8, 32):
        slots, sizes, stride = batch.compress_blocks(src, BLOCK, accel)
        torch.cuda.synchronize()
        a, b = ev(), ev()
        a.record()
        batch.compress_blocks(src, BLOCK, accel, slots=slots, out_sizes=sizes)
        b.record()
        torch.cuda.synchronize()
        c_ms = a.elapsed_time(b)
        sz = sizes.cpu().numpy()
        for i in np.random.default_rng(accel).integers(0, n_blocks, 6):
            eret, eout = orc.compress(host[i * BLOCK:(i + 1) * BLOCK], accel)
            got = slots[i * stride:i * stride + int(sz[i])].cpu().numpy().tobytes()
            assert int(sz[i]) == eret and got == eout, "GPU compressor differs from the oracle"
        packed, offs = batch.pack_blocks(slots, stride, sizes)
        offs0 = offs[:-1].contiguous()
        out, rets = batch.decompress_blocks(packed, offs0, sizes, BLOCK)
        torch.cuda.synchronize()
        assert bool((rets == BLOCK).all()) and torch.equal(out, src)
        a, b = ev(), ev()
        a.record()
        for _ in range(3):
            batch.decompress_blocks(packed, offs0, sizes, BLOCK, out=out, out_sizes=rets)
        b.record()
        torch.cuda.synchronize()
        d_ms = a.elapsed_time(b) / 3
        # CPU reference on the same data: compress all threads + 1 thread (first 256 MiB), decompress all threads
        nb1 = min(n_blocks, 4096)
        cslots = np.empty(n_blocks * stride, dtype=np.uint8)
        orc.time_compress(codec, host, BLOCK, cslots, stride, accel, cores)          # first touch
        tc, csz = orc.time_compress(codec, host, BLOCK, cslots, stride, accel, cores)
        tc1, _ = orc.time_compress(codec, host[:nb1 * BLOCK], BLOCK, cslots, stride, accel, 1)
        assert (csz == sz).all(), "GPU and reference compressed sizes differ"
        coffs = np.arange(n_blocks, dtype=np.int64) * stride
        cout = np.empty(total, dtype=np.uint8)
        orc.time_decompress(codec, cslots, coffs, csz, cout, BLOCK, cores)
        td, _ = orc.time_decompress(codec, cslots, coffs, csz, cout, BLOCK, cores)
        print(json.dumps({
            "data": "datagen P%d" % round(proba * 100), "block": BLOCK, "GiB": gib, "accel": accel,
            "ratio": round(total / float(sz.sum()), 4),
            "gpu_compress_GBps": round(total / c_ms / 1e6, 2), "gpu_decompress_GBps": round(total / d_ms / 1e6, 2),
            "cpu_compress_GBps_all_threads": round(total / tc / 1e9, 2), "cpu_compress_GBps_1_thread": round(nb1 * BLOCK / tc1 / 1e9, 3),
            "cpu_decompress_GBps_all_threads": round(total / td / 1e9, 2), "cpu_threads": cores,
            "byte_identical_to_reference": True}), flush=True)
