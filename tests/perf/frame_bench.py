"""BASELINE config 4 shape on one GPU: one LZ4 frame of 4 MB independent blocks through the frame layer
(host buffers): compress, check against the reference's LZ4F_compressFrame digest when available, decompress.
Usage (under gpurun): python tests/perf/frame_bench.py [GiB] [blockSizeID]"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lz4_b200 import frame  # noqa: E402
from oracle.pyoracle import Oracle, Reference, have_reference  # noqa: E402

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
bsid = int(sys.argv[2]) if len(sys.argv) > 2 else 7
orc = Oracle()
n = int(gib * (1 << 30))
data = orc.datagen_mt(n, 64 << 20, 0.5, 0)
frame.compress_frame(data[:1 << 22], bsid, 0, True)        # warm-up (context, allocations)
t0 = time.perf_counter(); f = frame.compress_frame(data, bsid, 0, True); t1 = time.perf_counter()
back = frame.decompress_frame(f, n)
t2 = time.perf_counter(); back = frame.decompress_frame(f, n); t3 = time.perf_counter()
assert back == data.tobytes()
row = {"GiB": gib, "blockSizeID": bsid, "frame_bytes": len(f), "ratio": round(n / len(f), 4),
       "compressFrame_host_GBps": round(n / (t1 - t0) / 1e9, 3), "decompressFrame_host_GBps": round(n / (t3 - t2) / 1e9, 3)}
if have_reference() and Reference().have_frame():
    ref = Reference()
    t4 = time.perf_counter(); rf = ref.compress_frame(data, bsid, 0, True); t5 = time.perf_counter()
    row["byte_identical_to_LZ4F_compressFrame"] = hashlib.sha256(rf).digest() == hashlib.sha256(f).digest()
    row["reference_compressFrame_GBps_1_thread"] = round(n / (t5 - t4) / 1e9, 3)
    t6 = time.perf_counter(); rb = ref.decompress_frame(f, n); t7 = time.perf_counter()
    assert rb == data.tobytes()
    row["reference_decompress_GBps_1_thread"] = round(n / (t7 - t6) / 1e9, 3)
print(json.dumps(row))
