"""BASELINE config 4 shape on one GPU: one LZ4 frame of 4 MB independent blocks through the frame layer (HOST buffers,
the copies inside the timed region): LZ4B200_compressFrame_host / LZ4B200_decompressFrame_host called on numpy memory
(pageable) and on pinned memory; checked against the reference's LZ4F_compressFrame digest when available.
Usage (under gpurun): python tests/perf/frame_bench.py [GiB] [blockSizeID]"""
import ctypes as C
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lz4_b200 import _lib  # noqa: E402
from oracle.pyoracle import Oracle, Reference, have_reference  # noqa: E402

import torch  # noqa: E402

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
bsid = int(sys.argv[2]) if len(sys.argv) > 2 else 7
lib = _lib.load()
orc = Oracle()
n = int(gib * (1 << 30))
data = orc.datagen_mt(n, 64 << 20, 0.5, 0)
cap = int(lib.LZ4B200_compressFrameBound(n, bsid))
row = {"GiB": gib, "blockSizeID": bsid}


def run(tag, src_ptr, frame_ptr, back_ptr):
    def comp():
        r = int(lib.LZ4B200_compressFrame_host(src_ptr, n, frame_ptr, cap, bsid, 0, 1))
        assert r > 0, r
        return r
    comp()                                                      # warm-up (context, allocations)
    t = []
    for _ in range(3):
        t0 = time.perf_counter(); fb = comp(); t.append(time.perf_counter() - t0)
    consumed = C.c_int64(0)
    d = []
    for _ in range(3):
        t0 = time.perf_counter()
        r = int(lib.LZ4B200_decompressFrame_host(frame_ptr, fb, back_ptr, n, C.byref(consumed)))
        d.append(time.perf_counter() - t0)
        assert r == n, r
    row["compressFrame_host_GBps_" + tag] = round(n / min(t) / 1e9, 3)
    row["decompressFrame_host_GBps_" + tag] = round(n / min(d) / 1e9, 3)
    return fb


frame = np.empty(cap, dtype=np.uint8)
back = np.empty(n, dtype=np.uint8)
fb = run("pageable", data.ctypes.data, frame.ctypes.data, back.ctypes.data)
assert (back == data).all()
row["frame_bytes"] = fb
row["ratio"] = round(n / fb, 4)
p_src = torch.empty(n, dtype=torch.uint8, pin_memory=True); p_src.numpy()[:] = data
p_frame = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
p_back = torch.empty(n, dtype=torch.uint8, pin_memory=True)
fb2 = run("pinned", p_src.data_ptr(), p_frame.data_ptr(), p_back.data_ptr())
assert fb2 == fb and bool((p_back.numpy() == data).all()) and bool((p_frame.numpy()[:fb] == frame[:fb]).all())
if have_reference() and Reference().have_frame():
    ref = Reference()
    t4 = time.perf_counter(); rf = ref.compress_frame(data, bsid, 0, True); t5 = time.perf_counter()
    row["byte_identical_to_LZ4F_compressFrame"] = hashlib.sha256(rf).digest() == hashlib.sha256(frame[:fb].tobytes()).digest()
    row["reference_compressFrame_GBps_1_thread"] = round(n / (t5 - t4) / 1e9, 3)
    t6 = time.perf_counter(); rb = ref.decompress_frame(frame[:fb].tobytes(), n); t7 = time.perf_counter()
    assert rb == data.tobytes()
    row["reference_decompress_GBps_1_thread"] = round(n / (t7 - t6) / 1e9, 3)
print(json.dumps(row))
