"""One-shot block API: the Python face of the drop-in C functions (lib/lz4.h:191,208,236).

Same names, argument meaning and error behaviour as the reference's C API; data are host bytes.
Each call is LZ4_* in liblz4_b200.so: host -> device copy, CUDA kernels, device -> host copy.
"""
import ctypes as C

import numpy as np

from . import _lib


def _u8(data):
    if isinstance(data, np.ndarray):
        assert data.dtype == np.uint8
        return np.ascontiguousarray(data)
    return np.frombuffer(bytes(data), dtype=np.uint8)


def LZ4_compressBound(input_size):
    return int(_lib.load().LZ4_compressBound(int(input_size)))


def LZ4_compress_fast(src, dst_capacity=None, acceleration=1):
    """Returns (ret, bytes): ret bytes written, 0 on failure -- lz4.h:236."""
    lib = _lib.load()
    s = _u8(src)
    cap = LZ4_compressBound(len(s)) if dst_capacity is None else int(dst_capacity)
    dst = np.empty(max(cap, 1), dtype=np.uint8)
    r = int(lib.LZ4_compress_fast(s.ctypes.data if len(s) else None, dst.ctypes.data, len(s), cap, int(acceleration)))
    return r, dst[:max(r, 0)].tobytes()


def LZ4_compress_default(src, dst_capacity=None):
    """lz4.h:191."""
    lib = _lib.load()
    s = _u8(src)
    cap = LZ4_compressBound(len(s)) if dst_capacity is None else int(dst_capacity)
    dst = np.empty(max(cap, 1), dtype=np.uint8)
    r = int(lib.LZ4_compress_default(s.ctypes.data if len(s) else None, dst.ctypes.data, len(s), cap))
    return r, dst[:max(r, 0)].tobytes()


def LZ4_decompress_safe(src, dst_capacity):
    """Returns (ret, bytes): ret decoded size or negative error -- lz4.h:208."""
    lib = _lib.load()
    s = _u8(src)
    cap = int(dst_capacity)
    dst = np.empty(max(cap, 1), dtype=np.uint8)
    sp = s.ctypes.data if len(s) else np.zeros(1, dtype=np.uint8).ctypes.data
    r = int(lib.LZ4_decompress_safe(sp, dst.ctypes.data, len(s), cap))
    return r, dst[:max(r, 0)].tobytes()
