"""lz4_b200 -- a B200 (sm_100a) LZ4 block codec behind the liblz4 C ABI.

    lz4_b200.block   one-shot LZ4_compress_default / LZ4_compress_fast / LZ4_decompress_safe
    lz4_b200.batch   batched device-pointer API on torch tensors (the measured path)
    lz4_b200.frame   one-shot LZ4 frames of independent blocks (LZ4F_compressFrame-compatible)
    lz4_b200.dist    block sharding across ranks + NCCL all-gather reassembly
    lz4_b200.build   in-tree build of liblz4_b200.so (nvcc + gcc)
"""
from . import _lib  # noqa: F401

__all__ = ["block", "batch", "frame", "dist", "build"]
