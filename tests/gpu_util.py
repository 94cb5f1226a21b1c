"""Helpers shared by the GPU parity tests."""
import hashlib

import numpy as np
import torch


def dev():
    return torch.device("cuda:0")


def to_dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev())


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


def pack_host_blocks(blocks):
    """list of bytes -> (u8 array, int64 offsets, int32 sizes)"""
    sizes = np.array([len(b) for b in blocks], dtype=np.int32)
    offs = np.zeros(len(blocks), dtype=np.int64)
    if len(blocks) > 1:
        offs[1:] = np.cumsum(sizes[:-1].astype(np.int64))
    buf = np.frombuffer(b"".join(blocks) + b"\x00" * 16, dtype=np.uint8).copy()
    return buf, offs, sizes


def decode_batch(blocks, caps):
    """Run blocks through LZ4B200_decompress_blocks with per-block capacities (device API).
    Returns list of (ret, bytes)."""
    from lz4_b200 import _lib
    lib = _lib.load()
    buf, offs, sizes = pack_host_blocks(blocks)
    caps = np.asarray(caps, dtype=np.int32)
    dst_off = np.zeros(len(blocks), dtype=np.int64)
    slot = (np.maximum(caps, 0).astype(np.int64) + 64)
    dst_off[1:] = np.cumsum(slot[:-1])
    total = int(slot.sum())
    d_buf, d_offs, d_sizes = to_dev(buf), to_dev(offs), to_dev(sizes)
    d_caps, d_dst_off = to_dev(caps), to_dev(dst_off)
    d_out = torch.full((total,), 0xA5, dtype=torch.uint8, device=dev())
    d_ret = torch.zeros(len(blocks), dtype=torch.int32, device=dev())
    ws = torch.empty(int(lib.LZ4B200_decompress_workspace_bytes(len(blocks))), dtype=torch.uint8, device=dev())
    rc = lib.LZ4B200_decompress_blocks(d_buf.data_ptr(), d_offs.data_ptr(), d_sizes.data_ptr(), d_out.data_ptr(),
                                       d_dst_off.data_ptr(), 0, d_caps.data_ptr(), 0, d_ret.data_ptr(), len(blocks),
                                       ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "decompress_blocks")
    torch.cuda.synchronize()
    out = d_out.cpu().numpy()
    ret = d_ret.cpu().numpy()
    res = []
    for i in range(len(blocks)):
        o, c = int(dst_off[i]), max(int(caps[i]), 0)
        guard = out[o + c:o + c + 64]
        assert (guard == 0xA5).all(), "block %d wrote past its capacity" % i
        r = int(ret[i])
        res.append((r, out[o:o + max(r, 0)].tobytes()))
    return res


def decode_batch_uniform(blocks, cap, wide=True):
    """Run blocks through LZ4B200_decompress_blocks with ONE capacity for all of them (device API); wide=True gives the
    workspace LZ4B200_decompress_workspace_bytes_for asks for (blocks above 64 KB: the tiles kernel), wide=False the
    small one (blocks above 64 KB: the generic kernel).  Returns list of (ret, bytes)."""
    from lz4_b200 import _lib
    lib = _lib.load()
    buf, offs, sizes = pack_host_blocks(blocks)
    n = len(blocks)
    stride = (cap + 64 + 15) // 16 * 16
    d_buf, d_offs, d_sizes = to_dev(buf), to_dev(offs), to_dev(sizes)
    d_out = torch.full((n * stride,), 0xA5, dtype=torch.uint8, device=dev())
    d_ret = torch.zeros(n, dtype=torch.int32, device=dev())
    nbytes = int(lib.LZ4B200_decompress_workspace_bytes_for(n, 0, cap)) if wide else int(lib.LZ4B200_decompress_workspace_bytes(n))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev())
    rc = lib.LZ4B200_decompress_blocks(d_buf.data_ptr(), d_offs.data_ptr(), d_sizes.data_ptr(), d_out.data_ptr(),
                                       None, stride, None, cap, d_ret.data_ptr(), n,
                                       ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "decompress_blocks")
    torch.cuda.synchronize()
    out = d_out.cpu().numpy()
    ret = d_ret.cpu().numpy()
    res = []
    for i in range(n):
        o = i * stride
        assert (out[o + cap:o + stride] == 0xA5).all(), "block %d wrote past its capacity" % i
        r = int(ret[i])
        res.append((r, out[o:o + max(r, 0)].tobytes()))
    return res
