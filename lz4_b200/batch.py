"""Batched block codec on device-resident torch tensors (the measured path).

torch is plumbing here: it owns the device memory and the stream; all work is done by
LZ4B200_* in liblz4_b200.so on raw pointers.
"""
import torch

from . import _lib


def _stream(stream=None):
    return stream if stream is not None else torch.cuda.current_stream()


_workspaces = {}        # device index -> cached decode workspace (grown on demand, reused by later calls)


def _workspace(device, nbytes, stream):
    """One workspace per device, allocated on the launch stream: calls on the same stream are ordered by the
    stream; a caller that decodes on several streams at once passes its own `workspace`."""
    ws = _workspaces.get(device.index)
    if ws is None or ws.numel() < nbytes:
        with torch.cuda.stream(stream):
            ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _workspaces[device.index] = ws
    return ws


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise ValueError("device tensors required (got a CPU tensor)")


def compress_bound(n):
    return int(_lib.load().LZ4_compressBound(int(n)))


def decompress_blocks(comp, offsets, sizes, block_capacity, out=None, out_stride=None, out_sizes=None,
                      workspace=None, stream=None, phases=3):
    """Decode len(sizes) independent blocks.

    comp: u8[*] device buffer; block i = comp[offsets[i] : offsets[i]+sizes[i]] (int64 / int32 device
    tensors).  Block i is written to out[i*out_stride : ...] with capacity block_capacity.
    Returns (out, out_sizes) -- out_sizes[i] is LZ4_decompress_safe's return value.
    """
    lib = _lib.load()
    _need_cuda(comp, offsets, sizes, out)
    n = int(sizes.numel())
    stride = int(out_stride if out_stride is not None else block_capacity)
    st = _stream(stream)
    with torch.cuda.stream(st):                     # temporaries belong to the stream the kernels run on
        if out is None:
            out = torch.empty(n * stride, dtype=torch.uint8, device=comp.device)
        if out_sizes is None:
            out_sizes = torch.empty(n, dtype=torch.int32, device=comp.device)
    ws_bytes = int(lib.LZ4B200_decompress_workspace_bytes_for(n, 0, int(block_capacity)))
    if workspace is None or workspace.numel() < ws_bytes:
        workspace = _workspace(comp.device, ws_bytes, st)
    assert offsets.dtype == torch.int64 and sizes.dtype == torch.int32 and comp.dtype == torch.uint8
    rc = lib.LZ4B200_decompress_blocks_phased(comp.data_ptr(), offsets.data_ptr(), sizes.data_ptr(), out.data_ptr(),
                                              None, stride, None, int(block_capacity), out_sizes.data_ptr(), n,
                                              workspace.data_ptr(), workspace.numel(), int(phases),
                                              st.cuda_stream)
    _lib.check(rc, "LZ4B200_decompress_blocks")
    return out, out_sizes


def compress_blocks(src, block_size, acceleration=1, slots=None, slot_stride=None, slot_capacity=None,
                    out_sizes=None, src_sizes=None, stream=None, mode="exact"):
    """Compress src (u8 device tensor) as ceil(len/block_size) independent blocks.

    mode "exact": byte-identical to LZ4_compress_fast (LZ4B200_compress_blocks); mode "parallel": the
    parallel-parse encoder (LZ4B200_compress_blocks_parallel: valid, deterministic, ratio within 2 %, much faster).

    Returns (slots, out_sizes, slot_stride): block i's bytes are slots[i*slot_stride : i*slot_stride+out_sizes[i]].
    """
    lib = _lib.load()
    _need_cuda(src, slots)
    total = int(src.numel())
    n = (total + block_size - 1) // block_size if total else 0
    cap = int(slot_capacity if slot_capacity is not None else compress_bound(block_size))
    stride = int(slot_stride if slot_stride is not None else ((cap + 15) // 16) * 16)
    st = _stream(stream)
    with torch.cuda.stream(st):
        if slots is None:
            slots = torch.empty(max(n, 1) * stride, dtype=torch.uint8, device=src.device)
        if out_sizes is None:
            out_sizes = torch.empty(max(n, 1), dtype=torch.int32, device=src.device)
        if src_sizes is None and n and total != n * block_size:
            src_sizes = torch.full((n,), block_size, dtype=torch.int32, device=src.device)
            src_sizes[-1] = total - (n - 1) * block_size
    if mode not in ("exact", "parallel"):
        raise ValueError("mode must be 'exact' or 'parallel'")
    fn = lib.LZ4B200_compress_blocks if mode == "exact" else lib.LZ4B200_compress_blocks_parallel
    rc = fn(src.data_ptr(), int(block_size), src_sizes.data_ptr() if src_sizes is not None else None,
            int(block_size), slots.data_ptr(), stride, cap, int(acceleration), out_sizes.data_ptr(), n, st.cuda_stream)
    _lib.check(rc, "LZ4B200_compress_blocks")
    return slots, out_sizes[:n], stride


def pack_blocks(slots, slot_stride, sizes, header_bytes=0, packed=None, stream=None):
    """Gather variable-size blocks into one contiguous stream.  Returns (packed, offsets[n+1])."""
    lib = _lib.load()
    _need_cuda(slots, sizes)
    n = int(sizes.numel())
    st = _stream(stream)
    with torch.cuda.stream(st):
        offsets = torch.empty(n + 1, dtype=torch.int64, device=slots.device)
        if packed is None:
            packed = torch.empty(n * (int(slot_stride) + header_bytes) + 16, dtype=torch.uint8, device=slots.device)
    rc = lib.LZ4B200_pack_blocks(slots.data_ptr(), int(slot_stride), sizes.data_ptr(), n, packed.data_ptr(),
                                 offsets.data_ptr(), int(header_bytes), st.cuda_stream)
    _lib.check(rc, "LZ4B200_pack_blocks")
    return packed, offsets
