"""Multi-GPU plumbing: block sharding across ranks and NCCL all-gather reassembly.

Blocks are independent units: rank r of R owns the contiguous block range
[r*N/R, (r+1)*N/R).  The codec itself needs no communication; the only exchange step is the
reassembly of the frame: one in-place all-gather of the fixed-size decoded shards, or -- for
compressed output -- an all-gather of the int32 size table followed by an all-gather of the
shards padded to the largest one.  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is the
transport; nothing here touches the bytes.
"""
import torch
import torch.distributed as dist


def shard_range(n_blocks, rank, world):
    """Contiguous block range [lo, hi) owned by `rank` (sizes differ by at most one block)."""
    lo = n_blocks * rank // world
    hi = n_blocks * (rank + 1) // world
    return lo, hi


def _all_gather_flat(out, shard, group=None):
    """all_gather_into_tensor where the backend has it (NCCL), list all_gather otherwise (gloo)."""
    world = dist.get_world_size(group)
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(out, shard, group=group)
    else:
        chunks = list(out.view(world, -1).unbind(0))
        dist.all_gather(chunks, shard, group=group)
    return out


def allgather_decoded(full, n_blocks, block_size, group=None):
    """Reassemble a decoded frame: every rank decoded its shard INTO ITS SLICE of `full`
    (u8[n_blocks*block_size]); after the call every rank holds the whole frame.

    Requires n_blocks % world == 0 (equal shards => one in-place all-gather, no padding)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if n_blocks % world:
        raise ValueError("allgather_decoded needs n_blocks divisible by the world size")
    per = n_blocks // world * block_size
    shard = full[rank * per:(rank + 1) * per]
    return _all_gather_flat(full[:world * per], shard, group)


def chunk_ranges(n_units, n_chunks):
    """[lo, hi) of each of n_chunks nearly equal contiguous pieces of range(n_units) (empty pieces dropped)."""
    out = []
    for k in range(n_chunks):
        lo, hi = n_units * k // n_chunks, n_units * (k + 1) // n_chunks
        if hi > lo:
            out.append((lo, hi))
    return out


class PeerFrame:
    """Every rank's copy of the frame, mapped into every other rank's address space (CUDA IPC).

    With it the exchange of a decoded chunk is R-1 device-to-device copies issued by the rank that decoded it, straight
    into the chunk's final place in each peer's frame (LZ4B200_peer_copy_async), on one stream per peer: copy engines, no
    SM taken from the codec's persistent kernels.  MEASURED (2 B200s of an NVSwitch node, 4 GiB per rank): 27 GB/s -- the
    copies into IPC-mapped memory do not take NVLink there -- against 650 GB/s for NCCL's send/recv kernels, so NCCL is the
    default exchange and this class is for systems where peer copies are fast (profiles/experiments_r02.txt).
    Collective construction: every rank of `group` must call it with its own `full`."""

    def __init__(self, full, group=None):
        from torch.multiprocessing.reductions import reduce_tensor
        from . import _lib
        self.lib = _lib.load()
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.full = full
        handles = [None] * self.world
        dist.all_gather_object(handles, reduce_tensor(full), group=group)
        self.peers = [None] * self.world                       # peers[r] = rank r's frame, writable from here
        self.streams = [torch.cuda.Stream(device=full.device) for _ in range(self.world - 1)]
        self.token = torch.zeros(1, dtype=torch.int32, device=full.device)
        err = None
        try:                                                   # (no collective inside: a rank that fails must not strand the others)
            for r, (fn, fargs) in enumerate(handles):
                if r != self.rank:
                    self.peers[r] = fn(*fargs)
                    rc = self.lib.LZ4B200_peer_copy_async(None, self.peers[r].device.index, None, 0, None)   # probe: enables peer access
                    if rc != 0:
                        raise RuntimeError("device %d cannot reach device %d directly (LZ4B200_peer_copy_async: %d)"
                                           % (full.device.index, self.peers[r].device.index, rc))
        except Exception as e:                                 # noqa: BLE001
            err = e
        ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=full.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)  # everybody has mapped everybody, or nobody uses the mapping
        if int(ok.item()) == 0:
            raise RuntimeError("peer mapping failed on at least one rank%s" % (": %s" % err if err else ""))

    def push(self, lo, hi):
        """Enqueue the copy of full[lo:hi] (this rank's freshly decoded bytes, ordered after the current stream's work so
        far) into the same place of every peer's frame."""
        ev = torch.cuda.Event()
        ev.record()
        for i in range(self.world - 1):
            r = (self.rank + 1 + i) % self.world               # staggered: no hot receiver
            st = self.streams[i]
            st.wait_event(ev)
            rc = self.lib.LZ4B200_peer_copy_async(self.peers[r].data_ptr() + lo, self.peers[r].device.index,
                                                  self.full.data_ptr() + lo, hi - lo, st.cuda_stream)
            if rc != 0:
                raise RuntimeError("LZ4B200_peer_copy_async failed (%d): %s" % (rc, self.lib.LZ4B200_last_cuda_error().decode()))

    def finish(self):
        """Order the current stream after this rank's pushes and after every peer's pushes into this rank."""
        cur = torch.cuda.current_stream()
        for st in self.streams:
            cur.wait_stream(st)
        dist.all_reduce(self.token, group=self.group)          # a rank contributes only after its own pushes (stream order)


def decode_and_allgather(full, blocks_per_rank, block_size, decode_chunk, n_chunks=1, group=None, peer=None):
    """Decode this rank's shard chunk by chunk and exchange every chunk as soon as it is decoded, so that the
    exchange of chunk k runs while chunk k+1 is being decoded (the ordered-writer role of lz4io.c:594-635, spread
    over the ranks: afterwards every rank holds the whole decoded frame).

    full            u8[world * blocks_per_rank * block_size]; rank r's shard is the r-th equal slice.
    decode_chunk    decode_chunk(lo, hi): enqueue, on the CURRENT stream, the decode of this rank's blocks [lo, hi)
                    into their final place in `full`.
    Exchange of one chunk = one grouped send/recv (NCCL P2P over NVLink): every rank receives its peers' bytes
    straight into their final place in `full` -- no staging buffer and no re-ordering copy, which a chunk-wise
    ncclAllGather (contiguous output per call) would need.  The collective stream waits for the decode of chunk k
    through the event torch.distributed records at issue time; the codec stream carries on with chunk k+1.
    n_chunks = 1 (default) is the serial form, decode then exchange: measured faster than any overlap today, because
    every extra chunk costs a whole scan kernel, whose time does not shrink with the batch.
    With `peer` (a PeerFrame over `full`) the exchange of a chunk is R-1 copy-engine copies into the peers' frames
    instead (no SM use; see PeerFrame).
    Returns the list of outstanding works (already waited for: the current stream is ordered after them)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = blocks_per_rank * block_size
    works = []
    for lo, hi in chunk_ranges(blocks_per_rank, n_chunks):
        decode_chunk(lo, hi)
        if world == 1:
            continue
        a, b = lo * block_size, hi * block_size
        if peer is not None:
            peer.push(rank * per + a, rank * per + b)
            continue
        ops = []
        for d in range(1, world):                              # peer order staggered per rank: no hot receiver
            to, frm = (rank + d) % world, (rank - d) % world
            ops.append(dist.P2POp(dist.isend, full[rank * per + a:rank * per + b], to, group))
            ops.append(dist.P2POp(dist.irecv, full[frm * per + a:frm * per + b], frm, group))
        works += dist.batch_isend_irecv(ops)
    for wk in works:
        wk.wait()
    if peer is not None and world > 1:
        peer.finish()
    return works


def allgather_compressed(packed, total_bytes, sizes, group=None):
    """Reassemble a compressed frame from per-rank packed shards.

    packed: this rank's contiguous compressed bytes (u8, at least total_bytes long);
    sizes : int32[blocks_per_rank] compressed size of each local block (same count on all ranks).
    Returns (all_sizes int32[R*blocks_per_rank], shards u8[R, max_shard], shard_bytes int64[R]):
    rank r's bytes are shards[r, :shard_bytes[r]]; concatenated in rank order they are the frame's
    block payloads in block order (the ordered-writer role of lz4io.c:594-635).
    """
    world = dist.get_world_size(group)
    dev = packed.device
    all_sizes = torch.empty(world * sizes.numel(), dtype=torch.int32, device=dev)
    _all_gather_flat(all_sizes, sizes.contiguous(), group)
    mine = torch.tensor([int(total_bytes)], dtype=torch.int64, device=dev)
    shard_bytes = torch.empty(world, dtype=torch.int64, device=dev)
    _all_gather_flat(shard_bytes, mine, group)
    max_shard = int(shard_bytes.max().item())
    max_shard = (max_shard + 15) // 16 * 16
    padded = torch.zeros(max_shard, dtype=torch.uint8, device=dev)
    padded[:int(total_bytes)] = packed[:int(total_bytes)]
    shards = torch.empty(world * max_shard, dtype=torch.uint8, device=dev)
    _all_gather_flat(shards, padded, group)
    return all_sizes, shards.view(world, max_shard), shard_bytes
