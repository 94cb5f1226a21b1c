"""N > 1 host logic on CPU: world_size 2, gloo backend.  The ranks shard the blocks of one frame
(lz4_b200.dist.shard_range), "decode"/"encode" their shard with the oracle (no GPU here), and the
reassembly collectives of lz4_b200.dist must rebuild exactly the single-process result."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BLOCK = 65536


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_blocks, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lz4_b200 import dist as ldist
    from oracle.pyoracle import Oracle
    orc = Oracle()
    data = orc.datagen(n_blocks * BLOCK, 0.5, 7)               # every rank can regenerate the frame
    lo, hi = ldist.shard_range(n_blocks, rank, world)
    # --- compress my shard, reassemble the compressed frame ---
    sizes, chunks = [], []
    for i in range(lo, hi):
        r, c = orc.compress(data[i * BLOCK:(i + 1) * BLOCK], 1)
        sizes.append(r)
        chunks.append(c)
    packed = torch.from_numpy(np.frombuffer(b"".join(chunks) + b"\0" * 16, dtype=np.uint8).copy())
    total = sum(sizes)
    all_sizes, shards, shard_bytes = ldist.allgather_compressed(packed, total, torch.tensor(sizes, dtype=torch.int32))
    frame = b"".join(shards[r, :int(shard_bytes[r])].numpy().tobytes() for r in range(world))
    # --- decode my shard into my slice of the full buffer, reassemble the decoded frame ---
    full = torch.zeros(n_blocks * BLOCK, dtype=torch.uint8)
    off = 0
    for j, i in enumerate(range(lo, hi)):
        r, out = orc.decompress(chunks[j], BLOCK)
        assert r == BLOCK
        full[i * BLOCK:(i + 1) * BLOCK] = torch.from_numpy(np.frombuffer(out, dtype=np.uint8).copy())
    ldist.allgather_decoded(full, n_blocks, BLOCK)
    ok_dec = bool((full.numpy() == data).all())
    # --- the overlapped form: decode chunk by chunk, exchange each chunk as soon as it is decoded ---
    if n_blocks % world == 0:
        full2 = torch.zeros(n_blocks * BLOCK, dtype=torch.uint8)
        per = n_blocks // world
        calls = []

        def decode_chunk(a, b):
            calls.append((a, b))
            for j in range(a, b):
                r, out = orc.decompress(chunks[j], BLOCK)
                assert r == BLOCK
                i = rank * per + j
                full2[i * BLOCK:(i + 1) * BLOCK] = torch.from_numpy(np.frombuffer(out, dtype=np.uint8).copy())

        ldist.decode_and_allgather(full2, per, BLOCK, decode_chunk, n_chunks=2)
        ok_dec = ok_dec and bool((full2.numpy() == data).all()) and calls == ldist.chunk_ranges(per, 2)
    if rank == 0:
        q.put((all_sizes.tolist(), frame, ok_dec, (lo, hi)))
    else:
        q.put((None, None, ok_dec, (lo, hi)))
    dist.destroy_process_group()


def test_shard_range_partitions_blocks():
    sys.path.insert(0, ROOT)
    from lz4_b200 import dist as ldist
    for n in [0, 1, 7, 8, 65536, 8191]:
        for world in [1, 2, 3, 8]:
            ranges = [ldist.shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            for a, b in zip(ranges, ranges[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_reassembly_matches_single_process(oracle):
    world, n_blocks = 2, 6
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_blocks, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[2] for r in results)                            # decoded frame identical on every rank
    sizes, frame = next((r[0], r[1]) for r in results if r[0] is not None)
    data = oracle.datagen(n_blocks * BLOCK, 0.5, 7)
    ref_sizes, ref_frame = [], b""
    for i in range(n_blocks):
        r, c = oracle.compress(data[i * BLOCK:(i + 1) * BLOCK], 1)
        ref_sizes.append(r)
        ref_frame += c
    assert sizes == ref_sizes and frame == ref_frame
