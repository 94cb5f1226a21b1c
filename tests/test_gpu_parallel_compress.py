"""GPU tests of the parallel-parse compressor (LZ4B200_compress_blocks_parallel, lz4_encode_par.cuh).

Its contract (include/lz4_b200.h): every block is a valid LZ4 block -- the ORACLE's LZ4_decompress_safe (and the
compiled reference where present) expands it to the input -- the output is deterministic, never exceeds the
capacity, returns 0 exactly when it does not fit, and the ratio stays within 2 % of the reference's at
acceleration 1 on the BASELINE generators (tests/datagen P50 / P90); acceleration is monotone.
The traps of tests/fuzzer.c:698-727 apply: last 5 bytes literal, no match start after n-12, offsets 1..65535.
"""
import numpy as np
import pytest
import torch

from gpu_util import to_dev

pytestmark = pytest.mark.gpu
BS = 65536


@pytest.fixture(scope="module")
def lib():
    from lz4_b200 import _lib
    lib = _lib.load()
    assert lib.LZ4B200_device_count() > 0, "GPU tests need a CUDA device"
    return lib


def gpu_compress(data, block=BS, accel=1, cap=None):
    from lz4_b200 import batch
    src = to_dev(np.frombuffer(data, dtype=np.uint8).copy() if isinstance(data, (bytes, bytearray)) else data)
    slots, sizes, stride = batch.compress_blocks(src, block, accel, mode="parallel", slot_capacity=cap)
    torch.cuda.synchronize()
    return slots.cpu().numpy(), sizes.cpu().numpy(), stride


def check_blocks(oracle, data, host, sizes, stride, block=BS):
    """every block decodes (oracle decoder, exact capacity) to its input; returns total compressed bytes"""
    raw = bytes(data)
    for i, n in enumerate(sizes):
        want = raw[i * block:(i + 1) * block]
        assert n > 0, (i, n)
        comp = host[i * stride:i * stride + n].tobytes()
        ret, out = oracle.decompress(comp, len(want))
        assert ret == len(want) and out == want, (i, len(want), ret)
        if len(want) >= 13:                                   # format rules a conformant encoder obeys (lz4.c:963-964)
            ret2, _ = oracle.decompress(comp, len(want) + 64)
            assert ret2 == len(want)
    return int(sizes.sum())


def test_round_trip_ratio_and_determinism_on_baseline_generators(lib, oracle):
    for proba, seed, ref_lo in ((0.5, 0, None), (0.9, 3, None), (0.2, 5, None)):
        n_blocks = 256
        d = oracle.datagen_mt(n_blocks * BS, 4 << 20, proba, seed)
        host, sizes, stride = gpu_compress(d)
        total = check_blocks(oracle, d.tobytes(), host, sizes, stride)
        ref = sum(oracle.compress(d[i * BS:(i + 1) * BS], 1)[0] for i in range(n_blocks))
        ratio, ref_ratio = n_blocks * BS / total, n_blocks * BS / ref
        assert ratio >= 0.98 * ref_ratio, (proba, ratio, ref_ratio)
        host2, sizes2, _ = gpu_compress(d)                    # deterministic: same bytes on a second run
        assert (sizes == sizes2).all()
        for i in (0, 1, 77, n_blocks - 1):
            assert (host[i * stride:i * stride + sizes[i]] == host2[i * stride:i * stride + sizes2[i]]).all()


def test_sizes_around_every_boundary(lib, oracle):
    rng = np.random.default_rng(3)
    words = [bytes(rng.integers(97, 123, int(rng.integers(2, 9)), dtype=np.uint8)) for _ in range(200)]
    text = b" ".join(words[int(i)] for i in rng.integers(0, 200, 14000))
    for n in (1, 4, 11, 12, 13, 14, 15, 16, 17, 31, 64, 100, 255, 4095, 4096, 4097, 4107, 4108, 4109, 8191, 8192, 8200,
              20000, 65523, 65524, 65525, 65535, 65536):
        for maker in (lambda k: oracle.datagen(k, 0.5, k).tobytes(), lambda k: text[:k], lambda k: b"\0" * k,
                      lambda k: bytes(rng.integers(0, 256, k, dtype=np.uint8)), lambda k: (b"abcdefg" * 10000)[:k]):
            raw = maker(n)
            host, sizes, stride = gpu_compress(raw, block=n)
            assert len(sizes) == 1 and 0 < sizes[0] <= oracle.compress_bound(n)
            ret, out = oracle.decompress(host[:sizes[0]].tobytes(), n)
            assert ret == n and out == raw, (n, ret)


def test_special_shapes_long_runs_and_long_literals(lib, oracle):
    rng = np.random.default_rng(9)
    shapes = [b"\0" * BS, b"ab" * (BS // 2), bytes(rng.integers(0, 256, BS, dtype=np.uint8)),
              bytes(rng.integers(0, 256, 30000, dtype=np.uint8)) + b"\0" * (BS - 30000),
              (bytes(rng.integers(0, 256, 3000, dtype=np.uint8)) + b"x" * 500) * 18 + b"y" * (BS - 18 * 3500),
              b"".join(bytes([i & 255]) * (1 + i % 300) for i in range(2000))[:BS].ljust(BS, b"z"),
              oracle.datagen(BS, 1.0, 1).tobytes(), oracle.datagen(BS, 0.99, 2).tobytes()]
    data = b"".join(shapes)
    host, sizes, stride = gpu_compress(data)
    check_blocks(oracle, data, host, sizes, stride)
    # periodic data must compress about as well as the reference does (long self-overlapping matches survive the windows)
    for i in (0, 1):
        assert sizes[i] <= 2 * oracle.compress(shapes[i], 1)[0] + 64


def test_limited_output_and_never_past_capacity(lib, oracle):
    from lz4_b200 import batch
    d = oracle.datagen_mt(8 * BS, 1 << 20, 0.5, 21)
    _, full, _ = gpu_compress(d)
    src = to_dev(d)
    for cap in (int(full.max()), int(full.max()) - 1, int(full.min()), int(full.min()) - 1, 1000, 1):
        stride = (cap + 15) // 16 * 16 + 64
        slots = torch.full((8 * stride,), 0xA5, dtype=torch.uint8, device=src.device)
        out_sizes = torch.zeros(8, dtype=torch.int32, device=src.device)
        batch.compress_blocks(src, BS, 1, slots=slots, slot_stride=stride, slot_capacity=cap, out_sizes=out_sizes, mode="parallel")
        torch.cuda.synchronize()
        host, sz = slots.cpu().numpy(), out_sizes.cpu().numpy()
        for i in range(8):
            assert (host[i * stride + cap:(i + 1) * stride] == 0xA5).all(), "wrote past dstCapacity"
            if full[i] <= cap:
                assert sz[i] == full[i]
                ret, out = oracle.decompress(host[i * stride:i * stride + sz[i]].tobytes(), BS)
                assert ret == BS and out == d[i * BS:(i + 1) * BS].tobytes()
            else:
                assert sz[i] == 0


def test_acceleration_is_monotone(lib, oracle):
    d = oracle.datagen_mt(64 * BS, 4 << 20, 0.5, 33)
    prev = None
    for accel in (1, 4, 5, 8, 16, 32, 64, 1000, 65537, 1 << 30):
        host, sizes, stride = gpu_compress(d, accel=accel)
        total = check_blocks(oracle, d.tobytes()[:4 * BS], host, sizes[:4], stride)
        tot = int(sizes.sum())
        if prev is not None:
            assert tot >= prev * 0.999, (accel, tot, prev)     # less search, never a (noticeably) better ratio
        prev = tot
    host0, sizes0, _ = gpu_compress(d, accel=0)                  # values < 1 behave as 1 (lz4.c:1386)
    host1, sizes1, _ = gpu_compress(d, accel=1)
    assert (sizes0 == sizes1).all()


def test_ragged_last_block_and_reference_decoder(lib, oracle):
    from oracle.pyoracle import Reference, have_reference
    total = 5 * BS + 12345
    d = oracle.datagen_mt(total, 1 << 20, 0.9, 44)
    host, sizes, stride = gpu_compress(d)
    assert len(sizes) == 6
    check_blocks(oracle, d.tobytes(), host, sizes, stride)
    if have_reference():
        ref = Reference()
        for i, n in enumerate(sizes):
            want = d[i * BS:(i + 1) * BS].tobytes()
            ret, out = ref.decompress(host[i * stride:i * stride + n].tobytes(), len(want))
            assert ret == len(want) and out == want
