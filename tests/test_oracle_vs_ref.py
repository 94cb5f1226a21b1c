"""Differential test: the oracle restatement vs the UNMODIFIED reference compiled into
oracle/_ref (skipped on machines where the reference has not been built)."""
import numpy as np


def test_datagen_identical(oracle, reference):
    for (n, p, s) in [(65536, 0.5, 0), (1 << 20, 0.9, 3), (100000, 0.0, 7), (300000, 1.0, 2), (1, 0.5, 0),
                      (777777, 0.25, 99)]:
        assert (oracle.datagen(n, p, s) == reference.datagen(n, p, s)).all()


def test_compress_and_decode_identical(oracle, reference):
    rng = np.random.default_rng(1)
    sizes = [0, 1, 5, 12, 13, 14, 20, 64, 100, 1000, 4096, 65535, 65536, 65546, 65547, 65548, 70000, 200000, 1 << 20]
    for trial in range(250):
        n = int(rng.choice(sizes))
        p = float(rng.choice([0.0, 0.1, 0.5, 0.9, 1.0]))
        d = oracle.datagen(n, p, trial)
        acc = int(rng.choice([1, 1, 1, 2, 8, 32, 1000, 65537, -3]))
        ro, bo = oracle.compress(d, acc)
        rr, br = reference.compress(d, acc)
        assert (ro, bo) == (rr, br), (n, p, acc)
        for cap in [rr, rr - 1, rr - 7, max(rr // 2, 0), 1, 0]:
            a = oracle.compress(d, acc, cap)
            b = reference.compress(d, acc, cap)
            assert a[0] == b[0] and (a[0] == 0 or a[1] == b[1]), (n, p, acc, cap)
        for cap in [n, n + 1, n + 100, n - 1, n - 10, n // 2]:
            if cap < 0:
                continue
            a = oracle.decompress(br, cap)
            b = reference.decompress(br, cap)
            assert a[0] == b[0] and (a[0] < 0 or a[1] == b[1]), (n, p, acc, cap)


def test_noisy_source_identical(oracle, reference):
    """tests/fuzzer.c:588-622 idea: corrupted blocks must give the same verdict, return value and
    bytes as the reference decoder (x86-64 build)."""
    rng = np.random.default_rng(2)
    accepted = 0
    for trial in range(1500):
        n = int(rng.choice([20, 64, 100, 300, 1000, 5000, 70000]))
        d = oracle.datagen(n, float(rng.choice([0.1, 0.5, 0.9])), 1000 + trial)
        _, br = reference.compress(d, int(rng.choice([1, 4])))
        b = bytearray(br)
        for _ in range(int(rng.integers(1, 6))):
            mode = rng.integers(0, 4)
            pos = int(rng.integers(0, len(b)))
            if mode == 0:
                b[pos] = int(rng.integers(0, 256))
            elif mode == 1:
                b[pos] = int(rng.choice([0, 0xFF, 0xF0, 0x0F, 0x10, 0x1F]))
            elif mode == 2:
                del b[pos:pos + int(rng.integers(1, 4))]
            else:
                b[pos:pos] = bytes(rng.integers(0, 256, int(rng.integers(1, 4)), dtype=np.uint8))
        if not b:
            continue
        for cap in [n, n + int(rng.integers(0, 80)), max(n - int(rng.integers(0, 80)), 0)]:
            a = oracle.decompress(bytes(b), cap)
            c = reference.decompress(bytes(b), cap)
            accepted += c[0] >= 0
            assert a[0] == c[0] and (a[0] < 0 or a[1] == c[1]), (trial, n, cap)
    assert accepted > 100
