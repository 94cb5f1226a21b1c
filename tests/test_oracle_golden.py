"""The oracle (oracle/lz4_oracle.c, datagen_oracle.c) against the committed golden vectors, which
are outputs of the compiled reference (tests/golden/make_golden.py).  Runs without the reference."""
import hashlib
import os

import numpy as np

from conftest import GOLDEN


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


def test_decode_known_answers(oracle, kat_decode):
    assert len(kat_decode) > 1000
    for c in kat_decode:
        ret, out = oracle.decompress(bytes.fromhex(c["block"]), c["cap"])
        assert ret == c["ret"], c
        if ret >= 0:
            assert out.hex() == c["out"], c


def test_compress_known_answers(oracle, kat_compress):
    for c in kat_compress:
        ret, out = oracle.compress(bytes.fromhex(c["src"]), c["accel"], c.get("cap"))
        assert ret == c["ret"], c
        if ret > 0:
            assert out.hex() == c["out"], c


def test_datagen_and_compress_digests(oracle, datagen_digests):
    cache = {}
    for row in datagen_digests["buffers"]:
        key = (row["size"], row["proba"], row["seed"])
        if key not in cache:
            cache[key] = oracle.datagen(*key)
            assert sha(cache[key]) == row["src_sha256"], key
        ret, out = oracle.compress(cache[key], row["accel"])
        assert ret == row["csize"], row
        assert sha(out) == row["comp_sha256"], row
        r2, back = oracle.decompress(out, row["size"])
        assert r2 == row["size"] and back == cache[key].tobytes()


def test_stream_of_64k_blocks(oracle, datagen_digests):
    s = datagen_digests["stream"]
    d = oracle.datagen(s["size"], s["proba"], s["seed"])
    assert sha(d) == s["src_sha256"]
    h = hashlib.sha256()
    for i, off in enumerate(range(0, len(d), s["block"])):
        ret, out = oracle.compress(d[off:off + s["block"]], s["accel"])
        assert ret == s["csizes"][i]
        h.update(out)
    assert h.hexdigest() == s["stream_sha256"]


def test_binary_fixture_block(oracle):
    with open(os.path.join(GOLDEN, "p50_seed0_64k.lz4block"), "rb") as f:
        blk = f.read()
    d = oracle.datagen(65536, 0.5, 0)
    ret, out = oracle.decompress(blk, 65536)
    assert ret == 65536 and out == d.tobytes()
    r2, c2 = oracle.compress(d, 1)
    assert c2 == blk
    st = oracle.block_stats(blk)
    assert st["literal_bytes"] + st["match_bytes"] == 65536


def test_fuzzer_unit_properties(oracle):
    """Block-API properties of tests/fuzzer.c:479-486, 502-504, 545-586, 627-639, 698-727."""
    rng = np.random.default_rng(5)
    for trial in range(40):
        n = int(rng.integers(1, 131072))
        d = oracle.datagen(n, float(rng.choice([0.1, 0.5, 0.9])), trial)
        bound = oracle.compress_bound(n)
        r, c = oracle.compress(d, 1)
        assert 0 < r <= bound
        assert oracle.compress(d, 1, r)[0] == r                  # exact capacity succeeds
        miss = int(rng.integers(1, 64))
        assert oracle.compress(d, 1, r - miss)[0] == 0           # too small -> 0
        r8, _ = oracle.compress(d, 8)
        assert r8 > 0 and oracle.compress(d, 8, r8 - 1)[0] == 0
        assert oracle.decompress(c, n) == (n, d.tobytes())
        assert oracle.decompress(c, n + 1) == (n, d.tobytes())
        assert oracle.decompress(c, n - 1)[0] < 0
        if n > 10:
            assert oracle.decompress(c, n - 10)[0] < 0
        assert oracle.decompress(c[:-1], n)[0] < 0
        assert oracle.decompress(c + b"\x00", n)[0] < 0
    assert oracle.compress(b"", 1, oracle.compress_bound(0)) == (1, b"\x00")   # fuzzer.c:1124-1131
    assert oracle.compress(b"", 1, 0)[0] == 0                                  # fuzzer.c:1134-1139
