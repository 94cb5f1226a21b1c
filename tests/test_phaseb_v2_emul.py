"""CPU check of the experimental phase B variant (lz4_b200/csrc/lz4_phaseb_v2.h, not in the default build).

The header is plain C++: the same text is compiled for the device (-DLZ4K_PHASEB_V2) and, here, by g++
into an emulator (tests/emul/phaseb_emul.cpp) that replays the kernel's loop lane by lane -- ballots,
dynamic hand-out, done flags -- with shuffled warp / lane order, on reference-compressed blocks, and
compares the assembled window with the decoded block.  This pins the index arithmetic of the variant
before it ever runs on a GPU; it says nothing about its speed.
"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle.pyoracle import Oracle, Reference, have_reference

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("g++ not available")
    so = str(tmp_path_factory.mktemp("pbemul") / "libpbemul.so")
    subprocess.run([gxx, "-O2", "-std=c++17", "-Wall", "-shared", "-fPIC", "-o", so,
                    os.path.join(HERE, "emul", "phaseb_emul.cpp")], check=True)
    lib = C.CDLL(so)
    lib.pb_emulate.restype = C.c_int
    lib.pb_emulate.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_uint32, C.POINTER(C.c_longlong)]
    return lib


def run(lib, comp, raw, head, seed):
    stats = (C.c_longlong * 4)()
    rc = lib.pb_emulate(bytes(comp), len(comp), bytes(raw), len(raw), head, seed, stats)
    return rc, list(stats)


def blocks():
    codec = Reference() if have_reference() else Oracle()
    out = []
    for proba, seed in ((0.5, 0), (0.9, 1), (0.2, 2), (0.99, 3), (1.0, 4)):
        data = codec.datagen(3 * 65536, proba, seed)
        for k in range(3):
            out.append(("P%g/%d" % (proba, k), bytes(data[k * 65536:(k + 1) * 65536])))
    rng = np.random.default_rng(7)
    # short periods (offsets 1..16) with long and short matches, ragged sizes
    for period in list(range(1, 17)) + [31, 33, 255]:
        seedb = bytes(rng.integers(0, 256, period, dtype=np.uint8))
        body = (seedb * (70000 // period + 1))[:int(rng.integers(3000, 65536))]
        out.append(("period%d" % period, body))
    words = [bytes(rng.integers(97, 123, int(rng.integers(2, 9)), dtype=np.uint8)) for _ in range(300)]
    text = b" ".join(words[int(i)] for i in rng.integers(0, 300, 14000))[:65536]
    out.append(("text", text))
    out.append(("text24k", text[:24000]))
    out.append(("tiny", b"abcabcabcabcabcabcabcabcabc" * 3))
    out.append(("mixed", bytes(rng.integers(0, 256, 5000, dtype=np.uint8)) + b"\x00" * 20000 + text[:20000] + b"ab" * 5000))
    return codec, out


def test_uniform_body_reassembles_reference_blocks(emul):
    codec, cases = blocks()
    base = tot = 0
    for name, raw in cases:
        ret, comp = codec.compress(raw, 1)
        comp = bytes(comp)
        if ret <= 0 or len(comp) > 65535:
            continue
        for head, seed in ((0, 0), (5, 1), (15, 2), (8, 3)):
            rc, st = run(emul, comp, raw, head, seed)
            if rc == -4:                      # more than 8192 sequences: not a fast-path block
                break
            assert rc == 0, (name, head, seed, rc, st)
            tot += 1
        else:
            base += 1
    assert base >= 30 and tot >= 120


def test_statistics_match_the_shipped_loop(emul):
    """Same work distribution as the shipped loop: the lock-step run (seed 0) reproduces the GPU-measured
    loop statistics (~740 warp-iterations per P50 block, ~30 % blocked)."""
    codec = Reference() if have_reference() else Oracle()
    data = codec.datagen(5 * 65536, 0.5, 0)
    its, blk, lanes = [], [], []
    for k in range(1, 5):
        raw = bytes(data[k * 65536:(k + 1) * 65536])
        _, comp = codec.compress(raw, 1)
        rc, st = run(emul, bytes(comp), raw, 0, 0)
        assert rc == 0
        its.append(st[0]); lanes.append(st[1]); blk.append(st[2])
    assert 550 < np.mean(its) < 900
    assert 0.15 < np.sum(blk) / np.sum(lanes) < 0.45
