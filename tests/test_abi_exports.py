"""CPU-side checks of the drop-in boundary: liblz4_b200.so loads without a GPU, exports every
symbol include/lz4_b200.h declares, and its host-only entry points follow lz4.c:749-752.  No
compute call is made here."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    from lz4_b200 import build, _lib
    build.build()
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "lz4_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"LZ4B200_API[^;(]*?\b(LZ4[A-Za-z0-9_]+)\s*\(", text)))


def test_header_declares_the_scope_symbols():
    names = declared_symbols()
    for must in ["LZ4_compress_default", "LZ4_compress_fast", "LZ4_decompress_safe", "LZ4_compressBound",
                 "LZ4_compress_fast_extState", "LZ4_compress_fast_extState_fastReset",
                 "LZ4_decompress_safe_usingDict", "LZ4B200_decompress_blocks", "LZ4B200_compress_blocks"]:
        assert must in names


def test_every_declared_symbol_is_exported_and_bound(lib):
    from lz4_b200 import _lib
    bound = {p[0] for p in _lib.PROTOTYPES}
    for name in declared_symbols():
        assert hasattr(lib, name), name + " not exported by liblz4_b200.so"
        assert name in bound, name + " has no ctypes prototype"
    raw = C.CDLL(_lib.LIB_PATH)
    for name in bound:
        getattr(raw, name)


def test_host_only_entry_points(lib, oracle):
    assert lib.LZ4_versionNumber() == 11000                      # lz4.h:131-135
    assert lib.LZ4_versionString() == b"1.10.0"
    assert lib.LZ4_sizeofState() == 16416                        # lz4.h:729-733
    for n in [0, 1, 12, 13, 255, 256, 65536, 4 << 20, 0x7E000000, 0x7E000001, -1, -100]:
        assert lib.LZ4_compressBound(n) == oracle.compress_bound(n), n
    assert lib.LZ4_compressBound(65536) == 65809


def test_no_cpu_fallback_without_device(lib):
    """On a machine without a CUDA device the codec must fail loudly, not compute on the CPU."""
    if lib.LZ4B200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    from lz4_b200 import block
    r, out = block.LZ4_compress_default(b"abcdabcdabcdabcdabcdabcd")
    assert r == 0 and out == b""
    r, out = block.LZ4_decompress_safe(bytes([0x30, 0x78, 0x79, 0x7A]), 3)
    assert r < 0
    assert b"no CUDA device" in lib.LZ4B200_last_cuda_error()
    assert lib.LZ4B200_launch_count() == 0


def test_product_does_not_reference_the_oracle():
    """The product path may not import, link or call anything under oracle/."""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "lz4_b200")):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".h")):
                if "oracle" in open(os.path.join(base, f), errors="ignore").read().replace("no CPU", ""):
                    bad.append(f)
    assert not bad, bad
