"""The reference's OWN fuzzer on the drop-in entry points, on a GPU.

oracle/_ref/fuzzer_b200 is tests/fuzzer.c of the reference linked so that LZ4_compress_default,
LZ4_compress_fast and LZ4_decompress_safe resolve to lz4_b200/liblz4_b200.so and everything else to the
reference (oracle/Makefile, INTEGRATION.md level 1).  `-s<seed>` skips the unit tests that pin the reference's
own parse; the fuzz loop checks, per cycle, that compressed data decodes to the input, that truncated /
overlong capacities fail exactly where the contract says, and that nothing is written past a capacity
(tests/fuzzer.c:479-727).  The binary is built where /root/reference exists and travels to the GPU box.
"""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FUZZER = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "fuzzer_b200")


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [1, 2, 2026])
def test_reference_fuzzer_on_the_drop_in_entry_points(seed):
    if not os.path.exists(FUZZER):
        pytest.skip("oracle/_ref/fuzzer_b200 not built (needs /root/reference at build time)")
    r = subprocess.run([FUZZER, "-s%d" % seed, "-i4"], capture_output=True, text=True, timeout=600)
    tail = (r.stdout + r.stderr)[-1500:]
    assert r.returncode == 0, tail
    assert "all tests completed successfully" in r.stdout + r.stderr, tail
