"""pytest configuration: `gpu` marker + shared fixtures.

CPU suite (`-m "not gpu"`): oracle vs golden vectors / compiled reference, host logic, C-ABI export
check.  GPU suite (`-m gpu`): parity of the CUDA path (through the C-ABI) against the oracle.
"""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    from oracle.pyoracle import Reference, have_reference
    if not have_reference():
        pytest.skip("oracle/_ref not built (reference sources absent)")
    return Reference()


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def kat_decode():
    return load_golden("kat_decode.json")["cases"]


@pytest.fixture(scope="session")
def kat_compress():
    return load_golden("kat_compress.json")["cases"]


@pytest.fixture(scope="session")
def datagen_digests():
    return load_golden("datagen_digests.json")
