"""pytest configuration: the ``gpu`` marker and shared fixtures.

``-m "not gpu"`` runs on a CPU-only box: the oracle against the golden vectors, the host logic,
and the C-ABI load/export check.  ``-m gpu`` are the parity tests proper (CUDA path vs oracle /
golden vectors through the C-ABI) and need a B200.
"""
import gzip
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


def load_golden(name):
    with gzip.open(os.path.join(GOLDEN, name), "rb") as f:
        return json.loads(f.read().decode())


@pytest.fixture(scope="session")
def kat():
    return load_golden("kat.json.gz")


@pytest.fixture(scope="session")
def golden_games():
    return load_golden("games.json.gz")


@pytest.fixture(scope="session")
def golden_probe():
    return load_golden("probe.json.gz")


@pytest.fixture(scope="session")
def oracle():
    from oracle import lib
    lib.build()
    return lib
