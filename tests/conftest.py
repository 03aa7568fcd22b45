"""pytest configuration: the ``gpu`` marker and shared fixture loaders.

``-m "not gpu"`` : oracle vs the golden vectors frozen from the reference, host
logic (tree flattening, sharding, gloo world_size=2), and that the C-ABI library
loads and exports every symbol of include/phynet_b200.h -- no compute calls.
``-m gpu``       : the parity tests proper; they call the CUDA path through the
C ABI and compare with the oracle / the golden vectors.
"""
import json
import os
import sys

import pytest

# read once when the device context is created: make the GPU tests exercise the chunked launch path
# with batches of a few hundred loci instead of thousands
os.environ.setdefault("PNB_CHUNK_MIN_JOBS", "128")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.join(ROOT, "tests") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "tests"))   # shared helpers (dropin_common)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def felsen_cases():
    return load_golden("felsen_cases.json")["cases"]


@pytest.fixture(scope="session")
def pmatrix_cases():
    return load_golden("pmatrix_cases.json")["cases"]


@pytest.fixture(scope="session")
def engine_case():
    return load_golden("engine_case.json")
