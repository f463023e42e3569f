import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def rfc_vectors():
    return json.load(open(os.path.join(GOLDEN, "rfc9496_vectors.json")))


@pytest.fixture(scope="session")
def sodium_vectors():
    return json.load(open(os.path.join(GOLDEN, "libsodium_vectors.json")))


@pytest.fixture(scope="session")
def c_oracle():
    from oracle import c_oracle as co
    co.load()
    return co


@pytest.fixture(scope="session")
def ctx():
    """A CUDA context through the C ABI.  No skip-on-failure: a GPU test without the CUDA library must fail."""
    import zkvm_b200 as zk
    c = zk.Context(0)
    yield c
    c.close()
