import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    from kzg_testlib import load_golden
    return load_golden()


@pytest.fixture(scope="session")
def setup_bytes():
    from kzg_testlib import SETUP
    with open(SETUP, "rb") as f:
        return f.read()


@pytest.fixture(scope="session")
def ref(setup_bytes):
    """C restatement of the reference path (oracle/kzg_ref.c)."""
    import kzg_ref
    return kzg_ref.RefSettings(setup_bytes)


@pytest.fixture(scope="session")
def pyoracle(setup_bytes):
    import kzg_oracle
    return kzg_oracle, kzg_oracle.load_settings(setup_bytes)


@pytest.fixture(scope="session")
def gpu_settings():
    """One GPU context for the whole session.  Window 8 keeps the table at 1.6 GB so the
    suite starts in seconds; test_gpu_parity.py::test_large_window covers c = 15."""
    import raiko_b200 as rk
    wb = int(os.environ.get("RAIKO_KZG_TEST_WINDOW_BITS", "8"))
    s = rk.KzgSettings(window_bits=wb)
    yield s
    s.close()
