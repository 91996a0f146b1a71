import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    from oracle_lib import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def pkg():
    import bra_pkg
    return bra_pkg.load()


@pytest.fixture(scope="session")
def dropin(pkg):
    """The B200 library behind the reference's per-stage C API (same wrapper class as for the reference itself)."""
    from oracle_lib import RefApi
    pkg.lib()  # raises if the library was not built
    return RefApi(pkg.LIB_PATH)


@pytest.fixture(scope="session")
def vocab(golden):
    return golden["vocab"]
