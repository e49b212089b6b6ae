import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle_lib import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reflib():
    from oracle_lib import RefLib
    if not RefLib.available():
        pytest.skip("oracle/_ref/libref.so not built (reference tree absent)")
    return RefLib()


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def ref_vectors(golden_dir):
    import json
    with open(os.path.join(golden_dir, "ref_vectors.json")) as f:
        return json.load(f)
