import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN, "ref_outputs.npz"))


@pytest.fixture(scope="session")
def ref_tests():
    return json.load(open(os.path.join(GOLDEN, "ref_tests.json")))


@pytest.fixture(scope="session")
def golden_data_dir():
    return os.path.join(GOLDEN, "data")


@pytest.fixture(scope="session")
def ref_lib():
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built (needs /root/reference; built in the authoring container)")
    return ref
