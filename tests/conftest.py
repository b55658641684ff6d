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
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def kernel_vectors():
    z = np.load(os.path.join(GOLDEN, "kernel_vectors.npz"))
    meta = json.loads(str(z["meta/scalars"]))
    return z, meta


@pytest.fixture(scope="session")
def ls_vectors():
    z = np.load(os.path.join(GOLDEN, "ls_vectors.npz"))
    return z, json.loads(str(z["cases"]))


@pytest.fixture(scope="session")
def price_goldens():
    with open(os.path.join(GOLDEN, "price_goldens.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def levy_goldens():
    with open(os.path.join(GOLDEN, "levy_goldens.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def dupire_vectors():
    z = np.load(os.path.join(GOLDEN, "dupire_vectors.npz"))
    return z, json.loads(str(z["meta"]))


@pytest.fixture(scope="session")
def american_goldens():
    with open(os.path.join(GOLDEN, "american_goldens.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def european_goldens():
    with open(os.path.join(GOLDEN, "european_goldens.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def greek_goldens():
    with open(os.path.join(GOLDEN, "greek_goldens.json")) as f:
        return json.load(f)
