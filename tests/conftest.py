import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def cameras():
    """Sample cameras: parameter values of the reference's samples/*.yaml and in-file test models."""
    return load_golden("cameras.json")


@pytest.fixture(scope="session")
def O():
    from oracle import oracle
    oracle.build()
    return oracle


def oracle_model(O, cam):
    return O.make_model(cam["model_id"], cam["params"], cam["width"], cam["height"])
