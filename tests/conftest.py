"""pytest configuration: registers the `gpu` marker and puts the repo root on sys.path."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on CPU")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def kat(golden_dir):
    import json
    with open(os.path.join(golden_dir, "thirdparty_kat.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def fixtures(golden_dir):
    import numpy as np
    return dict(np.load(os.path.join(golden_dir, "ref_fixtures.npz")))
