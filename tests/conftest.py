import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_runs():
    import numpy as np
    return np.load(os.path.join(REPO, "tests", "golden", "ref_runs.npz"))


@pytest.fixture(scope="session")
def sampler_kat():
    import numpy as np
    return np.load(os.path.join(REPO, "tests", "golden", "sampler_kat.npz"))
