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


@pytest.fixture(scope="session")
def golden_wide():
    import numpy as np
    return np.load(os.path.join(REPO, "tests", "golden", "ref_runs_wide.npz"))


def golden(name):
    import numpy as np
    return np.load(os.path.join(REPO, "tests", "golden", name))


@pytest.fixture(scope="session", autouse=True)
def _native_libraries_built():
    """Build (or refresh) the product library and the checker libraries once per session; both are no-ops when the
    binaries that travelled with the snapshot are newer than their sources."""
    import shutil
    from odam_b200 import build
    if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
        build.build()
    from oracle import c_oracle
    c_oracle.build()
