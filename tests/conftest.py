import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    return np.load(os.path.join(REPO, "tests", "golden", "reference_golden.npz"))


@pytest.fixture(scope="session")
def loop_golden():
    import numpy as np

    return np.load(os.path.join(REPO, "tests", "golden", "reference_loop_golden.npz"))
