import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    gdir = os.path.join(REPO, "tests", "golden")
    return {name: np.load(os.path.join(gdir, name + ".npz")) for name in ("operators", "losses", "detect",
                                                                           "priors_meta", "map", "extras", "crop")}
