import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Both libraries are build artefacts; build them once if missing."""
    import audiorenderingv2_b200 as arv
    import oracle
    if not os.path.exists(arv.LIB_PATH) or not os.path.exists(os.path.join(ROOT, "oracle", "_build", "liboracle.so")):
        import __graft_entry__
        __graft_entry__.build()
    oracle.lib()


@pytest.fixture(scope="session")
def golden_scenes():
    return np.load(os.path.join(GOLDEN, "scenes.npz"))


@pytest.fixture(scope="session")
def golden_receiver():
    d = np.load(os.path.join(GOLDEN, "receiver.npz"))
    return d["left"], d["right"]
