import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
# the peer-exchange tests run several "ranks" on ONE device, each on its own stream, whose kernels wait for each other:
# the streams must not share a hardware queue (must be set before CUDA initialises)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


@pytest.fixture(scope="session")
def golden_loader():
    return golden


def hits_to_lists(counts, nz):
    """(nl,) counts + (H,2) [line, triplet] rows -> list of ascending triplet index lists."""
    out = [[] for _ in range(len(counts))]
    for l, f in nz:
        out[int(l)].append(int(f))
    return out
