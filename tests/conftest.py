"""pytest configuration: `gpu` marker, import paths, golden-vector fixtures."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def stages():
    return np.load(os.path.join(GOLDEN, "stages.npz"))


@pytest.fixture(scope="session")
def configs_small():
    return np.load(os.path.join(GOLDEN, "configs_small.npz"))


@pytest.fixture(scope="session")
def configs_bundled():
    return np.load(os.path.join(GOLDEN, "configs_bundled.npz"))


@pytest.fixture(scope="session")
def lswarp():
    return np.load(os.path.join(GOLDEN, "configs_lswarp.npz"))


@pytest.fixture(scope="session")
def bundled_pair():
    b = np.load(os.path.join(GOLDEN, "bundled_pair.npz"))
    return b["im0"].astype(np.float32), b["im1"].astype(np.float32)


@pytest.fixture(scope="session")
def configs_extra():
    return np.load(os.path.join(GOLDEN, "configs_extra.npz"))


def _big(n):
    g = np.load(os.path.join(GOLDEN, "big_%d.npz" % n))
    return {"im0": g["im0"].astype(np.float32), "im1": g["im1"].astype(np.float32), "U": g["U"], "V": g["V"],
            "seed": int(g["seed"])}


@pytest.fixture(scope="session")
def big1024():
    """The reference's flow for one seeded synthetic 1024 x 1024 pair, full EX3 parameters (oracle/make_golden.py --big)."""
    return _big(1024)


@pytest.fixture(scope="session")
def big2048():
    return _big(2048)
