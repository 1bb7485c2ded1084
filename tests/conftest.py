import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


@pytest.fixture(scope="session")
def golden_layers():
    return load_golden("layers.npz")


@pytest.fixture(scope="session")
def golden_neg():
    return load_golden("neg_sampling.npz")


@pytest.fixture(scope="session")
def golden_layout():
    return load_golden("layout.npz")


@pytest.fixture(scope="session")
def golden_ablation():
    return load_golden("ablation.npz")
