import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The product library and the C oracle must exist; build them if a fresh checkout."""
    import subprocess
    from hypergef_b200 import _native
    if not os.path.exists(_native.LIB_PATH):
        subprocess.run(["make", "-s", "-j8", "-C", os.path.join(ROOT, "hypergef_b200", "csrc")], check=True)
    from oracle import oracle as orc
    orc.lib()


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def balancer_cases():
    d = load_golden("balancer")
    names = sorted({k.split("__")[0] for k in d.files})
    return d, names


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
