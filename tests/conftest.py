import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the oracle is test infrastructure: make sure its C library exists
    so = os.path.join(ROOT, "oracle", "libpsi_oracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def small_model():
    from psi_release_b200 import synthetic
    return synthetic.make_smplx_model(seed=1234, num_verts=431)


@pytest.fixture(scope="session")
def full_model():
    from psi_release_b200 import synthetic
    return synthetic.make_smplx_model(seed=1234, num_verts=10475)
