import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from lvt_b200 import capi  # noqa: E402

ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "liblvt_oracle.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure): built on demand from oracle/."""
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    return capi.Library(ORACLE_SO)


@pytest.fixture(scope="session")
def cuda():
    """The product library.  Must exist and must see a GPU -- no fallback."""
    import lvt_b200
    return lvt_b200.load()


@pytest.fixture(scope="session")
def rng():
    return np.random.default_rng(1234)
