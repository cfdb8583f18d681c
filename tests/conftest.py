import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    """the compiled, unmodified reference CPU path (oracle/_ref/libmgref_O0.so)"""
    from oracle.oracle import Oracle, ref_available
    if not ref_available("O0"):
        pytest.skip("oracle/_ref/libmgref_O0.so not built (needs /root/reference)")
    return Oracle("O0")


def golden(name):
    import numpy as np
    return np.load(os.path.join(GOLDEN, name))
