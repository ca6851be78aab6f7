import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def bmc():
    from _bmc_loader import load_pkg
    return load_pkg()


@pytest.fixture(scope="session")
def synth():
    from _bmc_loader import load_synth
    return load_synth()


@pytest.fixture(scope="session")
def orc():
    import oracle
    oracle.build()
    return oracle
