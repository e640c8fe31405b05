import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    import plsvo_b200

    return plsvo_b200


@pytest.fixture(scope="session")
def abi(pkg):
    return pkg.abi


@pytest.fixture(scope="session")
def synth(pkg):
    return pkg.synth


@pytest.fixture(scope="session")
def oracle(abi):
    import oracle_lib

    oracle_lib.build()
    oracle_lib.load(abi)
    return oracle_lib


@pytest.fixture(scope="session")
def gen_device():
    import torch

    return "cuda" if torch.cuda.is_available() else "cpu"
