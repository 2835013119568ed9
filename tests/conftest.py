import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_py
    oracle_py.build(reference=os.path.isdir("/root/reference"))
    return oracle_py


@pytest.fixture(scope="session")
def nlp():
    import nlp_b200
    nlp_b200.build.build()
    return nlp_b200
