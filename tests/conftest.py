import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) device; run with -m gpu on the GPU box")


def rel_err(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="session")
def labels28():
    # conf/conf.yaml:12-13 (28 inline labels; blank = index 28)
    return [" ", "'"] + [chr(ord("a") + i) for i in range(26)]
