import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    return np.load(ROOT / "tests" / "golden" / "hotpath_n16.npz")


@pytest.fixture(scope="session")
def port():
    from oracle import portapi
    portapi.lib()
    return portapi


@pytest.fixture(scope="session")
def ref():
    """the reference-built oracle; skipped where oracle/_ref/libthunder_ref.so does not exist"""
    from oracle import refapi
    if not refapi.available():
        pytest.skip("oracle/_ref/libthunder_ref.so not built (needs /root/reference)")
    refapi.lib()
    return refapi


@pytest.fixture(scope="session")
def ctx():
    from thunder_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()
