import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """Make sure both shared libraries exist (building is cheap and needs no GPU)."""
    import spand_public_b200 as S
    import oracle_lib as O
    if not os.path.exists(S.LIB_PATH):
        S.build()
    if not os.path.exists(O._LIB):
        O.build()
    yield
