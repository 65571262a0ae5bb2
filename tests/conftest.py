import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    from pkgload import load_pkg
    return load_pkg()


@pytest.fixture(scope="session")
def oracle():
    from oracle_util import Oracle
    return Oracle("fast")


@pytest.fixture(scope="session")
def oracle_strict():
    from oracle_util import Oracle
    return Oracle("strict")
