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


def _gpu_hardware_present() -> bool:
    """A CUDA device node exists.  Deliberately NOT a check of the product library: on a GPU box a library that fails to
    load must FAIL the gpu tests loudly, never skip them."""
    import glob
    return bool(glob.glob("/dev/nvidia[0-9]*"))


def pytest_collection_modifyitems(config, items):
    if _gpu_hardware_present():
        return
    skip = pytest.mark.skip(reason="no NVIDIA device on this machine (gpu tests run on the B200 box: pytest -m gpu)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
