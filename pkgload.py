"""Import helper: the package directory is named after the reference repo
(`kernelgen-perf-tests_b200/`), which is not a valid Python identifier, so it
is registered under the importable name `kernelgen_perf_tests_b200`."""
import importlib.util
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent / "kernelgen-perf-tests_b200"
PKG_NAME = "kernelgen_perf_tests_b200"


def load_pkg():
    if PKG_NAME in sys.modules:
        return sys.modules[PKG_NAME]
    spec = importlib.util.spec_from_file_location(
        PKG_NAME, PKG_DIR / "__init__.py", submodule_search_locations=[str(PKG_DIR)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[PKG_NAME] = mod
    spec.loader.exec_module(mod)
    return mod
