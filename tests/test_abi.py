"""CPU tests of the drop-in boundary: libb200stencil.so loads, exports every symbol that
include/b200_stencil.h declares, its test table agrees with the oracle's (two independent
restatements of what the reference drivers hard-code), and -- there being no GPU here -- every
compute entry point FAILS LOUDLY instead of falling back to anything."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "b200_stencil.h"


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_header_symbols_exported(pkg):
    lib = pkg.load()
    text = HEADER.read_text()
    declared = set(re.findall(r"\b(b200_[a-z_0-9]+)\s*\(", text))
    declared -= {"b200_ctx"}
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(pkg.capi.EXPORTS)
    assert lib.b200_api_version() == 1


def test_table_matches_oracle(pkg, oracle):
    assert pkg.TESTS == __import__("oracle_util").TESTS
    for t in pkg.TESTS:
        a, b = pkg.test_info(t), oracle.info(t)
        for k in ("name", "ndims", "narrays", "nscalars", "rotation"):
            assert a[k] == b[k], (t, k)
        assert pkg.capi.load().b200_test_by_name(t.encode()) == pkg.TEST_ID[t]
    assert pkg.capi.load().b200_test_by_name(b"whispering") == -1


def test_interior_points_match_baseline(pkg):
    # interior counts at 512x256x256 (2D tests 512x65536)
    want = {"laplacian": 32903160, "divergence": 32903160, "gradient": 32903160,
            "wave13pt": 32260032, "lapgsrb": 32260032, "tricubic2": 32260032,
            "uxx1": 32580581, "tricubic": 32580581, "vecadd": 33554432}
    for t, n in want.items():
        assert pkg.interior_points(t, 512, 256, 256) == n, t
    assert pkg.interior_points("gameoflife", 512, 65536, 1) == 33422340
    assert pkg.interior_points("jacobi", 512, 65536, 1) == 33422340
    assert pkg.interior_points("gaussblur", 512, 65536, 1) == 33290256
    assert pkg.interior_points("laplacian", 2, 2, 2) == 0


def test_interior_matches_oracle_writes(pkg, oracle):
    """The lo/hi interior box of the table is exactly the set of points the reference writes."""
    for t in pkg.TESTS:
        info = pkg.test_info(t)
        if t in ("vecadd", "matvec", "sincos", "matmul"):
            continue
        nx, ny, ns = (11, 9, 8) if info["ndims"] == 3 else (11, 13, 1)
        sc, arrays, _ = oracle.init(t, "double", nx, ny, ns)
        before = [a.copy() for a in arrays]
        oracle.sweep(t, "double", nx, ny, ns, sc, arrays)
        changed = np.zeros(nx * ny * ns, dtype=bool)
        for a, b in zip(arrays, before):
            changed |= a != b
        changed = changed.reshape(ns, ny, nx)
        lo, hi = info["lo"], info["hi"]
        box = np.zeros_like(changed)
        zs = slice(lo[2], ns - hi[2]) if info["ndims"] == 3 else slice(0, 1)
        box[zs, lo[1]:ny - hi[1], lo[0]:nx - hi[0]] = True
        assert np.array_equal(changed, box), t


@pytest.mark.skipif(_have_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(pkg):
    lib = pkg.load()
    n = C.c_int(-1)
    rc = lib.b200_device_count(C.byref(n))
    assert rc == 3 and n.value == 0          # B200_ERR_NO_DEVICE
    assert len(lib.b200_last_error()) > 0
    with pytest.raises(pkg.B200Error):
        pkg.Context(1)
    a = np.zeros(8 * 8 * 8)
    with pytest.raises(pkg.B200Error):
        pkg.sweep("laplacian", "double", 8, 8, 8, [0.1, 0.2], [a.ctypes.data, a.ctypes.data])


def test_bad_arguments(pkg):
    lib = pkg.load()
    assert lib.b200_get_test_info(99) is None or not lib.b200_get_test_info(99)
    assert lib.b200_sweep(None, None, None) == 1          # B200_ERR_ARG
    assert lib.b200_interior_points(99, 8, 8, 8) == 0


def test_bench_interior_points_match_library(pkg):
    """bench.py's reference arm computes interior points in Python (it must not map the product library): the table it
    uses has to agree with b200_interior_points for every stencil."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", ROOT / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for test in bench.INTERIOR:
        for dims in ((512, 256, 256), (5, 5, 5), (3, 3, 3), (130, 37, 29)):
            nd = pkg.test_info(test)["ndims"]
            nx, ny, ns = dims
            lib = pkg.interior_points(test, nx, ny * ns, 1) if nd == 2 else pkg.interior_points(test, nx, ny, ns)
            assert bench.interior_points_py(test, nx, ny, ns) == lib, (test, dims)
        assert bench.NARRAYS[test] == pkg.test_info(test)["narrays"]
