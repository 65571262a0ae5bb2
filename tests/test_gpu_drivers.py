"""GPU tests of the drop-in drivers (kernelgen-perf-tests_b200/drivers/bin/<test>_<real>): the
binaries a `b200` target directory of the suite builds.  Their stdout is parsed with the grammar
`benchmark` uses (benchmark:146-159,174-178,210-216,258-265 of the reference, restated here) and
the checksums are compared with the reference driver restated by the oracle -- which is itself
pinned to the README golden table (tests/test_oracle.py)."""
import os
import re
import subprocess
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
BIN = ROOT / "kernelgen-perf-tests_b200" / "drivers" / "bin"
NUM = r"[-+]?[0-9]*\.?[0-9]+(?:[eE][-+]?[0-9]+)?"
TESTS = ["laplacian", "wave13pt", "divergence", "gradient", "uxx1", "lapgsrb", "jacobi",
         "gaussblur", "gameoflife", "tricubic", "tricubic2", "vecadd", "matvec", "sincos", "matmul"]
TWO_D = {"jacobi", "gaussblur", "gameoflife", "matvec"}


def parse_like_benchmark(out: str, kernel: str) -> dict:
    """The fields `benchmark` extracts; a missing required line is what it reports as FAIL."""
    pats = {"i_mean": rf"initial mean = ({NUM})", "t_init": rf"init time = ({NUM}) sec",
            "t_alloc": rf"device buffer alloc time = ({NUM}) sec", "t_load": rf"data load time = ({NUM}) sec",
            "t_comp": rf"compute time = ({NUM}) sec", "t_save": rf"data save time = ({NUM}) sec",
            "t_free": rf"device buffer free time = ({NUM}) sec", "f_mean": rf"final mean = ({NUM})\s",
            "t_krn": rf"{kernel} kernel time = ({NUM})", "nreg_krn": rf"{kernel} regcount = (\d+)"}
    res = {}
    for k, p in pats.items():
        vals = [float(m) for m in re.findall(p, out)]
        res[k] = sum(vals) / len(vals) if vals else None        # benchmark averages all matches
    return res


def run_driver(test, real, args, env=None):
    exe = BIN / f"{test}_{real}"
    if not exe.exists():
        subprocess.run(["make", "-s", "-C", str(BIN.parent)], check=True)
    e = dict(os.environ)
    e["PROFILING_FNAME"] = test           # what benchmark passes from <test>/<target>/kernel
    e.update(env or {})
    p = subprocess.run([str(exe)] + [str(a) for a in args], capture_output=True, text=True, env=e, timeout=600)
    assert p.returncode == 0, p.stderr[-500:]
    return p.stdout


@pytest.mark.parametrize("real", ["float", "double"])
@pytest.mark.parametrize("test", TESTS)
def test_driver_contract_small(oracle, test, real):
    dims = (64, 1024) if test in TWO_D else (64, 32, 32)
    nt = 3
    out = run_driver(test, real, list(dims) + [nt])
    r = parse_like_benchmark(out, test)
    for k in ("i_mean", "t_comp", "f_mean"):                 # benchmark: missing -> FAIL
        assert r[k] is not None, f"{k} line missing:\n{out}"
    for k in ("t_init", "t_alloc", "t_load", "t_save", "t_free", "t_krn", "nreg_krn"):
        assert r[k] is not None, f"{k} line missing:\n{out}"
    nx, ny = dims[0], dims[1]
    ns = dims[2] if len(dims) == 3 else 1
    sc, im, fm = oracle.driver(test, real, nx, ny, ns, nt)
    assert "%f" % im == "%f" % r["i_mean"], out
    tol = 2e-6 if real == "double" else 5e-5
    assert abs(fm - r["f_mean"]) <= tol + abs(fm) * (1e-4 if real == "float" else 0), (fm, r["f_mean"])
    # coefficient header line is part of the cross-target diff-able output
    first = out.splitlines()[0]
    if sc and oracle.info(test)["nscalars"]:
        assert "%f" % sc[0] in first


@pytest.mark.parametrize("test,real", [("laplacian", "double"), ("laplacian", "float"), ("uxx1", "float"),
                                       ("gaussblur", "float"), ("matvec", "double"), ("tricubic", "double")])
def test_parallel_init_same_inputs(test, real):
    """B200_INIT_THREADS=N fills the host arrays with N threads in the reference's rand() draw order
    (drivers/kg_rand.h): the inputs are bit-identical, hence the very same final mean; the initial mean may move
    in its last digits (a `real` sum re-associated into per-thread partial sums)."""
    dims = (96, 2048) if test in TWO_D else (96, 40, 48)
    serial = parse_like_benchmark(run_driver(test, real, list(dims) + [3]), test)
    par = parse_like_benchmark(run_driver(test, real, list(dims) + [3], env={"B200_INIT_THREADS": "8"}), test)
    assert "%f" % serial["f_mean"] == "%f" % par["f_mean"], (serial, par)
    assert abs(serial["i_mean"] - par["i_mean"]) <= (2e-6 if real == "double" else 1e-4)


def test_pinned_host_arrays_same_result():
    """Page-locked host arrays (the default) vs B200_PINNED_HOST=0 (memalign, as in the reference): same means, every
    timing line still there, init time reported in both."""
    a = parse_like_benchmark(run_driver("wave13pt", "double", [96, 40, 48, 3], env={"B200_PINNED_HOST": "0"}), "wave13pt")
    b = parse_like_benchmark(run_driver("wave13pt", "double", [96, 40, 48, 3]), "wave13pt")
    assert a["t_init"] is not None and b["t_init"] is not None and b["t_init"] > 0
    assert "%f" % a["i_mean"] == "%f" % b["i_mean"] and "%f" % a["f_mean"] == "%f" % b["f_mean"]
    assert b["t_load"] is not None and b["t_save"] is not None


@pytest.mark.parametrize("test,dims", [("wave13pt", [130, 40, 48, 5]), ("gradient", [96, 40, 48, 2]), ("jacobi", [1024, 300, 7]),
                                       ("tricubic", [64, 24, 40, 3]), ("matmul", [132, 64, 140, 2])])
def test_verify_mode(test, dims):
    """B200_VERIFY=1 (SURVEY 8b env row): the driver runs the whole job a second time on an independent context and compares
    the results bit for bit; same stdout grammar otherwise."""
    out = run_driver(test, "double", dims, env={"B200_VERIFY": "1"})
    assert "b200 verify: second independent run bit-identical" in out, out
    r = parse_like_benchmark(out, test)
    assert r["f_mean"] is not None and r["t_comp"] is not None


def test_readme_checksums_laplacian_wave13pt():
    """README.md:119,127 -- `./laplacian 512 256 256 10` and wave13pt, double."""
    for test, (gi, gf) in {"laplacian": (0.000041, 0.000011), "wave13pt": (0.000024, 0.000173)}.items():
        r = parse_like_benchmark(run_driver(test, "double", [512, 256, 256, 10]), test)
        assert abs(r["i_mean"] - gi) < 0.6e-6 and abs(r["f_mean"] - gf) < 0.6e-6, (test, r)
        assert 0 < r["t_krn"] < r["t_comp"]


def test_no_timing_and_usage():
    out = run_driver("laplacian", "double", [16, 16, 16, 2], env={"NO_TIMING": "1"})
    assert "compute time" not in out and "initial mean" not in out and "final mean" in out
    out = run_driver("wave13pt", "double", [16, 16, 16, 2], env={"NO_TIMING": "1"})
    assert "initial mean" in out                      # wave13pt.c:756 prints it unconditionally
    p = subprocess.run([str(BIN / "laplacian_double"), "8", "8"], capture_output=True, text=True)
    assert p.returncode == 1 and p.stdout.startswith("Usage:")
    p = subprocess.run([str(BIN / "jacobi_double"), "8", "-3", "1"], capture_output=True, text=True)
    assert p.returncode == 1 and "Value for ny is invalid: -3" in p.stdout


def test_multi_gpu_driver_matches_single():
    """B200_NGPUS=2: z-slabs with the fused halo push; checksums identical to one GPU."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    for test, dims in (("laplacian", (130, 40, 50)), ("wave13pt", (128, 36, 64)), ("jacobi", (128, 700)),
                       ("uxx1", (64, 32, 40)), ("tricubic", (64, 24, 40)), ("gaussblur", (130, 300))):
        a = parse_like_benchmark(run_driver(test, "double", list(dims) + [5]), test)
        b = parse_like_benchmark(run_driver(test, "double", list(dims) + [5], env={"B200_NGPUS": "2"}), test)
        assert a["f_mean"] == b["f_mean"] and a["i_mean"] == b["i_mean"], (test, a, b)
