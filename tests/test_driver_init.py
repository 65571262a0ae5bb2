"""CPU tests of the drivers' input initialisation (kernelgen-perf-tests_b200/drivers/kg_rand.[ch], kg_init.h):
the parallel fill (B200_INIT_THREADS > 1) must reproduce glibc's unseeded rand() stream -- which DEFINES the
reference's inputs (e.g. laplacian/laplacian.c:112,158-165) -- bit for bit, for any thread count, any number of
interleaved arrays and any number of coefficient draws before the arrays."""
import subprocess
from pathlib import Path

import pytest

DRV = Path(__file__).resolve().parent.parent / "kernelgen-perf-tests_b200" / "drivers"
CC = ["gcc", "-O3", "-ffast-math", "-march=x86-64-v3", "-D_GNU_SOURCE", "-std=c99", "-Wall"]


@pytest.fixture(scope="module")
def bins(tmp_path_factory):
    d = tmp_path_factory.mktemp("kg")
    out = {}
    subprocess.run(CC + [str(DRV / "kg_rand_check.c"), str(DRV / "kg_rand.c"), f"-I{DRV}", "-o", str(d / "rc")], check=True)
    out["rand"] = d / "rc"
    for real in ("float", "double"):
        subprocess.run(CC + [f"-Dreal={real}", str(DRV / "init_check.c"), str(DRV / "kg_rand.c"), f"-I{DRV}",
                             "-o", str(d / f"ic_{real}"), "-lpthread"], check=True)
        out[real] = d / f"ic_{real}"
    return out


def test_kg_rand_is_glibc_rand(bins):
    p = subprocess.run([str(bins["rand"])], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and "kg_rand OK" in p.stdout, p.stdout


# (arrays, elements, coefficient draws before, threads): laplacian (2 arrays, 2 coefficients), wave13pt (3, 3),
# uxx1 (6, 2), gaussblur (2, 6), tricubic (5, 0), matvec/matmul groups (1 array), ragged splits, more threads than work
CASES = [(2, 1000003, 2, 7), (3, 262144, 3, 16), (6, 500001, 2, 5), (2, 99991, 6, 3), (5, 40000, 0, 64),
         (1, 77, 0, 8), (1, 8191, 123456, 2), (4, 4097, 1, 256)]


@pytest.mark.parametrize("real", ["float", "double"])
@pytest.mark.parametrize("na,n,before,threads", CASES)
def test_parallel_fill_bit_identical(bins, real, na, n, before, threads):
    p = subprocess.run([str(bins[real]), str(na), str(n), str(before), str(threads)], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and p.stdout.startswith("identical"), p.stdout
    s_serial, s_par = [float(v) for v in p.stdout.split()[1:3]]
    # the sums (-> "initial mean") agree up to re-association of a `real` sum
    tol = (1e-3 if real == "float" else 1e-9) * max(1.0, na * n) ** 0.5
    assert abs(s_serial - s_par) <= tol
