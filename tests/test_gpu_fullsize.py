"""Parity at BASELINE.json's full single-GPU sizes (1024^3 double = the ">= 1024^3 double grid" of
the north star, and 1024x1024x512 float = configs[2]), where a whole CPU sweep of every test would
take minutes and the host copies tens of GB.

Stencils are local, so the check is block-sampled: before the sweeps, sub-blocks of every input
array (full x extent -- uxx1's dth = 1/nx depends on it -- by a few rows and planes) are copied
out; the CPU oracle runs the same nt sweeps on each block as a small stand-alone grid; afterwards
the corresponding region of the GPU arrays must agree, everywhere at least nt * 2 points away
from the block faces that are not faces of the global grid (the oracle freezes its own shell, the
real grid only the global one).  Blocks sit on every kind of boundary (corners, edges, tile seams)
and in the middle.  Same bars as tests/test_gpu_parity.py.
"""
import numpy as np
import pytest

from parity_util import TOL, normwise, uxx1_bound

pytestmark = pytest.mark.gpu

SCAL = {"laplacian": [0.680375, -0.211234], "wave13pt": [0.680375, -0.211234 / 6, 0.566198 / 6],
        "divergence": [0.680375, -0.211234, 0.566198], "gradient": [0.680375, -0.211234, 0.566198],
        "uxx1": [0.680375, -0.211234], "lapgsrb": [0.680375, -0.211234 / 6, 0.566198 / 12, 0.59688 / 6],
        "jacobi": [0.680375, -0.211234 / 4, 0.566198 / 4],
        "gaussblur": [0.680375, -0.211234, 0.566198, 0.59688, 0.823295, -0.604897]}
LOCAL_TESTS = ["laplacian", "wave13pt", "divergence", "gradient", "uxx1", "lapgsrb", "jacobi", "gaussblur",
               "gameoflife", "tricubic", "tricubic2", "vecadd", "sincos"]
NT = 2
MARGIN = 2 * NT


def _blocks(n_y, n_z, by, bz, rng):
    """(y0, z0) of the sampled blocks: the four corners/edges of the (y,z) face, tile seams, random."""
    ys = [0, n_y - by, 24 - 3, 48 * 5 - 7, n_y // 2 - 5, int(rng.integers(1, n_y - by))]
    zs = [0, n_z - bz, n_z // 2 - 3, int(rng.integers(1, max(2, n_z - bz)))] if n_z > 1 else [0]
    out = [(ys[0], zs[0]), (ys[1], zs[min(1, len(zs) - 1)]), (ys[0], zs[min(1, len(zs) - 1)]), (ys[1], zs[0])]
    for i in range(2, len(ys)):
        out.append((ys[i], zs[i % len(zs)]))
    return [(max(0, min(y, n_y - by)), max(0, min(z, max(0, n_z - bz)))) for y, z in out]


def _valid(lo, b, n):
    """Comparable range inside a block [lo, lo+b) of an axis of extent n."""
    return (0 if lo == 0 else MARGIN), (b if lo + b == n else b - MARGIN)


def run_blocks(pkg, oracle, oracle_strict, test, real, nx, ny, ns):
    import torch
    from kernelgen_perf_tests_b200 import slab
    info = pkg.test_info(test)
    three_d = info["ndims"] == 3
    nz = ns if three_d else 1
    by, bz = (28, 14) if three_d else (72, 1)
    rng = np.random.default_rng(99)
    eng = slab.SlabEngine(pkg, test, real, nx, ny, nz, SCAL.get(test, []), seed=4321)
    try:
        views = [t.view(nz, ny, nx) for t in eng.t]
        blocks = _blocks(ny, nz, by, bz, rng)

        def grab():
            return [[v[z0:z0 + bz, y0:y0 + by, :].contiguous().cpu().numpy().reshape(-1) for v in views]
                    for (y0, z0) in blocks]

        before = grab()
        eng.run(NT)
        torch.cuda.synchronize()
        after = grab()
        o = oracle_strict if test == "gameoflife" else oracle
        worst = 0.0
        for (y0, z0), inp, got in zip(blocks, before, after):
            want = [a.copy() for a in inp]
            slot = o.run(test, real, nx, by, bz, NT, SCAL.get(test, []), want)
            assert slot == eng.result_slot()
            ya, yb = _valid(y0, by, ny)
            za, zb = _valid(z0, bz, nz) if three_d else (0, 1)
            sel = (slice(za, zb), slice(ya, yb), slice(None))
            bound = uxx1_bound(SCAL[test], inp, nx, by, bz, real, NT).reshape(bz, by, nx)[sel] if test == "uxx1" else None
            for q in range(len(got)):
                g = got[q].reshape(bz, by, nx)[sel]
                w = want[q].reshape(bz, by, nx)[sel]
                if test == "gameoflife":
                    assert np.array_equal(g, w), f"gameoflife/{real} block y0={y0}: slot {q} not bit-exact"
                elif test == "uxx1" and q in (0, 1):
                    err = np.abs(g.astype(np.float64) - w.astype(np.float64))
                    assert not (err > bound + 1e-300).any(), f"uxx1/{real} block y0={y0} z0={z0}: beyond the condition-aware bound"
                else:
                    scale = float(np.max(np.abs(want[q]))) or 1.0
                    e = float(np.max(np.abs(g.astype(np.float64) - w.astype(np.float64)))) / scale
                    worst = max(worst, e)
                    assert e <= TOL[real], f"{test}/{real} {nx}x{ny}x{nz} block y0={y0} z0={z0}: slot {q} normwise error {e:.3e}"
        return worst
    finally:
        eng.close()


@pytest.mark.parametrize("test", LOCAL_TESTS)
def test_1024cubed_double_blocks(pkg, oracle, oracle_strict, test):
    info = pkg.test_info(test)
    dims = (1024, 1024, 1024) if info["ndims"] == 3 else (1024, 1024 * 1024, 1)
    run_blocks(pkg, oracle, oracle_strict, test, "double", *dims)


@pytest.mark.parametrize("test", LOCAL_TESTS)
def test_config3_float_blocks(pkg, oracle, oracle_strict, test):
    info = pkg.test_info(test)
    dims = (1024, 1024, 512) if info["ndims"] == 3 else (1024, 1024 * 512, 1)
    run_blocks(pkg, oracle, oracle_strict, test, "float", *dims)
