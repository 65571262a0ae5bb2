"""Pins the oracle against the reference's DECLARATIVE stencil definitions (<test>/<test>.stc, PATUS DSL) -- the one
machine-readable definition of jacobi and sincos the reference holds that needs no Fortran compiler (missing #3:
`/root/reference/jacobi/jacobi.stc:1-13`, `sincos/sincos.stc`).  tests/stc_eval.py parses the .stc (grids, parameters,
domain, offsets, coefficients) and evaluates it with numpy; the oracle's sweep must agree on every array, which checks the
formula, the neighbour offsets AND the interior domain (points outside the .stc domain must stay untouched).
The C tests that have a .stc are cross-checked the same way.  Needs the reference tree (build container only); what
travels to the GPU box are the fixtures generated from the same evaluator (tests/golden/make_golden_stc.py)."""
from pathlib import Path

import numpy as np
import pytest

from stc_eval import BINDINGS, Stencil, stc_sweep

REF = Path("/root/reference")
GOLDEN = Path(__file__).resolve().parent / "golden"
pytestmark = pytest.mark.skipif(not REF.is_dir(), reason="reference tree not present")

# uxx1.stc and tricubic.stc write their domain with nx in all three dimensions ("2 .. nx-2, 2 .. nx-2, 2 .. nx-2"): they are
# evaluated on cubic grids, where that is what the C loops do too; uxx1 is left out (its .stc takes dth as a parameter and
# names the evolving field differently from the C kernel's u0/u1 pair -- covered by the compiled reference instead).
CUBIC = {"tricubic"}


def load(test):
    return Stencil((REF / test / f"{test}.stc").read_text())


def test_jacobi_stc_structure():
    """jacobi.stc:1-13: 9-point support, three coefficients, interior 1 .. n-2 in both dimensions."""
    s = load("jacobi")
    assert s.params == ["c0", "c1", "c2"] and list(s.grids) == ["U"] and s.grids["U"] == 2
    assert s.domain_box({"nx": 40, "ny": 30}) == [(1, 39), (1, 29)]
    offs = {o for g, o, t in s.offsets() if g == "U" and t == 0}
    assert offs == {(dx, dy) for dx in (-1, 0, 1) for dy in (-1, 0, 1)}


def test_sincos_stc_structure():
    s = load("sincos")
    assert s.params == [] and list(s.grids) == ["U", "V", "UV"]
    assert s.domain_box({"nx": 7, "ny": 5, "ns": 3}) == [(0, 7), (0, 5), (0, 3)]      # every point, no shell
    assert {(g, o) for g, o, t in s.offsets()} == {("U", (0, 0, 0)), ("V", (0, 0, 0))}


@pytest.mark.parametrize("real", ["double", "float"])
@pytest.mark.parametrize("test", sorted(BINDINGS))
def test_oracle_matches_stc_definition(oracle_strict, test, real):
    o = oracle_strict
    info = o.info(test)
    s = load(test)
    sizes = [(11, 11, 11), (9, 9, 9)] if test in CUBIC else ([(14, 10, 9), (7, 6, 5)] if info["ndims"] == 3 else [(18, 40, 1), (6, 5, 1)])
    for nx, ny, ns in sizes:
        scalars, inputs, _ = o.init(test, real, nx, ny, ns)
        want = [a.copy() for a in inputs]
        stc_sweep(s, test, nx, ny, ns, scalars, want)
        got = [a.copy() for a in inputs]
        o.sweep(test, real, nx, ny, ns, scalars, got)
        box = s.domain_box({"nx": nx, "ny": ny, "ns": ns})
        assert all(hi > lo for lo, hi in box), "test size must have an interior"
        tol = 2e-13 if real == "double" else 2e-6
        for q in range(len(got)):
            scale = max(float(np.max(np.abs(want[q]))), 1e-300)
            if test == "gameoflife":
                # 1 / (1 + P * 1e20): compare where the .stc value is not a rounding artefact of a cancelling P
                err = float(np.max(np.abs(got[q].astype(np.float64) - want[q].astype(np.float64)) /
                                   np.maximum(np.abs(want[q].astype(np.float64)), 1e-30)))
                assert err <= (1e-9 if real == "double" else 1e-1) or np.allclose(got[q], want[q], rtol=1e-6 if real == "double" else 0.2, atol=1e-30), (test, q, err)
                continue
            err = float(np.max(np.abs(got[q].astype(np.float64) - want[q].astype(np.float64)))) / scale
            assert err <= tol, f"{test}/{real} {nx}x{ny}x{ns}: slot {q}: oracle vs .stc definition differ by {err:.3e}"
        # the shell: what the .stc domain does not cover is bit-for-bit the input in both
        for q in range(len(got)):
            changed = got[q] != inputs[q]
            shape = (ns, ny, nx) if info["ndims"] == 3 else (ny, nx)
            mask = np.zeros(shape, bool)
            sl = tuple(slice(lo, hi) for lo, hi in reversed(box))
            mask[sl] = True
            assert not changed.reshape(shape)[~mask].any(), f"{test}: oracle writes outside the .stc domain (slot {q})"


@pytest.mark.parametrize("real", ["double", "float"])
@pytest.mark.parametrize("test", ["jacobi", "sincos"])
def test_fortran_fixtures_match_oracle(oracle, test, real):
    """The committed fixtures generated from the .stc definitions (these travel to the GPU box) against the oracle run."""
    fx = np.load(GOLDEN / f"{test}_{real}.npz")
    nx, ny, ns, nt = [int(v) for v in fx["dims"]]
    scalars = [float(v) for v in fx["scalars"]]
    n = len([k for k in fx.files if k.startswith("in")])
    work = [fx[f"in{q}"].copy() for q in range(n)]
    oracle.run(test, real, nx, ny, ns, nt, scalars, work)
    for q in range(n):
        ref = fx[f"stc{q}"] if f"stc{q}" in fx.files else fx[f"in{q}"]
        err = float(np.max(np.abs(work[q].astype(np.float64) - ref.astype(np.float64)))) / max(float(np.max(np.abs(ref))), 1e-300)
        assert err <= (1e-12 if real == "double" else 1e-5), (test, real, q, err)
