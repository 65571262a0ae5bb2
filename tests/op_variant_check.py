"""Helper (not a test module): run the given 3D tests through the C ABI on host arrays with whatever kernel forms the
environment of THIS process selects (B200_TILE_POLICY, ...; the library reads such switches once), compare with the
CPU oracle, print one JSON line {"worst": {"float": e, "double": e}, "sha": digest of the outputs}; exit 1 beyond the
tolerances of tests/parity_util.py.  usage: op_variant_check.py <test> [<test> ...]"""
import hashlib
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from pkgload import load_pkg          # noqa: E402
from oracle_util import Oracle        # noqa: E402
from parity_util import TOL, normwise  # noqa: E402

# several tiles in x and y for 12-, 24- and 48-row tiles, partial last tiles, odd pitches (scalar loader), tiny grids
SIZES = [(260, 101, 21), (128, 20, 12), (130, 37, 29), (63, 31, 29), (256, 70, 9), (16, 5, 5), (5, 5, 5)]


def main(tests):
    pkg = load_pkg()
    o = Oracle("fast")
    ctx = pkg.Context(1)
    sha = hashlib.sha256()
    worst = {"float": 0.0, "double": 0.0}
    try:
        for test in tests:
            for real in ("float", "double"):
                for nx, ny, ns in SIZES:
                    for nt in (1, 3):
                        scalars, inputs, _ = o.init(test, real, nx, ny, ns)
                        want = [a.copy() for a in inputs]
                        slot_o = o.run(test, real, nx, ny, ns, nt, scalars, want)
                        got = [a.copy() for a in inputs]
                        slot, _ = ctx.run_on_host_arrays(test, real, nx, ny, ns, scalars, got, nt)
                        assert slot == slot_o
                        for g, w in zip(got, want):
                            worst[real] = max(worst[real], float(normwise(g, w)))
                            sha.update(g.tobytes())
    finally:
        ctx.destroy()
    print(json.dumps({"tests": tests, "worst": worst, "sha": sha.hexdigest()}))
    return 0 if all(worst[r] <= TOL[r] for r in worst) else 1


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
