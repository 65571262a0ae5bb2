"""Generate tests/golden/{jacobi,sincos}_{float,double}.npz from the reference's declarative definitions
(/root/reference/jacobi/jacobi.stc, sincos/sincos.stc) with the interpreter in tests/stc_eval.py.  Run in the build
container:    python tests/golden/make_golden_stc.py

The two tests are Fortran in the reference (no gfortran here, README shows FAIL for their compiled targets), so unlike the
12 C tests (make_golden.py: outputs of the compiled reference) their fixtures hold the outputs of the .stc definition:
inputs drawn in the reference driver's rand() order (jacobi/main.c:90-113, sincos/main.c:95-106, restated by the oracle's
kgo_init), `nt` sweeps with the driver's buffer swap, evaluated in the precision of the arrays."""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
from oracle_util import Oracle          # noqa: E402
from stc_eval import Stencil, stc_sweep  # noqa: E402

REF = Path("/root/reference")
CASES = {"jacobi": (18, 40, 1, 3), "sincos": (14, 10, 9, 2)}


def main():
    o = Oracle("fast")
    for test, (nx, ny, ns, nt) in CASES.items():
        s = Stencil((REF / test / f"{test}.stc").read_text())
        rot = o.info(test)["rotation"]
        for real in ("float", "double"):
            scalars, arrays, i_mean = o.init(test, real, nx, ny, ns)
            out = {"scalars": np.array(scalars), "dims": np.array([nx, ny, ns, nt]), "i_mean": np.array(i_mean)}
            for q, a in enumerate(arrays):
                out[f"in{q}"] = a.copy()
            cur = [a.copy() for a in arrays]
            order = list(range(len(cur)))
            for _ in range(nt):
                stc_sweep(s, test, nx, ny, ns, scalars, [cur[i] for i in order])
                if rot == 2:
                    order[0], order[1] = order[1], order[0]
            for q, a in enumerate(cur):
                if not np.array_equal(a, arrays[q]):
                    out[f"stc{q}"] = a
            np.savez_compressed(HERE / f"{test}_{real}.npz", **out)
            print(test, real, [k for k in out if k.startswith("stc")])


if __name__ == "__main__":
    main()
