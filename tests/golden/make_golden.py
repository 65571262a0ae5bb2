"""Generate tests/golden/*.npz from the REFERENCE itself (oracle/_ref, i.e. the unmodified
sources under /root/reference compiled by oracle/build_ref.sh).  Run in the build container:

    python tests/golden/make_golden.py

Each fixture holds, for one (test, real) case: the rand()-initialised inputs, the scalars, and
the arrays after `nt` sweeps of the reference kernel driven with the reference driver's pointer
rotation (e.g. laplacian/laplacian.c:287-313), for both the shipped-flags build and the
strict-IEEE build, plus the i_mean / f_mean text the reference gcc-target binary prints.
The fixtures travel to the GPU box; /root/reference does not.
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
from oracle_util import Oracle, RefKernels, REF_C_TESTS, run_ref_binary  # noqa: E402

CASES3D = (14, 10, 9, 3)      # nx ny ns nt  (nx*8 % 16 == 0 -> TMA path on the GPU)
CASES2D = (18, 40, 1, 3)


def rotate_run(sweep, rot, nt, arrays):
    cur, idxs = list(arrays), [0, 1, 2]
    for _ in range(nt):
        sweep(cur)
        if rot == 2:
            cur[0], cur[1] = cur[1], cur[0]
            idxs[0], idxs[1] = idxs[1], idxs[0]
        elif rot == 3:
            cur = [cur[1], cur[2], cur[0]]
            idxs = [idxs[1], idxs[2], idxs[0]]
    return idxs


def main():
    o = Oracle("fast")
    refs = {"shipped": RefKernels("shipped"), "strict": RefKernels("strict")}
    for test in REF_C_TESTS:
        info = o.info(test)
        nx, ny, ns, nt = CASES3D if info["ndims"] == 3 else CASES2D
        for real in ("float", "double"):
            scalars, arrays, i_mean = o.init(test, real, nx, ny, ns)
            out = {"scalars": np.array(scalars), "dims": np.array([nx, ny, ns, nt])}
            for q, a in enumerate(arrays):
                out[f"in{q}"] = a.copy()
            for flav, ref in refs.items():
                work = [a.copy() for a in arrays]
                rotate_run(lambda cur: ref.sweep(test, real, nx, ny, ns, scalars, cur), info["rotation"], nt, work)
                for q, a in enumerate(work):
                    if not np.array_equal(a, arrays[q]):      # only arrays the sweeps wrote
                        out[f"{flav}{q}"] = a
            args = [nx, ny, ns, nt] if info["ndims"] == 3 else [nx, ny, nt]
            txt = run_ref_binary(test, real, args)
            out["i_mean"] = np.array(txt["i_mean"])
            out["f_mean"] = np.array(txt["f_mean"])
            np.savez_compressed(HERE / f"{test}_{real}.npz", **out)
            print(test, real, txt["i_mean"], txt["f_mean"])


if __name__ == "__main__":
    main()
