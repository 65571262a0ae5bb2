#!/usr/bin/env python
"""tools/one_sweep.py <test> <real> <nx> <ny> <ns> [nt] -- run nt sweeps through the context API on random host
arrays and compare with the CPU oracle (for compute-sanitizer / quick debugging on the GPU box)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from pkgload import load_pkg
from oracle_util import Oracle
from parity_util import normwise
pkg = load_pkg()
test, real = sys.argv[1], sys.argv[2]
nx, ny, ns = int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
nt = int(sys.argv[6]) if len(sys.argv) > 6 else 1
o = Oracle("strict" if test == "gameoflife" else "fast")
sc, inputs, _ = o.init(test, real, nx, ny, ns)
want = [a.copy() for a in inputs]
o.run(test, real, nx, ny, ns, nt, sc, want)
got = [a.copy() for a in inputs]
ctx = pkg.Context(1)
slot, stats = ctx.run_on_host_arrays(test, real, nx, ny, ns, sc, got, nt)
print(test, real, nx, ny, ns, "nt", nt, "err", [normwise(g, w) for g, w in zip(got, want)], stats)
ctx.destroy()
