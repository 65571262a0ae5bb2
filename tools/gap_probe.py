#!/usr/bin/env python
"""Isolated vs back-to-back sweep time (is the steady-state loss launch gaps, L2 write-back, clocks?)."""
import sys, statistics
sys.path.insert(0, ".")
import torch
from pkgload import load_pkg
pkg = load_pkg()
from kernelgen_perf_tests_b200 import slab
sc = {"laplacian": [0.68, -0.21], "wave13pt": [0.68, -0.035, 0.094], "divergence": [0.6, -0.2, 0.5]}
for test in sys.argv[1:] or ["wave13pt", "laplacian"]:
    eng = slab.SlabEngine(pkg, test, "double", 512, 256, 256, sc[test])
    eng.run(10); torch.cuda.synchronize()
    def timed(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.run(n); e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / n
    iso = [timed(1) for _ in range(30)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    isof = []
    for _ in range(20):
        flush.zero_(); torch.cuda.synchronize(); isof.append(timed(1))
    b10 = [timed(10) for _ in range(10)]
    b100 = [timed(100) for _ in range(3)]
    b1000 = timed(1000)
    print(f"{test}: isolated {statistics.median(iso):.1f} us, isolated after L2 flush {statistics.median(isof):.1f} us, "
          f"x10 {statistics.median(b10):.1f}, x100 {statistics.median(b100):.1f}, x1000 {b1000:.1f} us/sweep")
    eng.close()
