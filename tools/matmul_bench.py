#!/usr/bin/env python
"""tools/matmul_bench.py [n] -- TFLOP/s of the matmul test (C += A*B, n x n x n) through b200_sweep_loop:
hand-written tensor-core kernels vs the cuBLAS baseline (B200_MATMUL=cublas), float and double.
Needs the diagnostics library: make -C kernelgen-perf-tests_b200/csrc diag; B200_LIB=kernelgen-perf-tests_b200/libb200stencil_diag.so"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    import torch
    from pkgload import load_pkg
    pkg = load_pkg()
    pkg.load()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    for real, dt in (("double", torch.float64), ("float", torch.float32)):
        A = torch.rand(n * n, device="cuda", dtype=dt) * 2 - 1
        B = torch.rand(n * n, device="cuda", dtype=dt) * 2 - 1
        for mode in ("tensor", "cublas"):
            if mode == "cublas":
                os.environ["B200_MATMUL"] = "cublas"
            else:
                os.environ.pop("B200_MATMUL", None)
            Cm = torch.zeros(n * n, device="cuda", dtype=dt)
            ptrs = [A.data_ptr(), B.data_ptr(), Cm.data_ptr()]
            stream = torch.cuda.current_stream().cuda_stream
            pkg.capi.sweep_loop("matmul", real, n, n, n, [], ptrs, 1, stream=stream)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            pkg.capi.sweep_loop("matmul", real, n, n, n, [], ptrs, reps, stream=stream)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            ref = (A.view(n, n).T[:64].to(torch.float64) @ B.view(n, n).T.to(torch.float64)) * (reps + 1)
            err = float((Cm.view(n, n).T[:64].to(torch.float64) - ref).abs().max() / ref.abs().max())
            print(f"matmul {n}^3 {real:6s} {mode:6s}: {ms:8.2f} ms  {2 * n ** 3 / ms / 1e9:7.1f} TFLOP/s  normwise err {err:.2e}",
                  flush=True)
    os.environ.pop("B200_MATMUL", None)


if __name__ == "__main__":
    main()
