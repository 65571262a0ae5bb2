#!/usr/bin/env python
"""tools/matmul_bench.py [n] [reps] -- TFLOP/s of the matmul test (C += A*B, n x n x n) through b200_sweep_loop: the
hand-written tensor-core kernels (double: DMMA mma.sync; float: tcgen05 3xTF32), beside cuBLAS on the same operands
(torch.addmm_, SGEMM with allow_tf32 = False).  Also the command tools/make_profiles.sh captures under ncu (-k regex:matmul)."""
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
    torch.backends.cuda.matmul.allow_tf32 = False
    stream = torch.cuda.current_stream().cuda_stream
    for real, dt in (("double", torch.float64), ("float", torch.float32)):
        A = torch.rand(n * n, device="cuda", dtype=dt) * 2 - 1
        B = torch.rand(n * n, device="cuda", dtype=dt) * 2 - 1
        out = {}
        for mode in ("b200", "cublas"):
            Cm = torch.zeros(n * n, device="cuda", dtype=dt)
            ptrs = [A.data_ptr(), B.data_ptr(), Cm.data_ptr()]
            At, Bt, Ct = A.view(n, n), B.view(n, n), Cm.view(n, n)
            if mode == "b200":
                run = lambda k: pkg.capi.sweep_loop("matmul", real, n, n, n, [], ptrs, k, stream=stream)   # noqa: E731
            else:
                def run(k):
                    for _ in range(k):
                        Ct.addmm_(Bt, At)       # column-major C += A B  ==  row-major C^T += B^T A^T
            run(1)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run(reps)
            e1.record()
            torch.cuda.synchronize()
            out[mode] = 2 * n ** 3 / (e0.elapsed_time(e1) / reps) / 1e9
        print(f"matmul {real} {n}^3: b200 {out['b200']:.1f} TFLOP/s, cuBLAS {out['cublas']:.1f} TFLOP/s, ratio {out['b200'] / out['cublas']:.3f}", flush=True)


if __name__ == "__main__":
    main()
