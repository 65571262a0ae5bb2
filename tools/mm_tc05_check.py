#!/usr/bin/env python
"""tools/mm_tc05_check.py -- bring-up / accuracy / speed check of the tcgen05 float matmul (k_matmul_tc05.cu) through
b200_sweep_loop: normwise error against a float64 product for a ladder of shapes, then TFLOP/s at 4096 and 8192."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from pkgload import load_pkg
pkg = load_pkg()
pkg.load()
stream = torch.cuda.current_stream().cuda_stream
torch.manual_seed(1)


def run(m, k, n, sweeps=1, c_init=False):
    A = torch.rand(m * k, device="cuda") * 2 - 1       # column-major m x k
    B = torch.rand(k * n, device="cuda") * 2 - 1       # column-major k x n
    C0 = (torch.rand(m * n, device="cuda") * 2 - 1) if c_init else torch.zeros(m * n, device="cuda")
    C = C0.clone()
    pkg.capi.sweep_loop("matmul", "float", m, k, n, [], [A.data_ptr(), B.data_ptr(), C.data_ptr()], sweeps, stream=stream)
    torch.cuda.synchronize()
    ref = C0.view(n, m).t().double() + sweeps * (A.view(k, m).t().double() @ B.view(n, k).t().double())
    got = C.view(n, m).t().double()
    err = float((got - ref).abs().max() / ref.abs().max())
    return err


for shape in [(128, 32, 128), (128, 64, 128), (128, 256, 128), (256, 128, 256), (384, 96, 640), (132, 36, 140), (1024, 1024, 1024),
              (2048, 8192, 256), (300, 512, 140)]:
    try:
        e = run(*shape)
        e2 = run(*shape, sweeps=2, c_init=True)
        print(f"{shape}: normwise err {e:.3e}   (2 sweeps, C0 != 0: {e2:.3e})", flush=True)
    except Exception as ex:
        print(shape, "FAILED", str(ex)[:200], flush=True)
        break

for n in (4096, 8192):
    A = torch.rand(n * n, device="cuda") * 2 - 1
    B = torch.rand(n * n, device="cuda") * 2 - 1
    C = torch.zeros(n * n, device="cuda")
    ptrs = [A.data_ptr(), B.data_ptr(), C.data_ptr()]
    pkg.capi.sweep_loop("matmul", "float", n, n, n, [], ptrs, 1, stream=stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pkg.capi.sweep_loop("matmul", "float", n, n, n, [], ptrs, 3, stream=stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    torch.backends.cuda.matmul.allow_tf32 = False
    At, Bt, Ct = A.view(n, n), B.view(n, n), torch.zeros(n, n, device="cuda")
    Ct.addmm_(Bt, At)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        Ct.addmm_(Bt, At)
    e1.record()
    torch.cuda.synchronize()
    ms_c = e0.elapsed_time(e1) / 3
    print(f"n={n}: b200 {2 * n ** 3 / ms / 1e9:.1f} TFLOP/s ({ms:.3f} ms)   cuBLAS SGEMM {2 * n ** 3 / ms_c / 1e9:.1f} TFLOP/s ({ms_c:.3f} ms)", flush=True)
