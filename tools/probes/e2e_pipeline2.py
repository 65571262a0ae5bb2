"""Follow-up of e2e_pipeline.py: which ingredient of the C ABI job stops the return copy from overlapping the next upload?
Two streams, alternating; per step: H2D 2 x 268 MB (+ optional extra copies), D2H 268 MB; variants toggle one ingredient."""
import ctypes as C
import glob
import importlib
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
pkg = importlib.import_module("kernelgen-perf-tests_b200")
capi = pkg.capi
torch.zeros(1, device="cuda")
nx, ny, ns = 512, 256, 256
n = nx * ny * ns
row_b, plane_b, arr_b = nx * 8, nx * ny * 8, n * 8
rt = C.CDLL(glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))[0])
rt.cudaMemcpy2DAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
H2D, D2H = 1, 2


def make_lane(portable):
    L = dict(s=torch.cuda.Stream(), d=torch.empty(3 * n, dtype=torch.float64, device="cuda"))
    if portable:
        L["pb"] = [capi.PinnedBuffer(n, np.float64) for _ in range(3)]
        L["h"] = [p.array.ctypes.data for p in L["pb"]]
    else:
        L["t"] = [torch.empty(n, dtype=torch.float64, pin_memory=True) for _ in range(3)]
        L["h"] = [t.data_ptr() for t in L["t"]]
    return L


def bench(name, portable=False, xedge=False, yshell=False, small=0, same_buf=False, steps=12):
    lanes = [make_lane(portable) for _ in range(2)]

    def step(i):
        L = lanes[i % 2]
        L["s"].synchronize()
        st, d, h = L["s"].cuda_stream, L["d"].data_ptr(), L["h"]
        for q in range(2):
            assert rt.cudaMemcpyAsync(d + q * arr_b, h[q], arr_b, H2D, st) == 0
        if xedge:
            assert rt.cudaMemcpy2DAsync(d + 2 * arr_b + row_b - 16, row_b, h[2] + row_b - 16, row_b, 32, ny * ns - 1, H2D, st) == 0
        if yshell:
            assert rt.cudaMemcpy2DAsync(d + 2 * arr_b, plane_b, h[2], plane_b, 2 * row_b, ns, H2D, st) == 0
        for k in range(small):
            assert rt.cudaMemcpyAsync(d + 2 * arr_b + k * 4096, h[2] + k * 4096, 16, H2D, st) == 0
        assert rt.cudaMemcpyAsync(h[2] if same_buf else h[0], d + 2 * arr_b, arr_b, D2H, st) == 0

    for i in range(2):
        step(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        step(i)
    torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter() - t0) / steps * 1e3:.2f} ms/step", flush=True)
    for L in lanes:
        for p in L.get("pb", []):
            p.free()


bench("baseline: 2 x 268 MB up, 268 MB down (torch pinned memory)")
bench("+ b200_host_alloc (cudaHostAllocPortable) buffers", portable=True)
bench("+ x-edge 2D copy, 32 B x 65535 rows", xedge=True)
bench("+ y-shell 2D copy, 8 KB x 256 rows", yshell=True)
bench("+ 4 copies of 16 B", small=4)
bench("+ the return copy lands in the buffer the shell came from", xedge=True, yshell=True, same_buf=True)
bench("all of it, portable buffers", portable=True, xedge=True, yshell=True, small=4, same_buf=True)
