"""How long does b200_load_shell take for an output buffer (wave13pt 512x256x256 double, slot 2), next to a whole-array
b200_load and to the individual copies it is made of?  Wall clock around synchronous calls, 10 repetitions each."""
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
torch.cuda.init()
torch.zeros(1, device="cuda")

nx, ny, ns = 512, 256, 256
n = nx * ny * ns
ctx = capi.Context(1)
ctx.plan("wave13pt", "double", nx, ny, ns, [0.1, 0.2, 0.3])
ctx.alloc()
hb = capi.PinnedBuffer(n, np.float64)
hb.array[:] = 1.0


def wall(fn, reps=10):
    fn(); fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


print(f"b200_load whole array (268 MB): {wall(lambda: ctx.load_array(2, hb.array)):.3f} ms")
print(f"b200_load_shell slot 2:         {wall(lambda: ctx.load_array_shell(2, hb.array)):.3f} ms")

cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
rt = C.CDLL(cands[0])
rt.cudaMemcpy2DAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
d = torch.empty(n, dtype=torch.float64, device="cuda")
dp, hp = d.data_ptr(), hb.array.ctypes.data
row_b, plane_b = nx * 8, nx * ny * 8


def c2d(doff, pitch, width, height):
    rc = rt.cudaMemcpy2DAsync(dp + doff, pitch, hp + doff, pitch, width, height, 1, None)
    assert rc == 0, rc


print(f"x edges: 2D copy 32 B x {ny * ns - 1} rows, pitch {row_b}: {wall(lambda: c2d(row_b - 16, row_b, 32, ny * ns - 1)):.3f} ms")
print(f"y shell: 2D copy {2 * row_b} B x {ns} rows, pitch {plane_b}: {wall(lambda: c2d(0, plane_b, 2 * row_b, ns)):.3f} ms")
print(f"z shell: 2 planes contiguous ({2 * plane_b} B): {wall(lambda: rt.cudaMemcpyAsync(dp, hp, 2 * plane_b, 1, None)):.3f} ms")
print(f"contiguous 2 MB: {wall(lambda: rt.cudaMemcpyAsync(dp, hp, 2 << 20, 1, None)):.3f} ms")
# alternative for the x edges: gather on the host into a pinned staging buffer, one contiguous copy
stage = capi.PinnedBuffer((ny * ns) * 4, np.float64)
v = hb.array.reshape(ny * ns, nx)


def gather():
    s = stage.array.reshape(ny * ns, 4)
    s[:, :2] = v[:, :2]
    s[:, 2:] = v[:, -2:]


t0 = time.perf_counter()
for _ in range(10):
    gather()
print(f"host gather of the x edges with numpy (2 MB): {(time.perf_counter() - t0) / 10 * 1e3:.3f} ms")
