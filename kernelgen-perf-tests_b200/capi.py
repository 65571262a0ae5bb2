"""ctypes binding of libb200stencil.so -- one Python function per C-ABI entry point of
include/b200_stencil.h.  No compute happens here and there is no fallback: if the shared
library is missing, `load()` raises (build it with `python -c "import __graft_entry__ as g; g.build()"`
or `make -C kernelgen-perf-tests_b200/csrc`).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
# B200_LIB selects an experiment build of the same library (tools/ A/B measurements); default: the product
LIB_PATH = Path(os.environ["B200_LIB"]) if os.environ.get("B200_LIB") else PKG_DIR / "libb200stencil.so"

TESTS = ["laplacian", "wave13pt", "divergence", "gradient", "uxx1", "lapgsrb", "jacobi",
         "gaussblur", "gameoflife", "tricubic", "tricubic2", "vecadd", "matvec", "sincos", "matmul"]
TEST_ID = {n: i for i, n in enumerate(TESTS)}
F32, F64 = 0, 1
_DT = {"float": F32, "double": F64, "f32": F32, "f64": F64, np.float32: F32, np.float64: F64,
       np.dtype("float32"): F32, np.dtype("float64"): F64, F32: F32, F64: F64}
NP_DTYPE = {F32: np.float32, F64: np.float64}

MAX_ARRAYS, MAX_SCALARS = 8, 8

# every symbol include/b200_stencil.h declares (checked by tests/test_abi.py)
EXPORTS = ["b200_get_test_info", "b200_test_by_name", "b200_last_error", "b200_api_version",
           "b200_interior_points", "b200_device_count", "b200_sweep", "b200_sweep_loop", "b200_sweep2_supported", "b200_sweep2_profitable", "b200_sweep2", "b200_sweep_loop2", "b200_slab_loop", "b200_kernel_info",
           "b200_launch_count", "b200_init", "b200_plan", "b200_alloc", "b200_load", "b200_run",
           "b200_slot_interior_dead", "b200_load_shell", "b200_result_slot", "b200_save", "b200_rewind", "b200_set_async", "b200_sync", "b200_free", "b200_destroy", "b200_host_alloc",
           "b200_host_free", "b200_device_alloc", "b200_device_free", "b200_ipc_export",
           "b200_ipc_import", "b200_ipc_close", "b200_signal", "b200_wait", "b200_load_shell_slab"]
IPC_HANDLE_BYTES = 64


class B200Error(RuntimeError):
    pass


class TestInfo(C.Structure):
    _fields_ = [("name", C.c_char_p), ("ndims", C.c_int), ("narrays", C.c_int), ("nscalars", C.c_int),
                ("rotation", C.c_int), ("lo", C.c_int * 3), ("hi", C.c_int * 3), ("nread", C.c_int),
                ("nwritten", C.c_int), ("zghost_lo", C.c_int), ("zghost_hi", C.c_int),
                ("exchange_slot", C.c_int)]


class SweepDesc(C.Structure):
    _fields_ = [("test", C.c_int), ("dtype", C.c_int), ("nx", C.c_int), ("ny", C.c_int), ("ns", C.c_int),
                ("scalars", C.c_double * MAX_SCALARS), ("out_begin", C.c_int), ("out_end", C.c_int),
                ("push_lo", C.c_void_p), ("push_lo_src_plane", C.c_int), ("push_lo_dst_plane", C.c_int),
                ("push_lo_count", C.c_int),
                ("push_hi", C.c_void_p), ("push_hi_src_plane", C.c_int), ("push_hi_dst_plane", C.c_int),
                ("push_hi_count", C.c_int), ("reverse_order", C.c_int),
                ("wait_flag", C.c_void_p * 2), ("wait_value", C.c_ulonglong),
                ("signal_flag", C.c_void_p * 2), ("signal_value", C.c_ulonglong)]


class Stats(C.Structure):
    _fields_ = [("kernel_ms_per_sweep", C.c_double), ("kernel_ms_total", C.c_double),
                ("regs_per_thread", C.c_int), ("launches", C.c_int), ("kernel_name", C.c_char_p),
                ("ngpus", C.c_int)]


_lib = None


def load() -> C.CDLL:
    """Load libb200stencil.so (once).  Fails loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise B200Error(f"{LIB_PATH} not built: the b200 target has no fallback path "
                        f"(run `make -C {PKG_DIR / 'csrc'}`)")
    L = C.CDLL(str(LIB_PATH), mode=os.RTLD_GLOBAL if hasattr(os, "RTLD_GLOBAL") else 0)
    L.b200_get_test_info.restype = C.POINTER(TestInfo)
    L.b200_get_test_info.argtypes = [C.c_int]
    L.b200_test_by_name.restype = C.c_int
    L.b200_test_by_name.argtypes = [C.c_char_p]
    L.b200_last_error.restype = C.c_char_p
    L.b200_api_version.restype = C.c_int
    L.b200_interior_points.restype = C.c_ulonglong
    L.b200_interior_points.argtypes = [C.c_int] * 4
    L.b200_device_count.argtypes = [C.POINTER(C.c_int)]
    L.b200_sweep.argtypes = [C.POINTER(SweepDesc), C.POINTER(C.c_void_p), C.c_void_p]
    L.b200_sweep_loop.argtypes = [C.POINTER(SweepDesc), C.POINTER(C.c_void_p), C.c_int, C.c_void_p]
    L.b200_sweep2_supported.argtypes = [C.c_int]
    L.b200_sweep2_profitable.argtypes = [C.c_int, C.c_int]
    L.b200_sweep2.argtypes = [C.POINTER(SweepDesc), C.POINTER(C.c_void_p), C.c_void_p]
    L.b200_sweep_loop2.argtypes = [C.POINTER(SweepDesc), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, C.c_void_p]
    L.b200_slab_loop.argtypes = [C.POINTER(SweepDesc), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                 C.c_int, C.c_ulonglong, C.c_void_p]
    L.b200_kernel_info.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_char_p)]
    L.b200_launch_count.restype = C.c_ulonglong
    L.b200_init.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    L.b200_plan.argtypes = [C.c_void_p] + [C.c_int] * 5 + [C.POINTER(C.c_double), C.c_int]
    L.b200_alloc.argtypes = [C.c_void_p]
    L.b200_load.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.b200_load_shell.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.b200_slot_interior_dead.argtypes = [C.c_int, C.c_int]
    L.b200_run.argtypes = [C.c_void_p, C.c_int, C.POINTER(Stats)]
    L.b200_result_slot.argtypes = [C.c_void_p]
    L.b200_save.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.b200_rewind.argtypes = [C.c_void_p]
    L.b200_set_async.argtypes = [C.c_void_p, C.c_int]
    L.b200_sync.argtypes = [C.c_void_p]
    L.b200_free.argtypes = [C.c_void_p]
    L.b200_destroy.argtypes = [C.c_void_p]
    L.b200_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    L.b200_host_free.argtypes = [C.c_void_p]
    L.b200_device_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    L.b200_device_free.argtypes = [C.c_void_p]
    L.b200_ipc_export.argtypes = [C.c_void_p, C.c_char_p]
    L.b200_ipc_import.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    L.b200_ipc_close.argtypes = [C.c_void_p]
    L.b200_signal.argtypes = [C.c_void_p, C.c_ulonglong, C.c_void_p]
    L.b200_wait.argtypes = [C.c_void_p, C.c_ulonglong, C.c_void_p]
    L.b200_load_shell_slab.argtypes = [C.c_int] * 7 + [C.c_void_p, C.c_void_p, C.c_void_p]
    _lib = L
    return L


def _check(rc: int):
    if rc != 0:
        raise B200Error(f"b200 error {rc}: {load().b200_last_error().decode()}")


def _tid(test) -> int:
    return test if isinstance(test, int) else TEST_ID[test]


def test_info(test) -> dict:
    i = load().b200_get_test_info(_tid(test)).contents
    return dict(name=i.name.decode(), ndims=i.ndims, narrays=i.narrays, nscalars=i.nscalars,
                rotation=i.rotation, lo=list(i.lo), hi=list(i.hi), nread=i.nread, nwritten=i.nwritten,
                zghost_lo=i.zghost_lo, zghost_hi=i.zghost_hi, exchange_slot=i.exchange_slot)


def interior_points(test, nx, ny, ns) -> int:
    return int(load().b200_interior_points(_tid(test), nx, ny, ns))


def device_count() -> int:
    n = C.c_int(0)
    _check(load().b200_device_count(C.byref(n)))
    return n.value


def launch_count() -> int:
    return int(load().b200_launch_count())


def kernel_info(test, dtype) -> dict:
    regs = C.c_int(0)
    name = C.c_char_p()
    _check(load().b200_kernel_info(_tid(test), _DT[dtype], C.byref(regs), C.byref(name)))
    return dict(regs=regs.value, name=name.value.decode())


def sweep(test, dtype, nx, ny, ns, scalars, device_ptrs, stream=0, out_range=None, push=None):
    """One sweep on DEVICE buffers (raw addresses, slot order).  Asynchronous on `stream`
    (a cudaStream_t address, 0 = default).  push = dict(lo=(ptr, src, dst, count), hi=(...))."""
    d = SweepDesc()
    d.test, d.dtype, d.nx, d.ny, d.ns = _tid(test), _DT[dtype], nx, ny, ns
    for i, v in enumerate(scalars):
        d.scalars[i] = float(v)
    if out_range is not None:
        d.out_begin, d.out_end = out_range
    if push:
        if push.get("lo"):
            d.push_lo, d.push_lo_src_plane, d.push_lo_dst_plane, d.push_lo_count = push["lo"]
        if push.get("hi"):
            d.push_hi, d.push_hi_src_plane, d.push_hi_dst_plane, d.push_hi_count = push["hi"]
    ptrs = (C.c_void_p * len(device_ptrs))(*device_ptrs)
    _check(load().b200_sweep(C.byref(d), ptrs, C.c_void_p(stream)))


def sweep_loop(test, dtype, nx, ny, ns, scalars, device_ptrs, niters, stream=0, out_range=None):
    """`niters` sweeps back to back with the reference rotation applied in C; returns the rotated
    pointer list (the order the next sweep would see)."""
    d = SweepDesc()
    d.test, d.dtype, d.nx, d.ny, d.ns = _tid(test), _DT[dtype], nx, ny, ns
    for i, v in enumerate(scalars):
        d.scalars[i] = float(v)
    if out_range is not None:
        d.out_begin, d.out_end = out_range
    ptrs = (C.c_void_p * len(device_ptrs))(*device_ptrs)
    _check(load().b200_sweep_loop(C.byref(d), ptrs, niters, C.c_void_p(stream)))
    return [int(p) if p else 0 for p in ptrs]


def sweep2_supported(test) -> bool:
    return bool(load().b200_sweep2_supported(_tid(test)))


def sweep2_profitable(test, nx) -> bool:
    return bool(load().b200_sweep2_profitable(_tid(test), nx))


def sweep_loop2(test, dtype, nx, ny, ns, scalars, device_ptrs, scratch_ptr, niters, stream=0):
    """b200_sweep_loop2: fused two-sweep passes (temporal blocking) + the remaining single sweeps.  Returns
    (rotated pointer list, scratch pointer) -- after a fused pass the old w0 buffer is the scratch."""
    d = SweepDesc()
    d.test, d.dtype, d.nx, d.ny, d.ns = _tid(test), _DT[dtype], nx, ny, ns
    for i, v in enumerate(scalars):
        d.scalars[i] = float(v)
    ptrs = (C.c_void_p * len(device_ptrs))(*device_ptrs)
    scr = C.c_void_p(scratch_ptr)
    _check(load().b200_sweep_loop2(C.byref(d), ptrs, C.byref(scr), niters, C.c_void_p(stream)))
    return [int(p) if p else 0 for p in ptrs], int(scr.value or 0)


def slab_loop(test, dtype, nx, ny, ns, scalars, device_ptrs, peer_lo, peer_hi, niters, first_sweep,
              out_range, push_lo, push_hi, wait_flags, signal_flags, stream=0):
    """`niters` sweeps of one slab with fused halo push and in-kernel neighbour ordering (b200_slab_loop).
    peer_lo / peer_hi: the neighbours' buffers in rotation order (0 = none); push_lo / push_hi =
    (src_plane, dst_plane, count); wait_flags / signal_flags: two addresses each (0 = none)."""
    d = SweepDesc()
    d.test, d.dtype, d.nx, d.ny, d.ns = _tid(test), _DT[dtype], nx, ny, ns
    for i, v in enumerate(scalars):
        d.scalars[i] = float(v)
    d.out_begin, d.out_end = out_range
    d.push_lo_src_plane, d.push_lo_dst_plane, d.push_lo_count = push_lo
    d.push_hi_src_plane, d.push_hi_dst_plane, d.push_hi_count = push_hi
    for i in range(2):
        d.wait_flag[i] = wait_flags[i] or None
        d.signal_flag[i] = signal_flags[i] or None
    n = len(device_ptrs)
    arr = (C.c_void_p * n)(*device_ptrs)
    plo = (C.c_void_p * n)(*[p or None for p in peer_lo])
    phi = (C.c_void_p * n)(*[p or None for p in peer_hi])
    _check(load().b200_slab_loop(C.byref(d), arr, plo, phi, niters, first_sweep, C.c_void_p(stream)))


def device_alloc(nbytes: int) -> int:
    p = C.c_void_p()
    _check(load().b200_device_alloc(C.byref(p), nbytes))
    return p.value


def device_free(ptr: int):
    _check(load().b200_device_free(C.c_void_p(ptr)))


def ipc_export(ptr: int) -> bytes:
    buf = C.create_string_buffer(IPC_HANDLE_BYTES)
    _check(load().b200_ipc_export(C.c_void_p(ptr), buf))
    return buf.raw


def ipc_import(handle: bytes) -> int:
    p = C.c_void_p()
    _check(load().b200_ipc_import(handle, C.byref(p)))
    return p.value


def ipc_close(ptr: int):
    _check(load().b200_ipc_close(C.c_void_p(ptr)))


def signal_flag(flag_ptr: int, value: int, stream: int = 0):
    _check(load().b200_signal(C.c_void_p(flag_ptr), value, C.c_void_p(stream)))


def wait_flag(flag_ptr: int, value: int, stream: int = 0):
    _check(load().b200_wait(C.c_void_p(flag_ptr), value, C.c_void_p(stream)))


def slot_interior_dead(test, slot: int) -> bool:
    """b200_slot_interior_dead: the first sweep overwrites this array's interior before anything reads it."""
    return bool(load().b200_slot_interior_dead(_tid(test), slot))


def load_shell_slab(test, dtype, nx, ny, split_n, mem_lo, mem_hi, dev_ptr: int, host_ptr: int, stream: int = 0):
    """b200_load_shell_slab: boundary shell of a caller-owned slab, host -> device, asynchronous on `stream`."""
    _check(load().b200_load_shell_slab(_tid(test), _DT[dtype], nx, ny, split_n, mem_lo, mem_hi, C.c_void_p(dev_ptr),
                                       C.c_void_p(host_ptr), C.c_void_p(stream)))


class PinnedBuffer:
    """Page-locked host memory from b200_host_alloc, viewed as a numpy array."""

    def __init__(self, nelem: int, dtype):
        self.dtype = np.dtype(dtype)
        self.ptr = C.c_void_p()
        _check(load().b200_host_alloc(C.byref(self.ptr), nelem * self.dtype.itemsize))
        buf = (C.c_char * (nelem * self.dtype.itemsize)).from_address(self.ptr.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=nelem)

    def free(self):
        if self.ptr:
            self.array = None
            load().b200_host_free(self.ptr)
            self.ptr = None


class Context:
    """The driver-phase API (b200_init / plan / alloc / load / run / save / free) on HOST buffers --
    exactly what the C drivers call."""

    def __init__(self, ngpus: int = 1):
        self.h = C.c_void_p()
        _check(load().b200_init(C.byref(self.h), ngpus))
        self.test = None

    def plan(self, test, dtype, nx, ny, ns, scalars):
        self.test, self.dtype = _tid(test), _DT[dtype]
        self.nx, self.ny, self.ns = nx, ny, ns
        sc = (C.c_double * MAX_SCALARS)(*[float(v) for v in scalars])
        _check(load().b200_plan(self.h, self.test, self.dtype, nx, ny, ns, sc, len(scalars)))

    def alloc(self):
        _check(load().b200_alloc(self.h))

    def load_array(self, slot: int, host: np.ndarray):
        assert host.flags.c_contiguous and host.dtype == NP_DTYPE[self.dtype]
        _check(load().b200_load(self.h, slot, host.ctypes.data_as(C.c_void_p)))

    def load_array_shell(self, slot: int, host: np.ndarray):
        """Boundary shell only (b200_load_shell): for the arrays the first sweep overwrites."""
        assert host.flags.c_contiguous and host.dtype == NP_DTYPE[self.dtype]
        _check(load().b200_load_shell(self.h, slot, host.ctypes.data_as(C.c_void_p)))

    def interior_dead(self, slot: int) -> bool:
        return bool(load().b200_slot_interior_dead(self.test, slot))

    def run(self, niters: int) -> dict:
        st = Stats()
        _check(load().b200_run(self.h, niters, C.byref(st)))
        return dict(kernel_ms_per_sweep=st.kernel_ms_per_sweep, kernel_ms_total=st.kernel_ms_total,
                    regs=st.regs_per_thread, launches=st.launches, ngpus=st.ngpus,
                    kernel_name=st.kernel_name.decode() if st.kernel_name else "")

    def result_slot(self) -> int:
        return load().b200_result_slot(self.h)

    def save_array(self, slot: int, host: np.ndarray):
        assert host.flags.c_contiguous and host.dtype == NP_DTYPE[self.dtype]
        _check(load().b200_save(self.h, slot, host.ctypes.data_as(C.c_void_p)))

    def rewind(self):
        """Next load / run / save is a fresh job on the same buffers (rotation state as after plan())."""
        _check(load().b200_rewind(self.h))

    def set_async(self, on: bool):
        """Phase calls only enqueue (host arrays must be pinned and stay untouched until sync())."""
        _check(load().b200_set_async(self.h, 1 if on else 0))

    def sync(self):
        _check(load().b200_sync(self.h))

    def free(self):
        _check(load().b200_free(self.h))

    def destroy(self):
        if self.h:
            load().b200_destroy(self.h)
            self.h = C.c_void_p()

    # convenience: the whole driver loop on host arrays (in slot order); arrays are updated in place
    def run_on_host_arrays(self, test, dtype, nx, ny, ns, scalars, arrays, niters, shell_loads=True) -> tuple[int, dict]:
        self.plan(test, dtype, nx, ny, ns, scalars)
        self.alloc()
        try:
            for q, a in enumerate(arrays):
                # like the C drivers: output buffers only need their shell when a sweep follows
                if niters >= 1 and shell_loads and self.interior_dead(q):
                    self.load_array_shell(q, a)
                else:
                    self.load_array(q, a)
            stats = self.run(niters)
            slot = self.result_slot()
            for q, a in enumerate(arrays):
                self.save_array(q, a)
        finally:
            self.free()
        return slot, stats
