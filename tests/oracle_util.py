"""ctypes access to the CPU oracle (oracle/libkgo*.so) and, when present, to the
reference sources compiled into oracle/_ref (libkgref_*.so).

TEST INFRASTRUCTURE.  Imported only by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
REF_DIR = ORACLE_DIR / "_ref"

TESTS = ["laplacian", "wave13pt", "divergence", "gradient", "uxx1", "lapgsrb", "jacobi",
         "gaussblur", "gameoflife", "tricubic", "tricubic2", "vecadd", "matvec", "sincos", "matmul"]
TEST_ID = {n: i for i, n in enumerate(TESTS)}
F32, F64 = 0, 1
NP_DTYPE = {F32: np.float32, F64: np.float64, "float": np.float32, "double": np.float64}
DT_ID = {"float": F32, "double": F64, np.float32: F32, np.float64: F64}


class _Info(C.Structure):
    _fields_ = [("name", C.c_char_p), ("ndims", C.c_int), ("narrays", C.c_int),
                ("nscalars", C.c_int), ("rotation", C.c_int)]


def build_oracle() -> None:
    """Compile oracle/libkgo*.so (and oracle/_ref when /root/reference exists)."""
    subprocess.run(["make", "-s", "-C", str(ORACLE_DIR), "all"], check=True)
    if Path("/root/reference").is_dir() and not (REF_DIR / "libkgref_shipped.so").exists():
        subprocess.run(["make", "-s", "-C", str(ORACLE_DIR), "ref"], check=True)


class Oracle:
    """One flavour of the restatement: 'fast' (libkgo.so), 'strict', or 'omp'."""

    def __init__(self, flavour: str = "fast"):
        name = {"fast": "libkgo.so", "strict": "libkgo_strict.so", "omp": "libkgo_omp.so"}[flavour]
        path = ORACLE_DIR / name
        if not path.exists():
            build_oracle()
        self.flavour = flavour
        self.lib = L = C.CDLL(str(path))
        L.kgo_info.restype = C.POINTER(_Info)
        L.kgo_info.argtypes = [C.c_int]
        L.kgo_array_len.restype = C.c_size_t
        L.kgo_array_len.argtypes = [C.c_int] * 5
        L.kgo_init.restype = C.c_double
        L.kgo_init.argtypes = [C.c_int] * 5 + [C.POINTER(C.c_double), C.POINTER(C.c_void_p)]
        L.kgo_sweep.restype = C.c_int
        L.kgo_sweep.argtypes = [C.c_int] * 5 + [C.POINTER(C.c_double), C.POINTER(C.c_void_p)]
        L.kgo_run.restype = C.c_int
        L.kgo_run.argtypes = [C.c_int] * 6 + [C.POINTER(C.c_double), C.POINTER(C.c_void_p)]
        L.kgo_final_mean.restype = C.c_double
        L.kgo_final_mean.argtypes = [C.c_int] * 5 + [C.POINTER(C.c_void_p), C.c_int]
        L.kgo_driver.restype = C.c_int
        L.kgo_driver.argtypes = [C.c_int] * 6 + [C.POINTER(C.c_double)] * 3
        L.kgo_max_threads.restype = C.c_int

    # -- helpers ---------------------------------------------------------
    def info(self, test: str):
        i = self.lib.kgo_info(TEST_ID[test]).contents
        return dict(name=i.name.decode(), ndims=i.ndims, narrays=i.narrays,
                    nscalars=i.nscalars, rotation=i.rotation)

    def array_len(self, test, slot, nx, ny, ns):
        return self.lib.kgo_array_len(TEST_ID[test], slot, nx, ny, ns)

    @staticmethod
    def _ptrs(arrays):
        return (C.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])

    def alloc(self, test, dtype, nx, ny, ns):
        n = self.info(test)["narrays"]
        return [np.empty(self.array_len(test, q, nx, ny, ns), dtype=NP_DTYPE[dtype]) for q in range(n)]

    def init(self, test, dtype, nx, ny, ns, reseed=True):
        """Reference-order rand() init.  Returns (scalars, arrays, i_mean)."""
        if reseed:
            self.lib.kgo_reseed()
        arrays = self.alloc(test, dtype, nx, ny, ns)
        sc = (C.c_double * 8)()
        im = self.lib.kgo_init(TEST_ID[test], DT_ID[dtype], nx, ny, ns, sc, self._ptrs(arrays))
        return list(sc)[: self.info(test)["nscalars"]], arrays, im

    def sweep(self, test, dtype, nx, ny, ns, scalars, arrays):
        sc = (C.c_double * 8)(*scalars)
        rc = self.lib.kgo_sweep(TEST_ID[test], DT_ID[dtype], nx, ny, ns, sc, self._ptrs(arrays))
        assert rc == 0

    def run(self, test, dtype, nx, ny, ns, nt, scalars, arrays) -> int:
        sc = (C.c_double * 8)(*scalars)
        slot = self.lib.kgo_run(TEST_ID[test], DT_ID[dtype], nx, ny, ns, nt, sc, self._ptrs(arrays))
        assert slot >= 0
        return slot

    def final_mean(self, test, dtype, nx, ny, ns, arrays, slot):
        return self.lib.kgo_final_mean(TEST_ID[test], DT_ID[dtype], nx, ny, ns, self._ptrs(arrays), slot)

    def driver(self, test, dtype, nx, ny, ns, nt):
        """Whole reference driver (rand init -> nt sweeps -> means).  (scalars, i_mean, f_mean)."""
        self.lib.kgo_reseed()
        sc = (C.c_double * 8)()
        im, fm = C.c_double(), C.c_double()
        rc = self.lib.kgo_driver(TEST_ID[test], DT_ID[dtype], nx, ny, ns, nt, sc, C.byref(im), C.byref(fm))
        assert rc == 0
        return list(sc), im.value, fm.value

    def max_threads(self):
        return self.lib.kgo_max_threads()


# ---------------------------------------------------------------------------
# oracle/_ref : the reference's own kernels, compiled from /root/reference
# ---------------------------------------------------------------------------
REF_C_TESTS = ["laplacian", "wave13pt", "divergence", "gradient", "uxx1", "lapgsrb",
               "gaussblur", "gameoflife", "tricubic", "tricubic2", "vecadd", "matvec"]


def _host_has_build_flags() -> bool:
    f = REF_DIR / "build_host_flags.txt"
    if not f.exists():
        return False
    need = set(f.read_text().split())
    try:
        line = next(l for l in open("/proc/cpuinfo") if l.startswith("flags"))
    except StopIteration:
        return False
    have = set(line.split(":", 1)[1].split())
    isa = {x for x in need if x.startswith(("avx", "amx", "fma", "bmi", "sse", "f16c", "movbe", "popcnt",
                                              "vaes", "vpclmul", "gfni", "sha", "adx", "lzcnt", "abm"))}
    return isa <= have


def ref_available(flavour: str = "shipped") -> bool:
    p = REF_DIR / f"libkgref_{flavour}.so"
    if not p.exists():
        return False
    if flavour == "shipped":
        return _host_has_build_flags()   # built -march=native
    return True


class RefKernels:
    """The reference's kernel functions (kgref_<test>_<f|d>) from oracle/_ref."""

    def __init__(self, flavour: str = "shipped"):
        if not ref_available(flavour):
            raise FileNotFoundError(f"oracle/_ref/libkgref_{flavour}.so not usable on this host")
        self.flavour = flavour
        self.lib = C.CDLL(str(REF_DIR / f"libkgref_{flavour}.so"))

    def sweep(self, test, dtype, nx, ny, ns, scalars, arrays):
        """Call the reference kernel with its own argument order
        (e.g. laplacian/laplacian.c:44-49; uxx1/uxx1.c:45-50)."""
        dt = DT_ID[dtype]
        real = C.c_float if dt == F32 else C.c_double
        fn = getattr(self.lib, f"kgref_{test}_{'f' if dt == F32 else 'd'}")
        fn.restype = None
        p = [C.c_void_p(a.ctypes.data) for a in arrays]
        s = [real(v) for v in scalars]
        i3 = [C.c_int(nx), C.c_int(ny), C.c_int(ns)]
        i2 = [C.c_int(nx), C.c_int(ny)]
        if test in ("laplacian", "lapgsrb", "wave13pt", "divergence", "gradient"):
            fn(*i3, *s, *p)
        elif test == "uxx1":      # uxx1(nx,ny,ns,c1,c2,u0,u1,d1,xx,xy,xz)
            fn(*i3, *s, *p)
        elif test in ("tricubic", "tricubic2", "vecadd"):
            fn(*i3, *p)
        elif test == "gaussblur":
            fn(*i2, *s, *p)
        elif test in ("gameoflife", "matvec"):
            fn(*i2, *p)
        else:
            raise KeyError(test)


def ref_binary(test: str, real: str) -> Path | None:
    p = REF_DIR / "bin" / f"{test}_{real}"
    if p.exists() and (_host_has_build_flags()):
        return p
    return None


def run_ref_binary(test: str, real: str, args, env_extra=None) -> dict:
    """Run a reference gcc-target driver and parse the stdout grammar
    benchmark:146-159,258-265 uses."""
    import re
    exe = ref_binary(test, real)
    assert exe is not None
    env = dict(os.environ)
    env.update(env_extra or {})
    out = subprocess.run([str(exe)] + [str(a) for a in args], capture_output=True, text=True, env=env,
                         check=True).stdout
    res = {"stdout": out}
    m = re.search(r"initial mean = (\S+)", out)
    if m:
        res["i_mean"] = float(m.group(1))
    m = re.search(r"final mean = (\S+)", out)
    if m:
        res["f_mean"] = float(m.group(1))
    m = re.search(r"compute time = (\S+) sec", out)
    if m:
        res["t_comp"] = float(m.group(1))
    return res
