"""GPU parity tests proper: the CUDA path, called through the C ABI exactly as the C drivers
call it (b200_init / plan / alloc / load / run / save on HOST buffers), against
  * the committed golden fixtures (outputs of the reference itself, tests/golden/), and
  * the CPU oracle on the same rand()-initialised inputs, at ragged / odd / TMA-incompatible
    sizes and at the README size 512x256x256.
Bars: gameoflife bit-exact vs the strict-IEEE build; every other test normwise
max|gpu-ref|/max|ref| <= 1e-12 (double) / 1e-5 (float) (FMA contraction + re-association);
uxx1 with the condition-aware per-point bound of parity_util.uxx1_bound.
"""
from pathlib import Path

import numpy as np
import pytest

from parity_util import TOL, normwise, uxx1_bound

pytestmark = pytest.mark.gpu

GOLDEN_DIR = Path(__file__).resolve().parent / "golden"
STENCILS = ["laplacian", "wave13pt", "divergence", "gradient", "uxx1", "lapgsrb", "jacobi",
            "gaussblur", "gameoflife", "tricubic", "tricubic2", "vecadd", "matvec", "sincos"]
FIXTURE_TESTS = [t for t in STENCILS if t not in ("jacobi", "sincos")]


@pytest.fixture(scope="module")
def ctx(pkg):
    c = pkg.Context(1)
    yield c
    c.destroy()


def gpu_run(ctx, test, real, nx, ny, ns, nt, scalars, arrays):
    work = [a.copy() for a in arrays]
    slot, stats = ctx.run_on_host_arrays(test, real, nx, ny, ns, scalars, work, nt)
    return slot, work, stats


def check(test, real, nx, ny, ns, nt, scalars, inputs, got, want, strict_want=None):
    if test == "gameoflife":
        ref = strict_want if strict_want is not None else want
        for q in range(len(got)):
            assert np.array_equal(got[q], ref[q]), f"gameoflife/{real}: slot {q} not bit-exact vs strict oracle"
        return
    if test == "uxx1":
        bound = uxx1_bound(scalars, inputs, nx, ny, ns, real, nt)
        for q in (0, 1):
            err = np.abs(got[q].astype(np.float64) - want[q].astype(np.float64))
            bad = err > bound + 1e-300
            assert not bad.any(), (f"uxx1/{real}: slot {q}: {int(bad.sum())} points beyond the condition-aware bound; "
                                   f"worst ratio {float(np.max(err[bad] / np.maximum(bound[bad], 1e-300))):.3g}")
        for q in range(2, 6):
            assert np.array_equal(got[q], want[q])
        return
    for q in range(len(got)):
        e = normwise(got[q], want[q])
        assert e <= TOL[real], f"{test}/{real} {nx}x{ny}x{ns} nt={nt}: slot {q} normwise error {e:.3e}"


@pytest.mark.parametrize("real", ["float", "double"])
@pytest.mark.parametrize("test", FIXTURE_TESTS)
def test_golden_fixtures(ctx, pkg, test, real):
    """Against outputs of the reference itself (shipped flags and strict build)."""
    fx = np.load(GOLDEN_DIR / f"{test}_{real}.npz")
    nx, ny, ns, nt = [int(v) for v in fx["dims"]]
    scalars = [float(v) for v in fx["scalars"]]
    n = len([k for k in fx.files if k.startswith("in")])
    inputs = [fx[f"in{q}"] for q in range(n)]
    slot, got, stats = gpu_run(ctx, test, real, nx, ny, ns, nt, scalars, inputs)
    shipped = [fx[f"shipped{q}"] if f"shipped{q}" in fx.files else fx[f"in{q}"] for q in range(n)]
    strict = [fx[f"strict{q}"] if f"strict{q}" in fx.files else fx[f"in{q}"] for q in range(n)]
    check(test, real, nx, ny, ns, nt, scalars, inputs, got, shipped, strict)
    if test != "gameoflife":
        check(test, real, nx, ny, ns, nt, scalars, inputs, got, strict)
    else:
        # vs the shipped (fast-math) build only a pointwise-relative bound is meaningful
        rtol = 1e-8 if real == "double" else 1e-2
        assert np.allclose(got[0], shipped[0], rtol=rtol, atol=0) and np.allclose(got[1], shipped[1], rtol=rtol, atol=0)
    pairs = (nt - 2) // 2 if (nt >= 4 and pkg.capi.sweep2_profitable(test, nx)) else 0
    assert stats["launches"] == nt - pairs          # a fused two-sweep pass is one launch


@pytest.mark.parametrize("real", ["float", "double"])
@pytest.mark.parametrize("test", ["jacobi", "sincos"])
def test_golden_fixtures_fortran_tests(ctx, test, real):
    """jacobi and sincos are Fortran in the reference (no gfortran in the build container): their fixtures hold the outputs
    of the reference's DECLARATIVE definitions (jacobi/jacobi.stc:1-13, sincos/sincos.stc) evaluated by tests/stc_eval.py
    (generator: tests/golden/make_golden_stc.py), on inputs drawn in the reference drivers' rand() order."""
    fx = np.load(GOLDEN_DIR / f"{test}_{real}.npz")
    nx, ny, ns, nt = [int(v) for v in fx["dims"]]
    scalars = [float(v) for v in fx["scalars"]]
    n = len([k for k in fx.files if k.startswith("in")])
    inputs = [fx[f"in{q}"] for q in range(n)]
    slot, got, stats = gpu_run(ctx, test, real, nx, ny, ns, nt, scalars, inputs)
    want = [fx[f"stc{q}"] if f"stc{q}" in fx.files else fx[f"in{q}"] for q in range(n)]
    check(test, real, nx, ny, ns, nt, scalars, inputs, got, want)


SIZES_3D = [(128, 20, 12), (130, 37, 29), (63, 31, 29), (260, 19, 11), (16, 5, 5), (5, 5, 5)]
SIZES_2D = [(128, 70, 1), (130, 517, 1), (63, 301, 1), (512, 300, 1), (5, 5, 1)]


@pytest.mark.parametrize("real", ["float", "double"])
@pytest.mark.parametrize("test", STENCILS)
def test_vs_oracle_ragged_sizes(ctx, oracle, oracle_strict, test, real):
    """Same rand() inputs, odd / ragged / tiny extents (TMA path and scalar-loader path), 3 sweeps
    with the driver's buffer rotation, every array compared (untouched shells included)."""
    info = oracle.info(test)
    for nx, ny, ns in (SIZES_3D if info["ndims"] == 3 else SIZES_2D):
        for nt in (1, 3):
            scalars, inputs, _ = oracle.init(test, real, nx, ny, ns)
            o = oracle_strict if test == "gameoflife" else oracle
            want = [a.copy() for a in inputs]
            slot_o = o.run(test, real, nx, ny, ns, nt, scalars, want)
            slot, got, stats = gpu_run(ctx, test, real, nx, ny, ns, nt, scalars, inputs)
            assert slot == slot_o, f"{test}: result slot {slot} != reference remap {slot_o}"
            check(test, real, nx, ny, ns, nt, scalars, inputs, got, want)


@pytest.mark.parametrize("test", ["laplacian", "wave13pt", "lapgsrb", "uxx1", "divergence", "gradient",
                                  "gameoflife", "gaussblur", "jacobi", "vecadd", "matvec"])
def test_readme_size_double(ctx, pkg, test):
    """512 256 256 (2D: 512 x 65536), double, 2 sweeps, element-wise vs the threaded oracle, and the
    driver-level checksums: i_mean / f_mean the way the reference prints them."""
    from oracle_util import Oracle
    o = Oracle("omp")
    os_ = Oracle("strict")
    info = o.info(test)
    nx, ny, ns = (512, 256, 256) if info["ndims"] == 3 else (512, 65536, 1)
    nt = 2
    scalars, inputs, _ = o.init(test, "double", nx, ny, ns)
    want = [a.copy() for a in inputs]
    (os_ if test == "gameoflife" else o).run(test, "double", nx, ny, ns, nt, scalars, want)
    slot, got, stats = gpu_run(ctx, test, "double", nx, ny, ns, nt, scalars, inputs)
    check(test, "double", nx, ny, ns, nt, scalars, inputs, got, want)
    fm_gpu = o.final_mean(test, "double", nx, ny, ns, got, slot)
    fm_ref = o.final_mean(test, "double", nx, ny, ns, want, slot)
    assert "%f" % fm_gpu == "%f" % fm_ref


FUSED_SIZES = [(128, 70, 1), (130, 517, 1), (63, 301, 1), (512, 300, 1), (1024, 1203, 1), (248, 92, 1), (5, 5, 1), (9, 4, 1)]


@pytest.mark.parametrize("real", ["float", "double"])
@pytest.mark.parametrize("test", ["jacobi", "gaussblur", "gameoflife"])
def test_fused_two_sweep_passes(pkg, oracle, oracle_strict, test, real, monkeypatch):
    """Temporal blocking (b200_ops2d.cuh Fused2D, used by b200_run for the first nt-2 sweeps): bit-identical to
    single sweeps (B200_FUSE=0), within the bar of the oracle; every array compared, buffers NaN-poisoned.
    Sizes cover tile seams of the overlapped tiling (pitch 124 / 120 x 46 / 44 / 94 / 92), the scalar-loader
    path (odd nx) and grids smaller than one tile."""
    monkeypatch.setenv("B200_POISON", "1")
    c = pkg.Context(1)
    o = oracle_strict if test == "gameoflife" else oracle
    try:
        for nx, ny, ns in FUSED_SIZES:
            for nt in (4, 7, 10):
                scalars, inputs, _ = oracle.init(test, real, nx, ny, ns)
                monkeypatch.setenv("B200_FUSE", "1")
                a = [x.copy() for x in inputs]
                slot_a, st_a = c.run_on_host_arrays(test, real, nx, ny, ns, scalars, a, nt)
                monkeypatch.setenv("B200_FUSE", "0")
                b = [x.copy() for x in inputs]
                slot_b, st_b = c.run_on_host_arrays(test, real, nx, ny, ns, scalars, b, nt)
                if pkg.interior_points(test, nx, ny, ns):      # an empty interior launches nothing
                    assert st_b["launches"] == nt and st_a["launches"] == nt - (nt - 2) // 2, (st_a, st_b)
                assert slot_a == slot_b
                for q in range(len(a)):
                    assert np.array_equal(a[q], b[q]), f"{test}/{real} {nx}x{ny} nt={nt}: slot {q}: fused != single sweeps"
                want = [x.copy() for x in inputs]
                assert o.run(test, real, nx, ny, ns, nt, scalars, want) == slot_a
                check(test, real, nx, ny, ns, nt, scalars, inputs, a, want)
    finally:
        c.destroy()


MATMUL_SIZES = [(128, 128, 128), (256, 64, 384), (130, 37, 29), (131, 67, 259), (16, 5, 5), (5, 5, 5), (1, 1, 1), (300, 513, 140)]


@pytest.mark.parametrize("real", ["float", "double"])
def test_matmul_vs_oracle(ctx, oracle, real):
    """matmul/matmul.F90:56-68 -- C += A*B over nt sweeps (C accumulates, matmul/main.c:232-244):
    the hand-written tensor-core kernels (aligned and element-wise loaders, ragged tiles) against the
    sequential-sum restatement; also with a non-zero initial C.  (The product library has no library-GEMM
    path: the cuBLAS baseline lives in libb200stencil_diag.so and in bench.py's torch.matmul timing.)"""
    mode = "tensor"
    rng = np.random.default_rng(5)
    for nx, ny, ns in MATMUL_SIZES:
        for nt in (1, 3):
            scalars, inputs, _ = oracle.init("matmul", real, nx, ny, ns)
            if nt == 3:
                inputs[2][:] = rng.uniform(-1, 1, inputs[2].size).astype(inputs[2].dtype)
            want = [a.copy() for a in inputs]
            slot_o = oracle.run("matmul", real, nx, ny, ns, nt, scalars, want)
            slot, got, stats = gpu_run(ctx, "matmul", real, nx, ny, ns, nt, scalars, inputs)
            assert slot == slot_o == 2
            assert np.array_equal(got[0], inputs[0]) and np.array_equal(got[1], inputs[1])
            e = normwise(got[2], want[2])
            assert e <= TOL[real], f"matmul/{real}/{mode} {nx}x{ny}x{ns} nt={nt}: normwise error {e:.3e}"


def test_matmul_1024_float_accuracy(ctx):
    """3xTF32 keeps FP32-level accuracy where one TF32 pass would not: 1024^3 against a float64
    product of the same float32 inputs (a size the sequential oracle would take too long on)."""
    rng = np.random.default_rng(11)
    n = 1024
    A = rng.uniform(-1, 1, n * n).astype(np.float32)
    B = rng.uniform(-1, 1, n * n).astype(np.float32)
    Cm = np.zeros(n * n, np.float32)
    slot, got, _ = gpu_run(ctx, "matmul", "float", n, n, n, 1, [], [A, B, Cm])
    ref = (A.reshape(n, n).T.astype(np.float64) @ B.reshape(n, n).T.astype(np.float64)).T.reshape(-1)
    e = float(np.max(np.abs(got[2] - ref)) / np.max(np.abs(ref)))
    assert e <= 2e-6, e
    Ad, Bd = A.astype(np.float64), B.astype(np.float64)
    slot, gotd, _ = gpu_run(ctx, "matmul", "double", n, n, n, 1, [], [Ad, Bd, np.zeros(n * n)])
    ed = float(np.max(np.abs(gotd[2] - ref)) / np.max(np.abs(ref)))
    assert ed <= 1e-14, ed


@pytest.mark.parametrize("test", STENCILS)
def test_shell_only_loads_poisoned(pkg, oracle, oracle_strict, test, monkeypatch):
    """b200_load_shell: output buffers receive only their boundary shell.  Device buffers start as NaN
    patterns (B200_POISON), so a missed shell point or an unwritten interior point cannot hide; result
    must equal the run with whole-array loads bit for bit, and the oracle within the bar."""
    monkeypatch.setenv("B200_POISON", "1")
    info = oracle.info(test)
    c = pkg.Context(1)
    try:
        for real in ("double", "float"):
            for nx, ny, ns in ([(132, 21, 13), (64, 9, 7), (5, 5, 5)] if info["ndims"] == 3 else [(132, 75, 1), (8, 9, 1)]):
                scalars, inputs, _ = oracle.init(test, real, nx, ny, ns)
                a = [x.copy() for x in inputs]
                b = [x.copy() for x in inputs]
                c.run_on_host_arrays(test, real, nx, ny, ns, scalars, a, 3, shell_loads=True)
                c.run_on_host_arrays(test, real, nx, ny, ns, scalars, b, 3, shell_loads=False)
                for q in range(len(a)):
                    assert np.array_equal(a[q], b[q]), f"{test}/{real} {nx}x{ny}x{ns}: slot {q} differs with shell-only loads"
                want = [x.copy() for x in inputs]
                (oracle_strict if test == "gameoflife" else oracle).run(test, real, nx, ny, ns, 3, scalars, want)
                check(test, real, nx, ny, ns, 3, scalars, inputs, a, want)
    finally:
        c.destroy()


def test_degenerate_and_errors(ctx, pkg):
    a = [np.ones(27), np.ones(27) * 2]
    slot, got, _ = gpu_run(ctx, "wave13pt", "double", 3, 3, 3, 2, [0.1, 0.2, 0.3], a + [np.ones(27) * 3])
    assert np.array_equal(got[0], a[0])            # interior empty: nothing written
    with pytest.raises(pkg.B200Error):
        ctx.plan("laplacian", "double", 8, 8, 8, [0.1])       # wrong scalar count
    with pytest.raises(pkg.B200Error):
        ctx.plan("laplacian", "double", -1, 8, 8, [0.1, 0.2])


def test_kernel_info(pkg):
    for t in STENCILS:
        for real in ("float", "double"):
            ki = pkg.kernel_info(t, real)
            assert 16 <= ki["regs"] <= 255 and ki["name"] == t


@pytest.mark.parametrize("test,real,dims", [("wave13pt", "double", (130, 37, 29)), ("laplacian", "float", (128, 48, 40)),
                                            ("jacobi", "double", (256, 301, 1)), ("gradient", "float", (64, 31, 29))])
def test_async_contexts_match_sync(pkg, oracle, test, real, dims):
    """b200_set_async / b200_sync: two contexts whose phase calls only enqueue, driven alternately over several jobs
    with different inputs (what bench.py's e2e leg does), give bit for bit what the synchronous phases give."""
    nx, ny, ns = dims
    nt, jobs = 5, 4
    info = pkg.test_info(test)
    dt = np.float64 if real == "double" else np.float32
    scalars = [0.4, -0.05, 0.03, 0.02, 0.01, 0.005][:info["nscalars"]]
    rng = np.random.default_rng(11)
    inputs = [[rng.uniform(-1, 1, oracle.array_len(test, q, nx, ny, ns)).astype(dt) for q in range(info["narrays"])]
              for _ in range(jobs)]
    want = []
    sync_ctx = pkg.Context(1)
    try:
        for j in range(jobs):
            w = [a.copy() for a in inputs[j]]
            slot, _ = sync_ctx.run_on_host_arrays(test, real, nx, ny, ns, scalars, w, nt)
            want.append((slot, w))
    finally:
        sync_ctx.destroy()
    lanes = []
    for _ in range(2):
        c = pkg.Context(1)
        c.plan(test, real, nx, ny, ns, scalars)
        c.alloc()
        c.set_async(True)
        host = [pkg.capi.PinnedBuffer(a.size, dt) for a in inputs[0]]
        lanes.append((c, host, []))
    try:
        def collect(lane):
            c, host, pending = lane
            c.sync()
            for j, slot in pending:
                assert slot == want[j][0]
                assert np.array_equal(host[slot].array, want[j][1][slot]), f"{test}/{real}: job {j} differs in async mode"
            pending.clear()
        for j in range(jobs):
            lane = lanes[j % 2]
            collect(lane)                                  # the lane's previous job has left its host buffers
            c, host, pending = lane
            c.rewind()
            for q, h in enumerate(host):
                h.array[:] = inputs[j][q]
                if c.interior_dead(q):
                    c.load_array_shell(q, h.array)
                else:
                    c.load_array(q, h.array)
            c.run(nt)
            slot = c.result_slot()
            c.save_array(slot, host[slot].array)
            pending.append((j, slot))
        for lane in lanes:
            collect(lane)
    finally:
        for c, host, _ in lanes:
            c.set_async(False)
            c.free()
            c.destroy()
            for h in host:
                h.free()


def test_tricubic_row_variants():
    """Every form of the tricubic kernel that was measured on the way to the product form (k_tricubic.cu: 1-6 the
    separable x-y-z forms, 7-10 the outer-product forms) against the oracle.  They live in the DIAGNOSTICS library
    (libb200stencil_diag.so, B200_TRICUBIC_ROWS); the product library holds form 7 only and ignores the variable -- it is
    run too ("0").  Each form runs in its own interpreter (tests/tricubic_variant_check.py: tricubic + tricubic2, both
    precisions, ragged / odd / multi-tile sizes)."""
    import json
    import os
    import subprocess
    import sys
    root = Path(__file__).resolve().parent.parent
    diag = root / "kernelgen-perf-tests_b200" / "libb200stencil_diag.so"
    assert diag.exists(), "diagnostics library missing: make -C kernelgen-perf-tests_b200/csrc diag (build() does it)"
    out = {}
    for rows in ("0", "1", "2", "4", "6", "7", "8", "9"):
        env = dict(os.environ, B200_TRICUBIC_ROWS=rows)
        if rows != "0":
            env["B200_LIB"] = str(diag)
        else:
            env.pop("B200_LIB", None)
        p = subprocess.run([sys.executable, str(Path(__file__).resolve().parent / "tricubic_variant_check.py")],
                           capture_output=True, text=True, env=env, timeout=900)
        assert p.returncode == 0, f"rows={rows}: {p.stdout[-2000:]} {p.stderr[-2000:]}"
        out[rows] = json.loads(p.stdout.strip().splitlines()[-1])
        for real in ("float", "double"):
            assert out[rows]["worst"][real] <= TOL[real], out[rows]
    # the product library runs form 7: same bytes as the diagnostics library's form 7
    assert out["0"]["sha"] == out["7"]["sha"]
    assert len({out[r]["sha"] for r in ("1", "2", "4", "6")}) == 1 and len({out[r]["sha"] for r in ("7", "8", "9")}) == 1


@pytest.mark.parametrize("real", ["float", "double"])
@pytest.mark.parametrize("test", ["jacobi", "gaussblur", "gameoflife"])
def test_repeated_odd_runs_on_fused_context(pkg, oracle, oracle_strict, test, real, monkeypatch):
    """b200_run called again after an ODD number of sweeps on a context with a scratch buffer: the roles of
    the two buffers are then swapped, and the scratch carries the shell of the other one; b200_run realigns them with one
    leading single sweep.  run(5) + run(5) and run(3) + run(7) must equal run(10) bit for bit (fused and unfused), and the
    oracle within the bar."""
    nx, ny, ns = 260, 150, 1
    scalars, inputs, _ = oracle.init(test, real, nx, ny, ns)
    results = {}
    for fuse in ("1", "0"):
        monkeypatch.setenv("B200_FUSE", fuse)
        monkeypatch.setenv("B200_POISON", "1")
        for split in ((10,), (5, 5), (3, 7), (1, 4, 5)):
            c = pkg.Context(1)
            try:
                c.plan(test, real, nx, ny, ns, scalars)
                c.alloc()
                work = [a.copy() for a in inputs]
                for q, a in enumerate(work):
                    c.load_array(q, a)
                for n in split:
                    c.run(n)
                slot = c.result_slot()
                for q, a in enumerate(work):
                    c.save_array(q, a)
                results[(fuse, split)] = (slot, work)
            finally:
                c.free()
                c.destroy()
    base_slot, base = results[("0", (10,))]
    for key, (slot, work) in results.items():
        assert slot == base_slot, key
        for q in range(len(work)):
            assert np.array_equal(work[q], base[q]), f"{test}/{real}: B200_FUSE={key[0]} runs {key[1]}: slot {q} differs from run(10)"
    want = [a.copy() for a in inputs]
    (oracle_strict if test == "gameoflife" else oracle).run(test, real, nx, ny, ns, 10, scalars, want)
    check(test, real, nx, ny, ns, 10, scalars, inputs, base, want)


@pytest.mark.parametrize("real,tol", [("float", 1e-5), ("double", 1e-12)])
def test_matmul_8192(pkg, real, tol):
    """BASELINE configs[4]: matmul 8192^2 operands (8192^3 multiply-adds), C += A*B through the C ABI on device buffers;
    block-sampled rows and columns of C against a float64 product of the same inputs (the sequential oracle would need
    hours): 64 whole rows + 64 whole columns incl. the first/last of the matrix and tile seams (127/128, 4095/4096).
    Normwise bar 1e-5 (float) / 1e-12 (double) as the north star states."""
    import torch
    n = 8192
    dt = torch.float32 if real == "float" else torch.float64
    g = torch.Generator(device="cuda")
    g.manual_seed(3)
    A = torch.empty(n * n, device="cuda", dtype=dt).uniform_(-1, 1, generator=g)      # column-major nx x ny
    B = torch.empty(n * n, device="cuda", dtype=dt).uniform_(-1, 1, generator=g)      # column-major ny x ns
    C0 = torch.empty(n * n, device="cuda", dtype=dt).uniform_(-1, 1, generator=g)
    Cm = C0.clone()
    pkg.capi.sweep_loop("matmul", real, n, n, n, [], [A.data_ptr(), B.data_ptr(), Cm.data_ptr()], 1,
                        stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    Am, Bm = A.view(n, n).t().double(), B.view(n, n).t().double()        # A[i,k], B[k,j]
    Cg, Cs = Cm.view(n, n).t().double(), C0.view(n, n).t().double()      # C[i,j]
    rng = np.random.default_rng(5)
    rows = sorted({0, 1, 127, 128, 4095, 4096, n - 2, n - 1} | set(int(v) for v in rng.integers(0, n, 56)))
    cols = sorted({0, 1, 127, 128, 4095, 4096, n - 2, n - 1} | set(int(v) for v in rng.integers(0, n, 56)))
    ref_r = Cs[rows, :] + Am[rows, :] @ Bm
    ref_c = Cs[:, cols] + Am @ Bm[:, cols]
    scale = float(torch.max(torch.abs(ref_r)).item())
    e = max(float(torch.max(torch.abs(Cg[rows, :] - ref_r)).item()), float(torch.max(torch.abs(Cg[:, cols] - ref_c)).item())) / scale
    assert e <= tol, f"matmul/{real} 8192^3: normwise error {e:.3e}"


@pytest.mark.parametrize("test", ["tricubic", "tricubic2", "sincos"])
def test_readme_size_whole_grid(ctx, pkg, test):
    """512 256 256, whole grid, both precisions, 2 sweeps vs the threaded oracle -- the tests that
    test_readme_size_double leaves out (no whole-grid C1 check for tricubic / tricubic2 / sincos)."""
    from oracle_util import Oracle
    o = Oracle("omp")
    nx, ny, ns, nt = 512, 256, 256, 2
    for real in ("double", "float"):
        scalars, inputs, _ = o.init(test, real, nx, ny, ns)
        want = [a.copy() for a in inputs]
        o.run(test, real, nx, ny, ns, nt, scalars, want)
        slot, got, stats = gpu_run(ctx, test, real, nx, ny, ns, nt, scalars, inputs)
        check(test, real, nx, ny, ns, nt, scalars, inputs, got, want)


def test_async_refused_on_multi_gpu_context(pkg):
    """b200_set_async is a single-GPU facility: a multi-GPU context must refuse it."""
    if pkg.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    c = pkg.Context(2)
    try:
        with pytest.raises(pkg.B200Error):
            c.set_async(True)
    finally:
        c.destroy()
