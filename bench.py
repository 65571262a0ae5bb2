#!/usr/bin/env python
"""bench.py -- the headline benchmark of the b200 target.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch: `niters` sweeps (default 10, with the
reference driver's buffer rotation) of one stencil over one synthetic grid.  The N=1 workload
is BASELINE.json configs[1]: wave13pt, 512x256x256, double, niters=10.  For N>1 the grid is
weak-scaled in z (one 512x256x256 slab per GPU, z-slab decomposition, ghost planes pushed by
the sweep kernel itself into the neighbours' memory over NVLink).

Prints ONE JSON line (rank 0): GLUP/s (whole job), ms/step, the HBM roofline of the sweep
kernel, the end-to-end number through the C-ABI with host buffers, the CPU baseline (the
reference's own kernel compiled into oracle/_ref, timed on this box's host cores), clocks.
torch is used for device memory / streams / events / torch.distributed only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

BYTES = {"float": 4, "double": 8}
DTYPE_NAME = {"float": "f32", "double": "f64"}
# default-seed coefficients the reference drivers draw: first rand() values
DEFAULT_SCALARS = {
    "laplacian": [0.680375, -0.211234], "wave13pt": [0.680375, -0.211234 / 6, 0.566198 / 6],
    "divergence": [0.680375, -0.211234, 0.566198], "gradient": [0.680375, -0.211234, 0.566198],
    "uxx1": [0.680375, -0.211234], "lapgsrb": [0.680375, -0.211234 / 6, 0.566198 / 12, 0.59688 / 6],
    "jacobi": [0.680375, -0.211234 / 4, 0.566198 / 4],
    "gaussblur": [0.680375, -0.211234, 0.566198, 0.59688, 0.823295, -0.604897],
}


# interior of every test, (ndims, lo, hi): lo[d] <= idx < n[d] - hi[d] -- the loop bounds of the reference kernels
# (laplacian.c:81-99 ...).  Used by the reference arm, which must not map the product library.
INTERIOR = {
    "laplacian": (3, (1, 1, 1), (1, 1, 1)), "wave13pt": (3, (2, 2, 2), (2, 2, 2)), "divergence": (3, (1, 1, 1), (1, 1, 1)),
    "gradient": (3, (1, 1, 1), (1, 1, 1)), "uxx1": (3, (2, 2, 2), (1, 1, 1)), "lapgsrb": (3, (2, 2, 2), (2, 2, 2)),
    "jacobi": (2, (1, 1, 0), (1, 1, 0)), "gaussblur": (2, (2, 2, 0), (2, 2, 0)), "gameoflife": (2, (1, 1, 0), (1, 1, 0)),
    "tricubic": (3, (1, 1, 1), (2, 2, 2)), "tricubic2": (3, (2, 2, 2), (2, 2, 2)), "vecadd": (3, (0, 0, 0), (0, 0, 0)),
    "matvec": (2, (0, 0, 0), (0, 0, 0)), "sincos": (3, (0, 0, 0), (0, 0, 0)),
}
NARRAYS = {"laplacian": 2, "wave13pt": 3, "divergence": 4, "gradient": 4, "uxx1": 6, "lapgsrb": 2, "jacobi": 2,
           "gaussblur": 2, "gameoflife": 2, "tricubic": 5, "tricubic2": 5, "vecadd": 3, "matvec": 3, "sincos": 3}


def interior_points_py(test, nx, ny, ns):
    """Lattice-point updates per sweep (same number as b200_interior_points; asserted equal in tests/test_abi.py)."""
    nd, lo, hi = INTERIOR[test]
    ext = [nx, ny * ns if nd == 2 else ny, 1 if nd == 2 else ns]
    n = 1
    for d in range(3):
        n *= max(0, ext[d] - lo[d] - hi[d])
    return n


def make_config(test, real, nx, ny, ns, niters, world, halo):
    """The `config` object of the JSON line -- ONE function for both arms, so the driver sees the same dict."""
    return {"workload": f"{test} {nx}x{ny}x{ns} {real} niters={niters}" + (f" per GPU, z-slabs x{world}" if world > 1 else ""),
            "l2": f"inputs larger than L2 ({NARRAYS[test]} x {nx * ny * ns * BYTES[real] / 1e6:.0f} MB per GPU vs 126 MB)",
            "halo": halo if world > 1 else "none", "step": f"{niters} sweeps with buffer rotation"}


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--test", default="wave13pt")
    p.add_argument("--real", default="double", choices=["float", "double"])
    p.add_argument("--size", default="512x256x256", help="per-GPU grid nx x ny x ns")
    p.add_argument("--niters", type=int, default=10)
    p.add_argument("--suite", default="auto", choices=["auto", "none", "small", "full"],
                   help="also time every stencil (extra 'suite' key; N=1 only)")
    p.add_argument("--halo", default="push", choices=["push", "nccl"], help="N>1 ghost refresh")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu", action="store_true")
    p.add_argument("--no-parity", action="store_true", help="N>1: skip the multi == single bit-identity check")
    return p.parse_args()


def measured_peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            d = json.loads(f.read_text())
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[1]) for r in self.rows if len(r) > 2 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) > 2 and r[2].isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for n, v in zip(names, r[5:9]):
                if v.lower() == "active":
                    reasons.add(n)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


# ----------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own kernel (oracle/_ref), host cores
# ----------------------------------------------------------------------------------------------
def cpu_time_steps(test, real, nx, ny, ns, niters, steps, warmup):
    """Times `steps` x `niters` sweeps of the reference CPU kernel on all host threads.
    Returns (seconds per step, kind, cores, description)."""
    sys.path.insert(0, str(ROOT / "tests"))
    import numpy as np
    import oracle_util as ou
    o = ou.Oracle("omp")
    try:        # the OpenMP runtime may have been initialised (by torch) under torchrun's OMP_NUM_THREADS=1: set it explicitly
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(host_threads())
    except OSError:
        pass
    cores = o.max_threads()
    info = o.info(test)
    kind = "port"
    sweep = lambda cur: o.sweep(test, real, nx, ny, ns, scalars, cur)      # noqa: E731
    if test in ou.REF_C_TESTS and ou.ref_available("omp"):
        ref = ou.RefKernels("omp")
        kind = "reference"
        sweep = lambda cur: ref.sweep(test, real, nx, ny, ns, scalars, cur)   # noqa: E731
    scalars = DEFAULT_SCALARS.get(test, [])
    rng = np.random.default_rng(1)
    arrays = [rng.uniform(-1, 1, o.array_len(test, q, nx, ny, ns)).astype(ou.NP_DTYPE[real])
              for q in range(info["narrays"])]

    def one_step():
        cur = list(arrays)
        for _ in range(niters):
            sweep(cur)
            if info["rotation"] == 2:
                cur[0], cur[1] = cur[1], cur[0]
            elif info["rotation"] == 3:
                cur = [cur[1], cur[2], cur[0]]

    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    desc = (f"{steps} x {niters} sweeps of {test} {nx}x{ny}x{ns} {real}, "
            f"{'reference source + -fopenmp' if kind == 'reference' else 'oracle restatement + -fopenmp'}, {cores} threads")
    return dt, kind, cores, desc


def host_threads():
    """Host threads this process may use (affinity mask, not the machine total)."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def use_all_host_threads():
    """torch.distributed.run exports OMP_NUM_THREADS=1 to every rank; the CPU legs want all host threads.  Must run
    before the OpenMP runtime is loaded (libgomp reads the variable once), and the count is also set explicitly."""
    n = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(n)
    os.environ.pop("OMP_THREAD_LIMIT", None)
    return n


def run_reference_arm(args, nx, ny, ns):
    """--impl reference: the reference's own CPU kernel (oracle/_ref, built -fopenmp) on all host threads, with the
    caller's --steps / --warmup, the same `config` dict as the b200 arm.  Under torchrun rank 0 alone works.  The
    product library is NOT loaded here (interior points are computed in Python)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = max(1, int(os.environ.get("WORLD_SIZE", str(args.gpus))))
    use_all_host_threads()
    lups = interior_points_py(args.test, nx, ny, ns) * args.niters
    steps, warm = max(1, args.steps), max(0, args.warmup)
    dt, kind, cores, desc = cpu_time_steps(args.test, args.real, nx, ny, ns, args.niters, steps, warm)
    val = lups / dt / 1e9
    if world > 1:
        desc += f" (bounded sample: ONE of the {world} slabs of the weak-scaled grid; GLUP/s of a CPU sweep does not depend on ns)"
    line = {
        "impl": "reference", "metric": "GLUP/s", "value": val, "unit": "GLUP/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": DTYPE_NAME[args.real], "data": "synthetic",
        "config": make_config(args.test, args.real, nx, ny, ns, args.niters, world, args.halo),
        "cpu_baseline": {"value": val, "unit": "GLUP/s", "cores": cores, "kind": kind, "sample": desc},
        "e2e": {"value": val, "unit": "GLUP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# b200 arm
# ----------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    nx, ny, ns = [int(v) for v in args.size.lower().split("x")]
    if args.impl == "reference":
        run_reference_arm(args, nx, ny, ns)
        return

    import torch
    from pkgload import load_pkg
    pkg = load_pkg()
    pkg.load()                                   # fails loudly if the CUDA library is missing
    from kernelgen_perf_tests_b200 import slab as slabmod

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    test, real, niters = args.test, args.real, args.niters
    info = pkg.test_info(test)
    scalars = DEFAULT_SCALARS.get(test, [])
    eng = slabmod.SlabEngine(pkg, test, real, nx, ny, ns, scalars, world=world, rank=rank,
                             dist=dist, halo=args.halo, seed=1234 + rank)
    lups_step = eng.global_interior_points() * niters          # whole job, all ranks
    bytes_per_lup = (info["nread"] + info["nwritten"]) * BYTES[real]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        eng.run(niters)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = pkg.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        eng.run(niters)
    ev1.record()
    barrier()
    launches = pkg.launch_count() - l0
    ms_total = ev0.elapsed_time(ev1)
    if dist is not None:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = lups_step / (ms_step * 1e-3) / 1e9

    # roofline of the dominant (only) kernel: algorithmic bytes per launch / average launch duration
    peak, peak_src = measured_peaks()
    sweeps = args.steps * niters
    per_launch_lups = eng.local_interior_points()
    avg_launch_s = ms_total * 1e-3 / sweeps
    achieved = per_launch_lups * bytes_per_lup / avg_launch_s / 1e9
    # DRAM bytes per launch of the same kernel from an `ncu --set full` capture (dram__bytes_read + write, committed under
    # profiles/ by tools/collect_profiles.py; a number taken under a profiler cannot be measured inside this run).  Valid for
    # the exact (test, precision, per-GPU grid) key and for one GPU only: slabs launch a different kernel variant -> null.
    traffic = None
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists() and world == 1:
        try:
            traffic = json.loads(tf.read_text()).get(f"{test}_{real}_{nx}x{ny}x{ns}")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel": f"stream_kernel<{test}>",
                "bytes_per_lup": bytes_per_lup, "lups_per_launch": per_launch_lups,
                "avg_launch_us": avg_launch_s * 1e6}

    # end to end through the C ABI on HOST buffers (what a driver does): H2D of every array,
    # niters sweeps, D2H of the result array -- all inside the timed region, every step.
    e2e = None
    if not args.no_e2e:
        e2e = eng.e2e(niters, steps=max(3, min(args.steps, 10)), barrier=barrier, dist=dist)
        e2e["value"] = lups_step / e2e.pop("seconds_per_step") / 1e9
        e2e["unit"] = "GLUP/s"

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        dt, kind, cores, desc = cpu_time_steps(test, real, nx, ny, ns, niters, 3, 1)
        cpu = {"value": pkg.interior_points(test, nx, ny, ns) * niters / dt / 1e9, "unit": "GLUP/s",
               "cores": cores, "kind": kind, "sample": desc}

    # parity carried by the line itself (N > 1): the slab run == a single-GPU run of the same global grid, bit for bit --
    # the headline test at the bench size, plus one 2D test (row push).  Outside the timed region.
    parity = None
    if world > 1 and info["exchange_slot"] >= 0 and not args.no_parity:
        p3 = slabmod.multi_eq_single(pkg, dist, test, real, nx, ny, ns, scalars, niters, world, rank, halo=args.halo)
        p2 = slabmod.multi_eq_single(pkg, dist, "gameoflife", "double", 1024, 4096, 1, [], niters, world, rank, halo=args.halo)
        if rank == 0:
            parity = {"multi_eq_single": bool(p3["multi_eq_single"] and p2["multi_eq_single"]),
                      "bytes_compared": p3["bytes_compared"] + p2["bytes_compared"], "checks": [p3, p2]}

    suite = None
    want_suite = args.suite if args.suite != "auto" else ("full" if world == 1 else "none")
    if rank == 0 and world == 1 and want_suite != "none":
        def cpu_row(t, r, a, b, c, nit):
            dt, kind, cores, _ = cpu_time_steps(t, r, a, b, c, nit, 1, 1)
            return {"cpu_glups": round(pkg.interior_points(t, a, b, c) * nit / dt / 1e9, 3), "cpu_kind": kind, "cpu_cores": cores}

        suite = slabmod.suite_table(pkg, peak, full=(want_suite == "full"), scalars=DEFAULT_SCALARS,
                                    cpu_fn=cpu_row if (want_suite == "full" and not args.no_cpu) else None)
        if want_suite == "full":
            suite += slabmod.matmul_table(pkg)

    if rank == 0:
        line = {
            "metric": "GLUP/s", "value": value, "unit": "GLUP/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": DTYPE_NAME[real], "data": "synthetic",
            "config": make_config(test, real, nx, ny, ns, niters, world, args.halo),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu,
        }
        if parity is not None:
            line["parity"] = parity
        if suite is not None:
            line["suite"] = suite
        print(json.dumps(line), flush=True)
    eng.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
