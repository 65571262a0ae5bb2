#!/usr/bin/env python
"""tools/scaling_suite.py -- per-stencil GLUP/s and % of the HBM roofline on N GPUs (BASELINE.json's metric:
"GLUP/s + % HBM roofline per stencil at 1/2/4/8 B200").  Run under torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/scaling_suite.py [out.json] [--size 512x256x256] [--niters 10]

Weak scaling: every rank holds one slab of the given size (z-slabs; y-slabs for the 2D tests), ghost planes are
pushed by the sweep kernels into the neighbours' memory over NVLink.  Timed on the device, max over ranks.
One process group for all tests (the start-up cost is paid once)."""
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    import torch
    import torch.distributed as dist
    from pkgload import load_pkg
    from bench import DEFAULT_SCALARS, measured_peaks
    pkg = load_pkg()
    pkg.load()
    from kernelgen_perf_tests_b200 import slab as slabmod
    args = [a for a in sys.argv[1:]]
    size, niters, out, only = "512x256x256", 10, None, None
    while args:
        a = args.pop(0)
        if a == "--size":
            size = args.pop(0)
        elif a == "--niters":
            niters = int(args.pop(0))
        elif a == "--tests":
            only = args.pop(0).split(",")
        else:
            out = a
    nx, ny, ns = [int(v) for v in size.split("x")]
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    peak, _ = measured_peaks()
    rows = []

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the first engine of a process group pays one-time costs (module load of the halo-push kernels, IPC
    # mappings, NCCL warm-up) that 10 warm-up sweeps do not always cover: run a throw-away test first
    tests = [t for t in pkg.TESTS if t != "matmul" and (only is None or t in only)]
    for i, test in enumerate([tests[0]] + tests):
        discard = i == 0
        info = pkg.test_info(test)
        for real in ("double", "float"):
            dims = (nx, ny, ns) if info["ndims"] == 3 else (nx, ny * ns, 1)
            try:
                eng = slabmod.SlabEngine(pkg, test, real, dims[0], dims[1], dims[2], DEFAULT_SCALARS.get(test, []), world=world,
                                         rank=rank, dist=dist if world > 1 else None, halo="push", seed=77 + rank)
                eng.run(niters)
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 3
                e0.record()
                for _ in range(reps):
                    eng.run(niters)
                e1.record()
                barrier()
                ms = e0.elapsed_time(e1)
                if world > 1:
                    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    ms = float(t.item())
                sec = ms * 1e-3 / (reps * niters)
                lups = eng.global_interior_points()
                bpl = (info["nread"] + info["nwritten"]) * (4 if real == "float" else 8)
                if discard:
                    eng.close()
                    continue
                rows.append({"test": test, "real": real, "n_gpus": world, "slab": "x".join(str(d) for d in dims),
                             "us_per_sweep": round(sec * 1e6, 2), "glups": round(lups / sec / 1e9, 1),
                             "frac_per_gpu": round(lups * bpl / sec / 1e9 / world / peak, 4),
                             "exchange": bool(world > 1 and info["exchange_slot"] >= 0)})
                eng.close()
            except Exception as e:      # noqa: BLE001
                rows.append({"test": test, "real": real, "n_gpus": world, "error": str(e)[:200]})
                torch.cuda.synchronize()
            if rank == 0 and rows and not discard:
                print(rows[-1], flush=True)
    if rank == 0 and out:
        Path(out).write_text(json.dumps(rows, indent=1))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
