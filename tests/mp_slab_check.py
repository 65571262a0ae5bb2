"""Multi-process check of the z-slab engine, launched by the tests as

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/mp_slab_check.py <backend> [halo]

backend = gloo : CPU.  Exercises the host logic only (SlabLayout + exchange_halos): each rank
                 sweeps its slab with the ORACLE (test infrastructure) and the gathered result
                 must equal the oracle run on the undivided grid, bit for bit.
backend = nccl : one GPU per rank.  SlabEngine (C-ABI sweeps, halo = push | nccl); the gathered
                 result must equal the single-GPU run of the same library, bit for bit.
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from pkgload import load_pkg  # noqa: E402

CASES = [("laplacian", (34, 18, 13)), ("wave13pt", (32, 20, 17)), ("lapgsrb", (36, 18, 15)),
         ("tricubic", (32, 16, 12)), ("uxx1", (32, 18, 11)), ("divergence", (32, 18, 11)),
         ("jacobi", (34, 41, 1)), ("gaussblur", (32, 37, 1)), ("gameoflife", (34, 29, 1))]
SCAL = {"laplacian": [0.3, 0.1], "wave13pt": [0.6, -0.03, 0.09], "lapgsrb": [0.5, 0.03, 0.02, -0.01],
        "uxx1": [0.4, -0.2], "divergence": [0.6, -0.2, 0.5], "jacobi": [0.5, 0.1, 0.02],
        "gaussblur": [0.6, 0.2, 0.1, 0.05, 0.03, 0.01]}
# nccl backend only: slabs large enough for several z-chunks / tile rows per CTA, so that the "ends first" walk of the halo-
# pushing kernels (interior units take no part in the neighbour ordering) is exercised, not only one-item grids
BIG_CASES = [("laplacian", (256, 96, 120)), ("wave13pt", (256, 60, 150)), ("lapgsrb", (128, 50, 90)), ("tricubic", (128, 40, 60)),
             ("jacobi", (256, 2500, 1)), ("gaussblur", (384, 1500, 1)), ("gameoflife", (256, 1800, 1))]
NT = 4


def global_arrays(info, nx, ny, n_split, real, seed):
    rng = np.random.default_rng(seed)
    n = nx * ny * n_split if info["ndims"] == 3 else nx * n_split
    return [rng.uniform(-1, 1, n).astype(np.float64 if real == "double" else np.float32) for _ in range(info["narrays"])]


def main():
    backend = sys.argv[1]
    halo = sys.argv[2] if len(sys.argv) > 2 else "nccl"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    pkg = load_pkg()
    from kernelgen_perf_tests_b200 import slab
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank))))
    else:
        dist.init_process_group("gloo")
    failures = []
    for real in ("double", "float"):
        for test, (nx, ny, per_rank) in CASES + (BIG_CASES if backend == "nccl" else []):
            info = pkg.test_info(test)
            sc = SCAL.get(test, [])
            split_local = per_rank if info["ndims"] == 3 else ny
            n_split = split_local * world
            full = global_arrays(info, nx, ny, n_split, real, 11)
            gdims = (nx, ny, n_split) if info["ndims"] == 3 else (nx, n_split, 1)
            if backend == "nccl":
                eng = slab.SlabEngine(pkg, test, real, nx, ny, per_rank, sc, world=world, rank=rank, dist=dist, halo=halo)
                for q, a in enumerate(full):
                    eng.set_global(q, a)
                torch.cuda.synchronize()
                dist.barrier()
                eng.run(NT)
                torch.cuda.synchronize()
                mine = [eng.get_owned(q) for q in range(info["narrays"])]
                if halo == "push" and info["exchange_slot"] >= 0:
                    # a SECOND job on the same engine, loaded the way the end-to-end leg loads (SlabEngine.load_host: output
                    # buffers receive only their boundary shell through b200_load_shell_slab, not even their ghost planes;
                    # the device buffers are NaN-poisoned first): must give the same arrays again
                    L = eng.layout
                    host = []
                    for q, a in enumerate(full):
                        part = torch.from_numpy(np.ascontiguousarray(a[L.mem_lo * eng.unit:L.mem_hi * eng.unit])).pin_memory()
                        host.append(part)
                    for t in eng.t:
                        t.fill_(float("nan"))
                    torch.cuda.synchronize()
                    dist.barrier()
                    eng.rewind()
                    eng.load_host(host)
                    eng.run(NT)
                    torch.cuda.synchronize()
                    again = [eng.get_owned(q) for q in range(info["narrays"])]
                    for q in range(info["narrays"]):
                        if not np.array_equal(again[q], mine[q]):
                            failures.append((test, real, q, "second job via load_host differs", int((again[q] != mine[q]).sum())))
                eng.close()
            else:
                from oracle_util import Oracle
                o = Oracle("strict")
                L = slab.SlabLayout(info, n_split, world, rank)
                unit = nx * ny if info["ndims"] == 3 else nx
                loc = [torch.from_numpy(a[L.mem_lo * unit:L.mem_hi * unit].copy()) for a in full]
                ldims = (nx, ny, L.mem_n) if info["ndims"] == 3 else (nx, L.mem_n, 1)
                idxs = [0, 1, 2]
                for _ in range(NT):
                    cur = [loc[idxs[q]] if q < info["rotation"] else loc[q] for q in range(info["narrays"])]
                    o.sweep(test, real, *ldims, sc, [t.numpy() for t in cur])
                    out_pos = 2 if info["rotation"] == 3 else 1
                    if info["exchange_slot"] >= 0 and world > 1:
                        slab.exchange_halos(dist, L, cur[out_pos].view(L.mem_n, -1))
                    if info["rotation"] == 2:
                        idxs[0], idxs[1] = idxs[1], idxs[0]
                    elif info["rotation"] == 3:
                        idxs = [idxs[1], idxs[2], idxs[0]]
                a0, b0 = (L.own_lo - L.mem_lo) * unit, (L.own_hi - L.mem_lo) * unit
                mine = [t[a0:b0].numpy() for t in loc]
            gathered = [None] * world
            dist.all_gather_object(gathered, mine)
            if rank == 0:
                got = [np.concatenate([g[q] for g in gathered]) for q in range(info["narrays"])]
                want = [a.copy() for a in full]
                if backend == "nccl":
                    ctx = pkg.Context(1)
                    ctx.run_on_host_arrays(test, real, *gdims, sc, want, NT)
                    ctx.destroy()
                else:
                    o.run(test, real, *gdims, NT, sc, want)
                for q in range(info["narrays"]):
                    if not np.array_equal(got[q], want[q]):
                        failures.append((test, real, q, int((got[q] != want[q]).sum())))
    dist.barrier()
    if rank == 0:
        print("MP_SLAB_CHECK", "FAIL " + repr(failures) if failures else f"OK backend={backend} halo={halo} world={world}")
    dist.destroy_process_group()
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
