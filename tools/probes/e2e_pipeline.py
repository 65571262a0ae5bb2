"""Why does the two-context e2e leg of bench.py hide only the sweeps and not the return copy?  Times the same job
(wave13pt 512x256x256 double: 2 arrays + a shell up, 10 sweeps, 1 array down) per step
  (a) with plain torch streams: H2D -> a small kernel -> D2H on each of `depth` streams, alternating, and
  (b) through the C ABI contexts in asynchronous mode, depth 1..3,
so the two can be compared on one box."""
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
nx, ny, ns, niters = 512, 256, 256, 10
n = nx * ny * ns


def run(step, sync_all, depth, steps=12):
    for i in range(depth):
        step(i)
    sync_all()
    t0 = time.perf_counter()
    for i in range(steps):
        step(i)
    sync_all()
    return (time.perf_counter() - t0) / steps * 1e3


# (a) torch streams
for depth in ((2,) if os.environ.get("PROBE_SHORT") else (1, 2, 3)):
    lanes = []
    for _ in range(depth):
        lanes.append(dict(s=torch.cuda.Stream(), h_in=torch.empty(2 * n, dtype=torch.float64, pin_memory=True),
                          h_out=torch.empty(n, dtype=torch.float64, pin_memory=True),
                          d_in=torch.empty(2 * n, dtype=torch.float64, device="cuda"), d_out=torch.empty(n, dtype=torch.float64, device="cuda")))

    def step(i):
        L = lanes[i % depth]
        L["s"].synchronize()
        with torch.cuda.stream(L["s"]):
            L["d_in"].copy_(L["h_in"], non_blocking=True)
            for _ in range(4):
                torch.add(L["d_in"][:n], L["d_in"][n:], out=L["d_out"])
            L["h_out"].copy_(L["d_out"], non_blocking=True)

    print(f"torch streams, depth {depth}: {run(step, torch.cuda.synchronize, depth):.2f} ms/step", flush=True)
    del lanes

# (b) C ABI contexts
for depth in ((2,) if os.environ.get("PROBE_SHORT") else (1, 2, 3)):
    lanes = []
    for _ in range(depth):
        host = [capi.PinnedBuffer(n, np.float64) for _ in range(3)]
        for h in host:
            h.array[:] = 0.5
        ctx = capi.Context(1)
        ctx.plan("wave13pt", "double", nx, ny, ns, [0.1, 0.2, 0.3])
        ctx.alloc()
        ctx.set_async(True)
        lanes.append((ctx, host))
    dead = [lanes[0][0].interior_dead(q) for q in range(3)]

    def step(i, parts=("load", "run", "save")):
        ctx, hb = lanes[i % depth]
        ctx.sync()
        ctx.rewind()
        if "load" in parts:
            for q, h in enumerate(hb):
                (ctx.load_array_shell if dead[q] else ctx.load_array)(q, h.array)
        if "run" in parts:
            ctx.run(niters)
        if "save" in parts:
            slot = ctx.result_slot()
            ctx.save_array(slot, hb[slot].array)

    def sync_all():
        for ctx, _ in lanes:
            ctx.sync()
        torch.cuda.synchronize()

    print(f"C ABI contexts, depth {depth}: {run(step, sync_all, depth):.2f} ms/step", flush=True)
    if depth == 2 and not os.environ.get("PROBE_SHORT"):
        for parts in (("load",), ("save",), ("load", "save"), ("load", "run")):
            print(f"   depth 2, only {'+'.join(parts)}: {run(lambda i: step(i, parts), sync_all, depth):.2f} ms/step", flush=True)
    for ctx, hb in lanes:
        ctx.set_async(False)
        ctx.free()
        ctx.destroy()
        for h in hb:
            h.free()
