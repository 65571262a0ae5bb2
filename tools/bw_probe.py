#!/usr/bin/env python
"""Device-memory ceilings on this GPU (torch ops, CUDA events): copy, write-only, read-only."""
import torch
n = 1 << 30
a = torch.empty(n, dtype=torch.bfloat16, device="cuda").normal_()
b = torch.empty_like(a)
def t(f, reps=10):
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e-3
nb = n * 2
print("copy  GB/s", 2 * nb / t(lambda: b.copy_(a)) / 1e9)
print("write GB/s", nb / t(lambda: b.zero_()) / 1e9)
print("read  GB/s", nb / t(lambda: a.view(torch.int16).max()) / 1e9)
c = torch.empty_like(a)
print("add(2r1w) GB/s", 3 * nb / t(lambda: torch.add(a, b, out=c)) / 1e9)
