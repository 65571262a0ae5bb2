#!/usr/bin/env python
"""Copy / fill ceilings vs buffer size, back-to-back (steady state) -- what a perfect streaming kernel of
the same footprint as one stencil sweep can reach."""
import torch
for mb in (134, 268, 1074, 4295):
    n = mb * 1000 * 1000 // 8
    a = torch.empty(n, dtype=torch.float64, device="cuda").normal_()
    b = torch.empty_like(a)
    c = torch.empty_like(a)
    def run(f, reps=50):
        for _ in range(5): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): f()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / reps
    nb = n * 8
    flip = [0]
    def pingpong():
        if flip[0] & 1: a.copy_(b)
        else: b.copy_(a)
        flip[0] += 1
    t_copy = run(pingpong)
    t_fill = run(lambda: b.zero_())
    t_add = run(lambda: torch.add(a, b, out=c))
    t_memcpy = run(lambda: torch.cuda.current_stream().synchronize() or None, 1)
    print(f"{mb:5d} MB: copy(ping-pong) {2*nb/t_copy/1e9:7.0f} GB/s ({t_copy*1e6:7.1f} us)   fill {nb/t_fill/1e9:7.0f} GB/s   add(2r1w) {3*nb/t_add/1e9:7.0f} GB/s")
