"""Does this box overlap host->device and device->host copies?  Times 512 MB H2D alone, 256 MB D2H alone and both at once
on two streams (pinned host memory), with CUDA events.  Explains the e2e leg of bench.py: if the two directions do not
overlap, pipelining steps over two contexts can only hide the sweeps, not the return copy."""
import torch

dev = torch.device("cuda:0")
n_in, n_out = 512 << 20, 256 << 20
h_in = torch.empty(n_in, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(n_out, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(n_in, dtype=torch.uint8, device=dev)
d_out = torch.empty(n_out, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    s1.synchronize(); s2.synchronize()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d(); d2h()


def wall(fn, reps=10):
    import time
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


t_in, t_out, t_both = wall(h2d), wall(d2h), wall(both)
print(f"H2D 512 MB alone: {t_in:.2f} ms = {n_in / t_in / 1e6:.1f} GB/s")
print(f"D2H 256 MB alone: {t_out:.2f} ms = {n_out / t_out / 1e6:.1f} GB/s")
print(f"both, two streams: {t_both:.2f} ms (sum {t_in + t_out:.2f}, max {max(t_in, t_out):.2f}) -> overlap {(t_in + t_out - t_both) / min(t_in, t_out):.2f} of the shorter copy")
