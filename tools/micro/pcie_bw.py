"""PCIe copy bandwidth of the box (pinned host memory): H2D alone, D2H alone, both at once. Context for the e2e number."""
import torch
n = 512 * 1024 * 1024
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n // 2, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n // 2, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        s1.synchronize(); s2.synchronize()
        b.record(); b.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d(); d2h()


t = timed(h2d); print(f"H2D alone   : {n / t / 1e6:.1f} GB/s")
t = timed(d2h); print(f"D2H alone   : {n / 2 / t / 1e6:.1f} GB/s")
t = timed(both); print(f"H2D + D2H/2 : {n / t / 1e6:.1f} GB/s in, {n / 2 / t / 1e6:.1f} GB/s out (the e2e traffic mix: 32 B in, 16 B out per ray)")
