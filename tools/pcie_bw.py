"""Profiling helper: pinned H2D / D2H bandwidth alone and in duplex (explains the e2e bound)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
dev = "cuda:0"
n = 720_000_000
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); d_in = torch.empty(n, dtype=torch.uint8, device=dev)
m = 500_000_000
h_out = torch.empty(m, dtype=torch.uint8).pin_memory(); d_out = torch.empty(m, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)


def d2h(size=m):
    with torch.cuda.stream(s2):
        h_out[:size].copy_(d_out[:size], non_blocking=True)
    torch.cuda.current_stream().wait_stream(s2)


def both(size=m):
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out[:size].copy_(d_out[:size], non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)


ms = timed(h2d); print(f"H2D 720 MB alone: {ms:.2f} ms = {n / ms / 1e6:.1f} GB/s")
ms = timed(d2h); print(f"D2H 500 MB alone: {ms:.2f} ms = {m / ms / 1e6:.1f} GB/s")
ms = timed(both); print(f"H2D 720 MB + D2H 500 MB duplex: {ms:.2f} ms (H2D-equivalent {n / ms / 1e6:.1f} GB/s)")
ms = timed(lambda: both(m // 2)); print(f"H2D 720 MB + D2H 250 MB duplex: {ms:.2f} ms (H2D-equivalent {n / ms / 1e6:.1f} GB/s)")
