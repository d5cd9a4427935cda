"""What a chunked duplex pipeline can reach on this box, without any of our kernels: H2D of 245 MB and D2H of
250 MB in n chunks on two streams, chunk k's D2H released by chunk k's H2D (+ an optional dummy kernel),
captured as a CUDA graph.  Reference point for finaletoolkit_b200/pipeline.py (tools/e2e_sweep.py)."""
import sys
import time

import torch

dev = torch.device("cuda:0")
H2D, D2H = 245_020_384, 249_654_241
h_in = torch.empty(H2D, dtype=torch.uint8).pin_memory()
h_out = torch.empty(D2H, dtype=torch.uint8).pin_memory()
d_in = [torch.empty(H2D, dtype=torch.uint8, device=dev) for _ in range(2)]
d_out = torch.empty(D2H, dtype=torch.uint8, device=dev)
d_tmp = torch.empty(D2H, dtype=torch.uint8, device=dev)


def build(n, compute, dep=True):
    s_in, s_c, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    cap = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    bi = [H2D * k // n for k in range(n + 1)]
    bo = [D2H * k // n for k in range(n + 1)]
    with torch.cuda.graph(g, stream=cap):
        cur = torch.cuda.current_stream()
        for s in (s_in, s_c, s_out):
            s.wait_stream(cur)
        for k in range(n):
            with torch.cuda.stream(s_in):
                d_in[0][bi[k]:bi[k + 1]].copy_(h_in[bi[k]:bi[k + 1]], non_blocking=True)
                e_in = torch.cuda.Event(); e_in.record(s_in)
            with torch.cuda.stream(s_c):
                if dep:
                    s_c.wait_event(e_in)
                if compute:
                    d_out[bo[k]:bo[k + 1]].copy_(d_tmp[bo[k]:bo[k + 1]])       # a ~60 us HBM pass as the "kernel"
                e_c = torch.cuda.Event(); e_c.record(s_c)
            with torch.cuda.stream(s_out):
                s_out.wait_event(e_c)
                h_out[bo[k]:bo[k + 1]].copy_(d_out[bo[k]:bo[k + 1]], non_blocking=True)
        cur.wait_stream(s_in); cur.wait_stream(s_c); cur.wait_stream(s_out)
    return g


def wall(g, reps=5):
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        g.replay(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


for n in [int(a) for a in sys.argv[1:]] or [1, 4, 8, 16, 32]:
    print(f"chunks {n:3d}: independent legs {wall(build(n, False, dep=False)):.2f} ms | D2H(k) after H2D(k) {wall(build(n, False)):.2f} ms"
          f" | + dummy kernel {wall(build(n, True)):.2f} ms", flush=True)
