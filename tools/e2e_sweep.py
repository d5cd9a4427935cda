"""End-to-end pipeline timing helper: packed host columns -> H2D -> unpack + fused sweep -> D2H, for several
chunk counts, plus the bare PCIe legs of the same byte counts (python tools/e2e_sweep.py [chunks ...])."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from finaletoolkit_b200 import device as D
from finaletoolkit_b200.packed import PackedFragments
from finaletoolkit_b200.pipeline import StreamedContig
from finaletoolkit_b200.synth import synth_fragments_device

CLEN, N = 249_250_621, 80_000_000
dev = D.require_cuda("cuda:0")
st, sp, mq = synth_fragments_device(CLEN, N, 0, dev)
packed = PackedFragments(st.cpu().numpy(), sp.cpu().numpy(), mq.cpu().numpy(), None, max_len=600)
del st, sp, mq
edges = np.arange(0, CLEN + 5000, 5000, dtype=np.int64).clip(max=CLEN)
print(f"record_bytes {packed.record_bytes}, wire {packed.wire_bytes() / 1e6:.1f} MB", flush=True)


def wall(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


# bare PCIe legs with the same bytes
hw = packed.words
dw = torch.empty_like(hw, device=dev)
ho = torch.empty(CLEN, dtype=torch.int8).pin_memory()
do = torch.empty(CLEN, dtype=torch.int8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
print(f"H2D alone {wall(lambda: dw.copy_(hw, non_blocking=True)):.2f} ms; D2H alone {wall(lambda: ho.copy_(do, non_blocking=True)):.2f} ms", flush=True)


def both():
    with torch.cuda.stream(s1):
        dw.copy_(hw, non_blocking=True)
    with torch.cuda.stream(s2):
        ho.copy_(do, non_blocking=True)


print(f"H2D || D2H {wall(both):.2f} ms", flush=True)
for n_chunks in [int(a) for a in sys.argv[1:]] or [8, 16, 32, 64]:
    pipe = StreamedContig(None, None, None, edges[:-1], edges[1:], CLEN, 120, 120, 180, 30, max_frag_len=600,
                          n_chunks=n_chunks, device=dev, wps_dtype="int8", packed=packed)
    print(f"chunks {n_chunks}: {wall(pipe.run):.2f} ms/step", flush=True)
    del pipe
