"""Profiling helper (PYTHONPATH=. python tools/quick_motif.py [k] [breakpoint]): end-motif kernel at chr1 scale."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sys
import numpy as np, torch
from finaletoolkit_b200 import device as D
from finaletoolkit_b200.synth import synth_fragments, synth_twobit
clen = 249_250_621; n = 80_000_000
k = int(sys.argv[1]) if len(sys.argv) > 1 else 4
bp = len(sys.argv) > 2
st, sp, mq, sd = synth_fragments(clen, n, 0)
fr = D.ContigFragments(st, sp, mq, sd, device="cuda:0", max_len=600)
codes, nm = synth_twobit(clen, 0)
ref = D.PackedContig.from_codes(codes, nm, device="cuda:0")
win = [(s, s + 1_000_000) for s in range(0, clen - 1_000_000, 1_000_000)]
win.append((clen - clen % 1_000_000, clen))
ws, we = np.array([a for a, _ in win]), np.array([b for _, b in win])
f = lambda: D.end_motif_hist(fr, ref, ws, we, k=k, strand_mode=0, quality_threshold=30, pooled=True, breakpoint=bp)
for _ in range(3): f()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
for i in range(5):
    ev[i].record(); out = f()
ev[5].record(); torch.cuda.synchronize()
ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(5)]
print(f"k={k} breakpoint={bp} ms={np.median(ts):.3f} (min {min(ts):.3f}) GB/s={10 * n / np.median(ts) / 1e6:.1f} checksum={int(out.sum())}")
