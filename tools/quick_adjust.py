"""Profiling helper (run from the repo root: PYTHONPATH=. python tools/...): adjust_wps timing at chr1 scale (raw WPS from our kernel -> float32 -> adjust)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sys
import numpy as np, torch
from finaletoolkit_b200.device import ContigFragments, WpsPlan, adjust_segments
from finaletoolkit_b200.synth import synth_fragments
clen = int(sys.argv[2]) if len(sys.argv) > 2 else 249_250_621
n = int(sys.argv[1]) if len(sys.argv) > 1 else 80_000_000
st, sp, mq, sd = synth_fragments(clen, n, 0)
fr = ContigFragments(st, sp, mq, sd, device="cuda:0", max_len=600)
edges = np.arange(0, clen + 5000, 5000).clip(max=clen)
wps = WpsPlan(edges[:-1], edges[1:], clen, 180, "cuda:0").run(fr)
x = wps.to(torch.float32)
del wps
lens = np.diff(edges)[:-1]
x = x[: int(lens.sum())]
for mode, seg in (("5kb segments", lens), ("one merged segment", np.array([int(lens.sum())]))):
    import os
    kw = dict(savgol=os.environ.get("SG", "1") == "1", use_mean=os.environ.get("MEAN", "0") == "1")
    for _ in range(2):
        out, off = adjust_segments(x, seg, **kw)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for i in range(3):
        ev[i].record(); out, off = adjust_segments(x, seg, **kw)
    ev[3].record(); torch.cuda.synchronize()
    ms = min(ev[i].elapsed_time(ev[i + 1]) for i in range(3))
    npos = int(off[-1])
    print(f"{mode}: {ms:.2f} ms, {npos/ms/1e6:.2f} Gpos/s, {(4*x.numel()+8*npos)/ms/1e6:.0f} GB/s algorithmic; out[:3]={out[:3].tolist()}")
    del out
