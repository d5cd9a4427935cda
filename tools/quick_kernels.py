"""Kernel timing helper (python tools/quick_kernels.py [motif] [bp] [adjust] [delfi] [cleavage]): CUDA-event time of
the secondary kernels at chr1 scale (80 M fragments drawn on the GPU), median of 7 calls after 3 warm-ups."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from finaletoolkit_b200 import device as D
from finaletoolkit_b200.synth import synth_fragments_device, synth_twobit

what = set(sys.argv[1:]) or {"motif", "bp", "adjust"}
CLEN, N = 249_250_621, 80_000_000
dev = D.require_cuda("cuda:0")
st, sp, mq = synth_fragments_device(CLEN, N, 0, dev)
g = torch.Generator(device=dev); g.manual_seed(5)
sd = (torch.rand(N, generator=g, device=dev) < 0.5).to(torch.uint8)
fr = D.ContigFragments(st, sp, mq, sd, device=dev, max_len=600)


def timed(name, fn, nbytes):
    for _ in range(3):
        out = fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    print(f"{name}: {ms:.3f} ms (min {min(ts):.3f})  {nbytes / ms / 1e6:.0f} GB/s = {nbytes / ms / 1e6 / 6541.1:.1%} of measured HBM peak", flush=True)
    return out


if what & {"motif", "bp", "delfi"}:
    codes, nm = synth_twobit(CLEN, 0)
    ref = D.PackedContig.from_codes(codes, nm, device=dev)
    ws = np.arange(0, CLEN, 1_000_000, dtype=np.int64); we = np.minimum(ws + 1_000_000, CLEN)
if "motif" in what:
    o = timed("end_motifs k=4", lambda: D.end_motif_hist(fr, ref, ws, we, k=4, strand_mode=0, quality_threshold=30, pooled=True), 10 * N)
    print("  checksum", int(o.sum()))
    timed("end_motifs k=3 (runtime-K variant)", lambda: D.end_motif_hist(fr, ref, ws, we, k=3, strand_mode=0, quality_threshold=30, pooled=True), 10 * N)
if "bp" in what:
    o = timed("breakpoint_motifs k=6", lambda: D.end_motif_hist(fr, ref, ws, we, k=6, strand_mode=0, quality_threshold=30, pooled=True, breakpoint=True), 10 * N)
    print("  checksum", int(o.sum()))
if "delfi" in what:
    bs = np.arange(0, CLEN, 100_000, dtype=np.int64); be = np.minimum(bs + 100_000, CLEN)
    timed("delfi 100-kb bins", lambda: D.delfi_windows(fr, ref, bs, be, quality_threshold=30), 9 * N + CLEN * 3 // 8)
if "cleavage" in what:
    edges = np.arange(0, CLEN + 5000, 5000, dtype=np.int64).clip(max=CLEN)
    timed("cleavage", lambda: D.cleavage_intervals(fr, edges[:-1], edges[1:], CLEN, None, None, 30), 10 * N + 8 * CLEN)
if "hist" in what:
    edges = np.arange(0, CLEN + 5000, 5000, dtype=np.int64).clip(max=CLEN)
    timed("coverage counts (5-kb intervals)", lambda: D.interval_hist(fr, edges[:-1], edges[1:], "midpoint", None, None, 30), 9 * N)
    timed("coverage counts + pooled length histogram", lambda: D.interval_hist(fr, edges[:-1], edges[1:], "midpoint", None, None, 30, n_bins=601, pooled="hist"), 9 * N)
    timed("per-interval length histograms (frag_length_intervals)", lambda: D.interval_hist(fr, edges[:-1], edges[1:], "midpoint", None, None, 30, n_bins=601, first_seen=True), 9 * N)
if "adjust" in what:
    edges = np.arange(0, CLEN + 5000, 5000, dtype=np.int64).clip(max=CLEN)
    plan = D.WpsPlan(edges[:-1], edges[1:], CLEN, 180, dev)
    wps = plan.run(fr)
    ap = D.AdjustPlan(np.diff(plan.offsets), 1000, True, 21, 2, dev, skip_short=True)
    out = torch.empty(ap.n_total, dtype=torch.float64, device=dev)
    timed("adjust_wps fused (median 1000 + SG 21/2)", lambda: ap.run_rank(wps, 0, out), 4 * plan.n_positions + 8 * ap.n_total)
    print("  checksum", float(out.sum()), "flags", int(ap.run_rank(wps, 0, out)[1].sum()))
    ap0 = D.AdjustPlan(np.diff(plan.offsets), 1000, False, 21, 2, dev, skip_short=True)
    timed("adjust_wps median only", lambda: ap0.run_rank(wps, 0, out), 4 * plan.n_positions + 8 * ap.n_total)
