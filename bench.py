#!/usr/bin/env python
"""Benchmark of the B200-native FinaleToolkit hot path (contract: see DESIGN.md §6).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm (oracle port)

Headline workload (BASELINE.json configs[1]): synthetic chr1-scale 30x fragment set - one
249,250,621-bp contig, 80 M start-sorted fragments per GPU (SURVEY.md §8d seeds) - and one
"step" = ONE fused sweep over the fragments: L-WPS (window 120, fragments 120-180, mapq >= 30)
over the contig tiled by 49,851 5-kb intervals + per-interval coverage counts + the pooled
fragment-length histogram.  N > 1: every rank owns its own chr1-scale shard (contig-sharded, no
data-path collective; the one real exchange is the all_reduce of the job-wide histogram) ->
weak scaling.

One JSON line on stdout (rank 0).  `value` = fragments/s with inputs resident in HBM; `e2e` =
the same step from pinned HOST buffers (packed fragment columns H2D + unpack + kernels + D2H of
every result); `roofline` = the fused kernel's algorithmic bytes / its own CUDA-event time vs the
measured HBM peak; `parity` = the step's outputs checked against the oracle on >= 2000 intervals
of the exact benchmark input; `cpu_baseline` = the oracle's OpenMP port of the reference loops on
a bounded sample of the same intervals; `genome` = BASELINE.json configs[2]/[3] (24 b37 contigs,
1e9 fragments, LPT-sharded over the ranks: STRONG scaling) through the product drivers
`distributed.multi_wps_genome` (WPS + coverage + histogram, and WPS -> adjust_wps without leaving
HBM) and the genome-wide reductions.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

CONTIG_LEN = 249_250_621
N_FRAG = int(os.environ.get("FTK_BENCH_NFRAG", 80_000_000))
GENOME_NFRAG = int(os.environ.get("FTK_BENCH_GENOME_NFRAG", 1_000_000_000))
IVL = 5000
WINDOW, MIN_LEN, MAX_LEN, MAPQ = 120, 120, 180, 30
METRIC = "wps_fragments_per_sec"
UNIT = "fragments/s"
WORKLOAD = ("synthetic chr1-scale 30x: 249,250,621 bp, %d fragments/GPU; L-WPS (W=120, 120-180, mapq>=30) "
            "over 49,851 5-kb intervals + per-interval coverage + length histogram" % N_FRAG)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_intervals():
    edges = np.arange(0, CONTIG_LEN + IVL, IVL, dtype=np.int64).clip(max=CONTIG_LEN)
    return edges[:-1].copy(), edges[1:].copy()


# --------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Polls NVML (SM clock, throttle reasons) while the timed region runs."""

    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int, period=0.002):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.mask, self.max_mhz = [], 0, None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._stop_evt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def sample_now(self):
        if self.ok:
            try:
                self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:  # noqa: BLE001
                pass

    def finish(self):
        self.sample_now()   # the GPU has just drained the timed region: never report zero samples
        self._stop_evt.set()
        if self.is_alive():
            self.join()
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        reasons = [n for b, n in self.REASONS.items() if self.mask & b and n != "gpu_idle"]
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(self.samples)}


# ----------------------------------------------------------- CPU baseline leg
def cpu_sample_rate(st, sp, mq, target_s: float, seed=0, keep=False, min_intervals=0):
    """Time the oracle's OpenMP port of the reference loops on a bounded interval sample.
    With ``keep`` also returns what it computed (``idx, wps, offsets, coverage``) so the caller can
    check the GPU step against it - the sample always holds the first and the last interval."""
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    fr = O.Frags(st, sp, mq)
    fr.max_len = min(fr.max_len, 600)
    s_all, e_all = make_intervals()
    rng = np.random.default_rng(seed)
    perm = rng.permutation(len(s_all))
    edge = np.array([0, len(s_all) - 1])
    perm = np.concatenate([edge, perm[~np.isin(perm, edge)]])

    def run(idx):
        t0 = time.perf_counter()
        out, off = O.wps_intervals(fr, s_all[idx], e_all[idx], CONTIG_LEN, WINDOW, MIN_LEN, MAX_LEN, MAPQ, threads=cores)
        cov = O.interval_coverage(fr, s_all[idx], e_all[idx], None, None, "midpoint", MAPQ, threads=cores)
        return time.perf_counter() - t0, out, off, cov

    probe = perm[: max(4 * cores, 64)]
    t_probe, out, off, _ = run(probe)
    n = int(min(len(perm), max(len(probe), min_intervals, len(probe) * target_s / max(t_probe, 1e-6))))
    idx = perm[:n]
    t, out, off, cov = run(idx)
    pos = int(off[-1])
    pos_per_s = pos / t
    frag_per_s = pos_per_s * (len(st) / CONTIG_LEN)
    info = {"value": frag_per_s, "unit": UNIT, "cores": cores, "kind": "port",
            "positions_per_sec": pos_per_s, "seconds": t,
            "sample": f"{n} random 5-kb intervals ({pos} positions) of the same workload: brute-force WPS "
                      f"(reference frag/_wps.py:176-188 loops) + per-interval coverage, OpenMP over intervals "
                      f"like the reference's Pool; fragments/s = positions/s x fragments/position of the workload"}
    return (info, idx, out, off, cov) if keep else info


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from finaletoolkit_b200.synth import synth_fragments
    t0 = time.time()
    st, sp, mq, _ = synth_fragments(CONTIG_LEN, N_FRAG, 0)
    log(f"[reference] synthesised {N_FRAG} fragments in {time.time() - t0:.1f}s")
    per_step = float(os.environ.get("FTK_BENCH_REF_STEP_S", 2.0))
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_sample_rate(st, sp, mq, per_step, seed=i)
        if i >= args.warmup:
            vals.append(r)
    v = float(np.mean([r["value"] for r in vals]))
    ms = float(np.mean([r["seconds"] for r in vals]) * 1e3)
    base = dict(vals[-1]); base["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "each step = a bounded random sample of the workload's intervals"},
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------- per-kernel device times (CUPTI)
def kernel_times(fn, reps=3):
    """Device time of every kernel ``fn`` launches (torch.profiler / CUPTI sees the ctypes launches too):
    {kernel name: mean microseconds per call of fn}.  Empty dict if the profiler is unavailable."""
    import torch
    try:
        from torch.profiler import ProfilerActivity, profile
        fn(); torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
        out = {}
        for ev in prof.key_averages():
            dt = getattr(ev, "device_time_total", None)
            if dt is None:
                dt = getattr(ev, "cuda_time_total", 0.0)
            if dt and "ftk::" in ev.key:
                out[ev.key.split("(")[0]] = out.get(ev.key.split("(")[0], 0.0) + float(dt) / reps
        return out
    except Exception as e:  # noqa: BLE001
        log(f"[bench] torch.profiler unavailable: {e!r}")
        return {}


# ------------------------------------------------- secondary kernels (N=1 only)
def other_kernels(frags, wps_i32, dev, peak):
    """The other hot-path kernels on the same chr1-scale shard, each against its own algorithmic bytes
    (DESIGN.md §4).  ``kernel_us`` = device time of the op's own kernels (CUPTI, per call); ``call_ms`` = CUDA
    events around the Python wrapper (planning + uploads included).  ``frac`` uses the kernel time."""
    import torch
    from finaletoolkit_b200 import device as D
    from finaletoolkit_b200.synth import synth_twobit

    def timed(fn, reps=3):
        fn(); torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        for i in range(reps):
            ev[i].record(); fn()
        ev[reps].record(); torch.cuda.synchronize()
        return min(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))

    def row(fn, nbytes, units, unit_name, main):
        call_ms = timed(fn)
        kt = kernel_times(fn)
        k_us = sum(v for k, v in kt.items() if any(m in k for m in main))
        ms = k_us * 1e-3 if k_us > 0 else call_ms
        gbs = nbytes / (ms * 1e-3) / 1e9
        return {"kernel_ms": ms if k_us > 0 else None, "call_ms": call_ms, "kernels_us": {k: round(v, 1) for k, v in kt.items()},
                "algorithmic_bytes": nbytes, "achieved_gbs": gbs, "frac_of_measured_hbm_peak": gbs / peak,
                "timing": "kernel (CUPTI)" if k_us > 0 else "call (CUDA events)", unit_name + "_per_sec": units / (ms * 1e-3)}

    out = {}
    ivl_s, ivl_e = make_intervals()
    # adjust_wps straight from the int32 WPS in HBM: median window 1000 + Savitzky-Golay(21,2) over the full 5-kb intervals
    full = int(np.sum((ivl_e - ivl_s) == IVL))
    lens = np.full(full, IVL, dtype=np.int64)
    x = wps_i32[: full * IVL]
    n_out = int((lens - 1000).sum())
    out["adjust_wps(median1000+savgol21/2, fused rank kernel)"] = row(
        lambda: D.adjust_segments(x, lens), 4 * x.numel() + 8 * n_out, n_out, "positions", ["adjust_rank_kernel"])
    xf = x.to(torch.float32)
    out["adjust_wps(round-1 path: histogram median + separate savgol)"] = row(
        lambda: D.adjust_segments(xf, lens, impl="hist"), 4 * x.numel() + 8 * n_out, n_out, "positions",
        ["adjust_hist_kernel", "adjust_savgol_kernel", "adjust_generic_kernel"])
    del x, xf
    # cleavage profile over the same 5-kb tiling (10 B/fragment + 8 B/position)
    out["cleavage_profile"] = row(lambda: D.cleavage_intervals(frags, ivl_s, ivl_e, CONTIG_LEN, None, None, MAPQ),
                                  10 * frags.n + 8 * CONTIG_LEN, CONTIG_LEN, "positions", ["cleavage"])
    # end motifs k=4, both strands, genome-wide 1-Mb windows pooled (10 B/fragment + L2-resident reference)
    codes, nm = synth_twobit(CONTIG_LEN, 0)
    ref = D.PackedContig.from_codes(codes, nm, device=dev)
    if frags.strand is None:
        frags.strand = torch.ones(frags.n, dtype=torch.uint8, device=dev)
    win = [(s, s + 1_000_000) for s in range(0, CONTIG_LEN - 1_000_000, 1_000_000)]
    win.append((CONTIG_LEN - CONTIG_LEN % 1_000_000, CONTIG_LEN))
    ws, we = [a for a, _ in win], [b for _, b in win]
    out["end_motifs(k=4,both strands)"] = row(
        lambda: D.end_motif_hist(frags, ref, ws, we, k=4, strand_mode=0, quality_threshold=MAPQ, pooled=True),
        10 * frags.n, frags.n, "fragments", ["end_motif_kernel"])
    out["breakpoint_motifs(k=6,both strands)"] = row(
        lambda: D.end_motif_hist(frags, ref, ws, we, k=6, strand_mode=0, quality_threshold=MAPQ, pooled=True, breakpoint=True),
        10 * frags.n, frags.n, "fragments", ["end_motif_kernel"])
    # DELFI: 100-kb bins, short/long counts + GC content (9 B/fragment + 0.375 B/base of packed reference)
    bs = np.arange(0, CONTIG_LEN - 100_000, 100_000, dtype=np.int64); be = bs + 100_000
    rng = np.random.default_rng(5)
    r0 = np.sort(rng.integers(0, CONTIG_LEN - 20_000, 400)); blk = (r0, r0 + rng.integers(200, 20_000, 400))
    gaps = ((121_500_000, 124_500_000), [(0, 10_000), (CONTIG_LEN - 10_000, CONTIG_LEN)])
    out["delfi_windows(100kb bins)"] = row(
        lambda: D.delfi_windows(frags, ref, bs, be, blacklist=blk, gaps=gaps, quality_threshold=MAPQ),
        9 * frags.n + (3 * CONTIG_LEN) // 8, frags.n, "fragments", ["delfi_count_kernel", "delfi_gc_kernel"])
    # stand-alone coverage (per-interval counts only) and the unpack kernel of the wire format
    cov_set = D.IntervalSet(ivl_s.tolist(), ivl_e.tolist(), dev)
    cov = torch.zeros(cov_set.n, dtype=torch.int64, device=dev)
    out["coverage(per-interval counts)"] = row(
        lambda: D.interval_hist(frags, intersect_policy="midpoint", quality_threshold=MAPQ, ivl_set=cov_set, out=(cov, None, None)),
        9 * frags.n, frags.n, "fragments", ["interval_count_warp_kernel"])
    return out


# ------------------------------------------------- public API wall clock (N=1 only)
def api_wall():
    """Wall clock (second call) of the reference-facing Python API on a 5 M-fragment, 60-Mb contig with
    12,000 5-kb intervals: where the host-side tail (result download, statistics, text / bigWig writers)
    stands next to the kernels.  Fragments come from an in-memory FragmentTable (decode excluded)."""
    import tempfile
    import finaletoolkit_b200 as F
    from finaletoolkit_b200.frag import _multi_wps as MW
    from finaletoolkit_b200.io.fragments import FragmentTable
    from finaletoolkit_b200.synth import synth_fragments
    clen, n = 60_000_000, 5_000_000
    table = FragmentTable({"1": synth_fragments(clen, n, 1)})
    tmp = tempfile.mkdtemp(prefix="ftk_wall_")
    tiles = os.path.join(tmp, "tiles.bed")
    with open(tiles, "w") as fh:
        fh.write("".join(f"1\t{a}\t{a + 5000}\n" for a in range(0, clen, 5000)))
    sites = os.path.join(tmp, "sites.bed")
    with open(sites, "w") as fh:
        fh.write("".join(f"1\t{a}\t{a + 1}\t.\t0\t+\n" for a in range(2500, clen, 5000)))
    cs = os.path.join(tmp, "cs")
    open(cs, "w").write(f"1\t{clen}\n")

    def wall(fn):
        fn()
        t0 = time.perf_counter(); fn()
        return (time.perf_counter() - t0) * 1e3

    out = {"input": f"{n} fragments, {clen} bp, 12000 intervals (in-memory table: decode excluded)"}
    out["frag_length_intervals_ms"] = wall(lambda: F.frag_length_intervals(table, tiles, os.path.join(tmp, "fli.bed")))
    out["coverage_ms"] = wall(lambda: F.coverage(table, tiles, os.path.join(tmp, "cov.bed")))
    out["frag_length_bins_ms"] = wall(lambda: F.frag_length_bins(table, "1", 0, clen, bin_size=5))
    for ext in (".bed.gz", ".bw"):
        ms = wall(lambda: F.multi_wps(table, sites, chrom_sizes=cs, output_file=os.path.join(tmp, "wps" + ext)))
        t = dict(MW.LAST_TIMINGS)
        out["multi_wps->" + ext] = {"total_ms": ms, "compute_ms": t["compute"] * 1e3, "write_ms": t["write"] * 1e3,
                                    "total_over_write": ms / max(t["write"] * 1e3, 1e-9)}
    out["adjust_wps(.bw->.bw)_ms"] = wall(lambda: F.adjust_wps(os.path.join(tmp, "wps.bw"), sites, os.path.join(tmp, "adj.bw"), cs))
    return out


# ------------------------------------------------------------- genome block
class SynthGenomeTable:
    """FragmentTable stand-in for BASELINE.json configs[2]: 24 b37 contigs, fragments proportional to
    length, drawn and sorted ON the GPU of the rank that owns the contig (``synth_fragments_device``)."""

    def __init__(self, total: int, dev):
        from finaletoolkit_b200.synth import B37_CONTIGS
        self.sizes = list(B37_CONTIGS)
        glen = sum(n for _, n in self.sizes)
        self.counts = {c: int(round(total * n / glen)) for c, n in self.sizes}
        self.index = {c: i for i, (c, _) in enumerate(self.sizes)}
        self.dev = dev
        self._cache = {}

    @property
    def contigs(self):
        return [c for c, _ in self.sizes]

    def n_fragments(self, c):
        return self.counts.get(c, 0)

    shard_weight = n_fragments

    def device(self, c, device=None):
        if c not in self._cache:
            from finaletoolkit_b200.device import ContigFragments
            from finaletoolkit_b200.synth import synth_fragments_device
            st, sp, mq = synth_fragments_device(dict(self.sizes)[c], self.counts[c], self.index[c], self.dev)
            self._cache[c] = ContigFragments(st, sp, mq, None, device=self.dev, contig=c, max_len=600)
        return self._cache[c]


def genome_block(args, dev, rank, world, peak):
    """BASELINE.json configs[2] + [3] at this N (STRONG scaling: the job is fixed, the ranks share it)."""
    import torch
    import torch.distributed as dist
    from finaletoolkit_b200 import device as D
    from finaletoolkit_b200.distributed import multi_wps_genome, owned_contigs, tile_genome
    from finaletoolkit_b200.sharding import DistContext

    ctx = DistContext()
    table = SynthGenomeTable(GENOME_NFRAG, dev)
    sizes = table.sizes
    mine = owned_contigs(table, ctx)
    t0 = time.time()
    for c in mine:
        table.device(c)
    torch.cuda.synchronize()
    gen_s = time.time() - t0
    sites = {c: v for c, v in tile_genome(sizes, IVL).items()}
    my_frag = sum(table.counts[c] for c in mine)
    my_pos = sum(dict(sizes)[c] for c in mine)
    tot_pos = sum(n for _, n in sizes)
    tot_frag = sum(table.counts.values())
    plans = {}      # multi_wps_genome caches the rank's GenomeShard (merged tile table, buffers) in here
    adjust_kw = dict(median_window_size=1000, savgol=True, savgol_window_size=21, savgol_poly_deg=2)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world == 1:
            return [float(v)]
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(x.item()) for x in out]

    clock_log = {}

    def timed(fn, reps, name=None):
        for _ in range(2):
            fn()
        # (the NVML sampler is created BEFORE the barrier: its start-up time differs from rank to rank and
        # would otherwise shift e0 on some ranks, i.e. charge one rank's set-up to the others' first pass)
        sampler = ClockSampler(dev.index if dev.index is not None else 0, period=0.001) if name else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if sampler:
            sampler.start()
        barrier()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.current_stream().synchronize()
        if sampler:
            clock_log[name] = sampler.finish()
        ms = e0.elapsed_time(e1) / reps
        barrier()
        return ms

    reps = max(2, min(args.steps, 5))
    state = {}

    def pass_a():   # WPS + per-interval coverage + genome-wide length histogram (ONE packed all_reduce)
        state["a"] = None       # drop the previous pass's outputs first: the allocator reuses their blocks
        state["a"] = multi_wps_genome(table, sizes, sites, IVL, WINDOW, MIN_LEN, MAX_LEN, MAPQ, coverage=True,
                                      length_hist=True, ctx=ctx, device=dev, contigs=mine, plans=plans, n_bins=601,
                                      sync=False)     # the total stays on the device: nothing waits for the GPU

    def pass_b():   # WPS -> adjust_wps, device resident (no bigWig round trip), no collective
        state["b"] = None
        state["b"] = multi_wps_genome(table, sizes, sites, IVL, WINDOW, MIN_LEN, MAX_LEN, MAPQ, adjust=adjust_kw,
                                      ctx=ctx, device=dev, contigs=mine, plans=plans, keep_adjusted=False, sync=False)

    ms_a_local = timed(pass_a, reps, "a")
    ms_a = max_over_ranks(ms_a_local)
    ranks_a = gather_ranks(ms_a_local)
    state.pop("b", None)
    ms_b_local = timed(pass_b, max(2, reps // 2), "b")
    ms_b = max_over_ranks(ms_b_local)
    ranks_b = gather_ranks(ms_b_local)

    # ---- correctness of what was timed (outside the timed region)
    res_a, hist, total = state["a"]
    total = int(total.item())
    flagged = [f for r in (state.get("b") or ({},))[0].values() for f in (r.flags or [])]
    checks = {"ok": True}
    if flagged and bool(torch.stack(flagged).any().item()):     # the timed adjust pass must not have needed the fallback
        checks["ok"] = False; checks["adjust_rank_kernel_flagged_tiles"] = True
    local_hist = torch.zeros_like(hist)
    local_total = 0
    for c in mine:   # recompute this rank's share with the stand-alone kernels: the all_reduce must add up
        fr = table.device(c)
        cnt, h, _ = D.interval_hist(fr, sites[c][0], sites[c][1], "midpoint", None, None, MAPQ, n_bins=hist.numel(), pooled="hist")
        if not torch.equal(cnt, res_a[c].cov):
            checks["ok"] = False; checks.setdefault("cov_mismatch", []).append(c)
        local_hist += h[0]; local_total += int(cnt.sum())
    summed = local_hist.clone()
    tot_t = torch.tensor([local_total], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(summed); dist.all_reduce(tot_t)
    checks["hist_allreduce_equals_sum_of_ranks"] = bool(torch.equal(summed, hist))
    checks["coverage_total"] = int(tot_t.item())
    checks["ok"] = checks["ok"] and checks["hist_allreduce_equals_sum_of_ranks"] and int(tot_t.item()) == total
    # oracle spot check on the smallest contig this rank owns (fragments copied back to the host)
    from oracle import oracle as O
    small = min(mine, key=lambda c: table.counts[c]) if mine else None
    mism, n_chk = 0, 0
    if small is not None:
        fr = table.device(small)
        ofr = O.Frags(fr.start.cpu().numpy(), fr.stop.cpu().numpy(), fr.mapq.cpu().numpy())
        ofr.max_len = 600
        clen = dict(sizes)[small]
        rng = np.random.default_rng(rank)
        pick = np.unique(np.concatenate([[0, len(sites[small][0]) - 1], rng.integers(0, len(sites[small][0]), 30)]))
        s_, e_ = sites[small][0][pick], sites[small][1][pick]
        exp, off = O.wps_intervals(ofr, s_, e_, clen, WINDOW, MIN_LEN, MAX_LEN, MAPQ, threads=os.cpu_count() or 1)
        exp_cov = O.interval_coverage(ofr, s_, e_, None, None, "midpoint", MAPQ, threads=os.cpu_count() or 1)
        r = res_a[small]
        wps_h = torch.cat([r.wps[r.offsets[i]: r.offsets[i + 1]] for i in pick]).cpu().numpy().astype(np.int64)
        mism += int((wps_h != exp).sum()) + int((r.cov.cpu().numpy()[pick] != exp_cov).sum())
        n_chk = int(len(pick))
        # adjust_wps of four of those intervals vs numpy median + scipy savgol on the oracle's WPS
        rb = multi_wps_genome(table, sizes, sites, IVL, WINDOW, MIN_LEN, MAX_LEN, MAPQ, adjust=adjust_kw, ctx=ctx,
                              device=dev, contigs=[small], plans=plans, reduce=False)[0][small]
        worst = 0.0
        for j, i in enumerate(pick[:4]):
            if e_[j] - s_[j] < 1021:
                continue
            got = rb.adjusted[rb.adj_offsets[i]: rb.adj_offsets[i + 1]].cpu().numpy()
            ref = O.adjust_core(exp[off[j]: off[j + 1]].astype(np.float64), 1000, False, True, 21, 2)
            worst = max(worst, float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-9 / 1e-5))))
        checks["adjust_max_rel_err"] = worst
        if worst > 1e-5:
            mism += 1
    mm = torch.tensor([mism, n_chk], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(mm)
    checks["oracle_intervals_checked"] = int(mm[1].item())
    checks["oracle_mismatches"] = int(mm[0].item())
    checks["ok"] = checks["ok"] and int(mm[0].item()) == 0

    # ---- configs[3]: DELFI-style 5-Mb-bin coverage + genome-wide fragment-length distribution
    # (first-seen ordered dict of frag_length_bins) with the NCCL reductions
    bins5 = tile_genome(sizes, 5_000_000)
    nb_hist = 601
    from finaletoolkit_b200.distributed import genome_bin_counts

    def config4():      # ONE launch per rank over its contigs laid end to end + the three small collectives
        state["c4"] = genome_bin_counts(table, bins5, n_bins=nb_hist, quality_threshold=MAPQ, ctx=ctx, device=dev,
                                        contigs=mine, cache_key="bins5")

    ms_c4 = max_over_ranks(timed(config4, reps))
    bin_counts, ldict = state["c4"]
    checks["config4_bins_total"] = int(sum(int(v.sum()) for v in bin_counts.values()))
    checks["config4_lengths_total"] = int(sum(ldict.values()))
    checks["ok"] = checks["ok"] and checks["config4_bins_total"] == checks["config4_lengths_total"]

    bytes_a = 9 * tot_frag + 4 * tot_pos
    n_adj_out = 0
    for c in mine:      # outputs of adjust_wps: n - 1000 per interval that is long enough for both filters
        no = np.maximum(sites[c][1] - sites[c][0], 0) - 1000
        n_adj_out += int(no[no >= 21].sum())
    adj_out_t = torch.tensor([n_adj_out], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(adj_out_t)
    bytes_b = bytes_a + 4 * tot_pos + 8 * int(adj_out_t.item())
    return {
        "config": f"24 b37 contigs ({tot_pos} bp), {tot_frag} synthetic fragments proportional to length (drawn on the GPUs), "
                  f"5-kb tiling = {sum(len(v[0]) for v in sites.values())} intervals; contigs LPT-sharded over {world} rank(s)",
        "scaling": "strong", "n_gpus": world, "reps": reps, "generate_s": gen_s,
        "wps_cov_hist": {"ms": ms_a, "fragments_per_sec": tot_frag / (ms_a * 1e-3), "positions_per_sec": tot_pos / (ms_a * 1e-3),
                         "algorithmic_bytes": bytes_a, "achieved_gbs_per_gpu": bytes_a / world / (ms_a * 1e-3) / 1e9,
                         "frac_of_hbm_peak_per_gpu": bytes_a / world / (ms_a * 1e-3) / 1e9 / peak,
                         "ms_per_rank": ranks_a, "imbalance_max_over_mean": max(ranks_a) / (sum(ranks_a) / len(ranks_a)),
                         "launches_per_pass_this_rank": "per shard group: 1 range prepass + 1 fused persistent kernel; + 1 sum + 1 all_reduce",
                         "sm_mhz_per_rank": gather_ranks((clock_log.get("a") or {}).get("sm_mhz") or 0.0),
                         "collective": "one all_reduce(SUM) of [coverage total, 601-bin histogram]"},
        "wps_then_adjust": {"ms": ms_b, "fragments_per_sec": tot_frag / (ms_b * 1e-3), "positions_per_sec": tot_pos / (ms_b * 1e-3),
                            "algorithmic_bytes": bytes_b, "achieved_gbs_per_gpu": bytes_b / world / (ms_b * 1e-3) / 1e9,
                            "frac_of_hbm_peak_per_gpu": bytes_b / world / (ms_b * 1e-3) / 1e9 / peak,
                            "ms_per_rank": ranks_b, "imbalance_max_over_mean": max(ranks_b) / (sum(ranks_b) / len(ranks_b)),
                            "collective": "none", "note": "int32 WPS stays in HBM; adjust = fused rank-median + Savitzky-Golay kernel",
                            "sm_mhz_per_rank": gather_ranks((clock_log.get("b") or {}).get("sm_mhz") or 0.0),
                            "clock_reasons_this_rank": (clock_log.get("b") or {}).get("reasons")},
        "coverage5mb_plus_length_bins": {"ms": ms_c4, "bins": int(sum(len(v[0]) for v in bins5.values())), "fragments_per_sec": tot_frag / (ms_c4 * 1e-3),
                                         "collective": "all_reduce(SUM) of 5-Mb bin counts + SUM/MIN of the length histogram / first-seen keys"},
        "lpt": {"fragments_this_rank": my_frag, "positions_this_rank": my_pos, "contigs_this_rank": len(mine)},
        "checks": checks,
    }


# ------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from finaletoolkit_b200 import device as D
    from finaletoolkit_b200.packed import PackedFragments
    from finaletoolkit_b200.synth import synth_fragments

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"   # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=dev)
    D.require_cuda(dev)
    numa = "unbound"
    if world > 1 and os.environ.get("FTK_BENCH_BIND", "1") == "1":
        # one process per GPU: keep this rank (and the host columns it is about to allocate and pin)
        # on the CPUs / memory node next to its GPU, so N ranks do not pull their H2D traffic
        # through one socket.  N=1 stays unbound: its cpu_baseline leg uses every host core.
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
            numa = f"{len(os.sched_getaffinity(0))} cpus near gpu {local}"
        except Exception as e:  # noqa: BLE001 - placement is an optimisation, never a failure
            numa = f"unbound ({type(e).__name__})"

    peaks_path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"

    t0 = time.time()
    st, sp, mq, sd = synth_fragments(CONTIG_LEN, N_FRAG, rank)
    log(f"[rank {rank}] synthesised {N_FRAG} fragments in {time.time() - t0:.1f}s")
    # the form the host decoder hands over: packed columns in pinned memory (4.06 B per fragment)
    t0 = time.time()
    packed = PackedFragments(st, sp, mq, None, max_len=600)
    pack_s = time.time() - t0
    frags = packed.to_device(dev)
    assert torch.equal(frags.start.cpu(), torch.from_numpy(st)) and torch.equal(frags.stop.cpu(), torch.from_numpy(sp))

    ivl_s, ivl_e = make_intervals()
    plan = D.WpsPlan(ivl_s, ivl_e, CONTIG_LEN, MAX_LEN, dev)
    n_bins = frags.max_len + 1
    wps_out = torch.empty(plan.n_positions, dtype=torch.int32, device=dev)
    cov_out = torch.zeros(plan.n_intervals, dtype=torch.int64, device=dev)
    hist_out = torch.zeros(n_bins, dtype=torch.int64, device=dev)
    job_hist = torch.zeros_like(hist_out)
    launches_per_step = 2   # fragment-range prepass + the fused kernel

    def step(ev=None):
        if ev is not None:
            ev[0].record()
        plan.ranges_fused(frags, WINDOW, zero_counts=cov_out, zero_hist=hist_out)   # also clears the accumulators
        if ev is not None:
            ev[1].record()
        # ONE sweep over the fragments: WPS + per-interval coverage + pooled length histogram
        plan.run_fused(frags, WINDOW, MIN_LEN, MAX_LEN, MAPQ, None, None, MAPQ, n_bins=n_bins, out=wps_out,
                       counts=cov_out, hist=hist_out, ranges_ready=True)
        if ev is not None:
            ev[2].record()
        if world > 1:
            # the one real exchange of the path: the job-wide length histogram (4.8 KB) over NCCL, issued
            # asynchronously so that its ~30 us of latency overlap the next step's kernels; the next
            # step waits for it before it reuses job_hist, the timed region ends with a wait
            if pending[0] is not None:
                pending[0].wait()
            job_hist.copy_(hist_out)
            pending[0] = dist.all_reduce(job_hist, op=dist.ReduceOp.SUM, async_op=True)

    pending = [None]

    def barrier():
        if world > 1:
            if pending[0] is not None:
                pending[0].wait(); pending[0] = None
            dist.barrier()
        torch.cuda.synchronize()

    W_ = max(args.warmup, 3)
    for _ in range(W_):
        step()
    barrier()

    # ---- device-resident timed region
    K = args.steps
    kev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    e0.record()
    for k in range(K):
        step(kev[k])
    if pending[0] is not None:       # the last step's all_reduce belongs to the timed region
        pending[0].wait(); pending[0] = None
    e1.record()
    torch.cuda.current_stream().synchronize()
    clocks = sampler.finish()
    barrier()
    ms_total = e0.elapsed_time(e1)
    t_ms = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_total = float(t_ms.item())
    ms_step = ms_total / K
    wps_ms = float(np.mean([kev[k][1].elapsed_time(kev[k][2]) for k in range(K)]))
    rng_ms = float(np.mean([kev[k][0].elapsed_time(kev[k][1]) for k in range(K)]))
    value = world * N_FRAG / (ms_step * 1e-3)
    pos_per_s = world * plan.n_positions / (ms_step * 1e-3)

    # ---- the N>1 collective is checked, not just timed: job_hist == sum over ranks of the local histograms
    dist_parity = None
    if world > 1:
        gathered = [torch.zeros_like(hist_out) for _ in range(world)]
        dist.all_gather(gathered, hist_out)
        ok = bool(torch.equal(torch.stack(gathered).sum(0), job_hist)) and int(hist_out.sum()) == int(cov_out.sum())
        dist_parity = {"job_hist_equals_sum_of_rank_hists": ok, "ranks": world,
                       "job_hist_total": int(job_hist.sum().item())}

    # ---- end-to-end: pinned PACKED host columns -> chunked H2D -> unpack + fused sweep -> D2H of every result
    from finaletoolkit_b200.pipeline import StreamedContig
    wire = os.environ.get("FTK_BENCH_WIRE", "int8")
    n_chunks = int(os.environ.get("FTK_BENCH_CHUNKS", 8))

    def make_pipe(w):
        return StreamedContig(None, None, None, ivl_s, ivl_e, CONTIG_LEN, WINDOW, MIN_LEN, MAX_LEN, MAPQ, max_frag_len=600,
                              n_chunks=n_chunks, device=dev, wps_dtype=w, packed=packed)
    # WPS crosses PCIe in the narrowest integer type that holds it exactly: int8 if the warm-up pass
    # raises no overflow flag (|WPS| <= 127 at this depth), else int16.  Chosen outside the timed region.
    pipe = make_pipe(wire)
    try:
        pipe.run()
    except OverflowError:
        del pipe
        wire = "int16"
        pipe = make_pipe(wire)
    h2d, d2h = pipe.h2d_bytes, pipe.d2h_bytes
    E = max(3, min(K, 10))
    for _ in range(2):
        pipe.run()
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(E):
        pipe.run()          # returns after the results are in pinned host memory
    g1.record()
    barrier()
    t_e = torch.tensor([g0.elapsed_time(g1) / E], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_ms = float(t_e.item())
    e2e_value = world * N_FRAG / (e2e_ms * 1e-3)
    checksum = int(pipe.h_wps.sum(dtype=torch.int64)) if rank == 0 else 0
    assert torch.equal(pipe.h_wps[: plan.n_positions].to(torch.int32), wps_out.cpu()) and \
        torch.equal(pipe.h_cov[: plan.n_intervals], cov_out.cpu()) and torch.equal(pipe.h_hist[0], hist_out.cpu()), \
        "e2e pipeline disagrees with the resident path"
    e2e_kernel_launches = pipe.kernel_launches
    n_e2e_chunks = len(pipe.chunks)
    del pipe
    torch.cuda.empty_cache()

    # ---- parity of the timed step vs the oracle on the exact benchmark input (+ the cpu_baseline timing)
    cpu_info, parity = None, None
    if rank == 0 and not args.no_cpu:
        target = float(os.environ.get("FTK_BENCH_CPU_S", 20.0)) if world == 1 else 3.0
        cpu_info, idx, o_wps, o_off, o_cov = cpu_sample_rate(st, sp, mq, target, keep=True,
                                                             min_intervals=2000 if world == 1 else 256)
        offs = plan.offsets
        sel = torch.from_numpy(np.concatenate([np.arange(offs[i], offs[i + 1]) for i in idx])).to(dev)
        got = wps_out[sel].cpu().numpy().astype(np.int64)
        mism = int((got != o_wps).sum()) + int((cov_out.cpu().numpy()[idx] != o_cov).sum())
        from oracle import oracle as O
        ofr = O.Frags(st, sp, mq); ofr.max_len = 600
        hist_exp = O.length_dist(ofr, 0, CONTIG_LEN, None, None, "midpoint", MAPQ)   # the tiling covers [0, contig)
        hh = hist_out.cpu().numpy()
        hist_ok = {int(i): int(hh[i]) for i in np.flatnonzero(hh)} == hist_exp
        parity = {"intervals": int(len(idx)), "positions": int(o_off[-1]), "mismatches": mism + (0 if hist_ok else 1),
                  "includes_first_and_last_interval": True, "length_histogram_equals_oracle": bool(hist_ok),
                  "against": "oracle port of the reference loops (brute force per position), same input arrays"}
        if world > 1:
            cpu_info = None   # the baseline number is an N=1 figure; the sample above was only the checker
    wps_for_extras = wps_out if (world == 1 and not args.no_extras) else None
    if wps_for_extras is None:
        del wps_out

    genome = None
    if not args.no_genome:
        try:
            genome = genome_block(args, dev, rank, world, peak)
        except Exception as e:  # noqa: BLE001 - never sink the headline line
            import traceback
            log(traceback.format_exc())
            genome = {"error": repr(e)}

    if rank == 0:
        algo_bytes = 9 * N_FRAG + 4 * plan.n_positions
        achieved = algo_bytes / (wps_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        prof = os.path.join(REPO, "profiles", "r2_wps_hex_fused_ncu.json")
        if os.path.exists(prof) and N_FRAG == 80_000_000:
            traffic = json.load(open(prof)).get("dram_bytes_per_launch")
            traffic_src = "profiles/r2_wps_hex_fused_ncu.json (one ncu --set full capture of this kernel on this workload)"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W_,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "fragments_per_gpu": N_FRAG, "positions_per_gpu": plan.n_positions,
                       "intervals_per_gpu": int(len(ivl_s)), "sharding": f"contig-per-rank x{world}",
                       "collective": ("none (N=1)" if world == 1 else
                                      "one NCCL all_reduce(SUM) of the 601-bin job-wide length histogram per step"),
                       "l2": "no flush: per-step working set 1.7 GB >> 126 MB L2",
                       "step": "fragment-range prepass + ONE fused kernel (WPS + coverage + length histogram)"},
            "positions_per_sec": pos_per_s,
            "roofline": {"bound": "hbm", "kernel": "wps_hex_kernel<false,int,FUSE>", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": wps_ms,
                         "ranges_prepass_ms": rng_ms},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms, "steps": E, "chunks": n_e2e_chunks, "wps_dtype_on_the_wire": wire,
                    "fragment_wire_format": f"packed {packed.record_bytes + 0.0625} B/fragment ({packed.n_raw} escaped blocks; "
                                            "finaletoolkit_b200/packed.py), unpacked on the GPU",
                    "host_pack_seconds_outside_timed_region": pack_s, "host_placement": numa,
                    "gpu_launches_per_step": e2e_kernel_launches},
            "gpu_launches": launches_per_step * K, "clocks": clocks, "wps_checksum": checksum,
            "parity": parity, "dist_parity": dist_parity, "genome": genome,
        }
        if world == 1 and not args.no_extras:
            try:
                line["other_kernels"] = other_kernels(frags, wps_for_extras, dev, peak)
            except Exception as e:  # noqa: BLE001 - secondary numbers must never sink the headline line
                line["other_kernels"] = {"error": repr(e)}
            try:
                line["api_wall"] = api_wall()
            except Exception as e:  # noqa: BLE001
                line["api_wall"] = {"error": repr(e)}
        line["cpu_baseline"] = cpu_info
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _claim_stdout():
    """Keep the real stdout for the ONE JSON line: libraries (NCCL's version banner, torchrun
    notices) write to fd 1 too, so fd 1 is pointed at stderr and the line goes to a saved dup."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


def main():
    global print
    real_stdout = _claim_stdout()
    _print = print

    def print(*a, **k):  # noqa: A001 - rank 0's JSON line -> the real stdout
        k.setdefault("file", real_stdout)
        _print(*a, **k)
        real_stdout.flush()

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / parity leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary kernels' timings")
    ap.add_argument("--no-genome", action="store_true", help="skip the genome-scale (configs 3/4) block")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
