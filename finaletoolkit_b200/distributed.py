"""Genome-wide features across the GPUs of one box: one process per GPU, contigs sharded by LPT.

Launch with ``python -m torch.distributed.run --nproc-per-node N ...`` (NCCL).  Per-position /
per-interval outputs (WPS, adjust_wps, interval coverage, interval statistics) stay on the rank
that owns the contig - no collective.  Genome-wide reductions use exactly one packed
``all_reduce(SUM)`` (+ one ``MIN`` for first-seen order), see ``sharding.py``.

Reference semantics reproduced: ``coverage(normalize=True)`` total = ``single_coverage`` over the
whole file (frag/_coverage.py:215-227, 254); ``frag_length_bins`` genome-wide dict in stream order
(frag/_frag_length.py:408-421); ``end_motifs`` 4^k counts summed over contigs' 1 Mb windows
(frag/_motif_common.py:599-609); DELFI bins are independent per contig (frag/_delfi.py:283-294).
"""
from __future__ import annotations

import numpy as np

from .sharding import DistContext, genome_length_dict, lpt_pack

__all__ = ["owned_contigs", "genome_total_coverage", "genome_length_distribution", "genome_end_motif_counts",
           "genome_delfi_windows", "ContigWps", "multi_wps_genome", "adjust_wps_genome", "tile_genome",
           "gather_to_writer"]


def owned_contigs(table, ctx: DistContext | None = None):
    """Contigs of ``table`` this rank owns (LPT by ``shard_weight``: fragment count, or compressed byte
    span for a lazy tabix-backed table; deterministic and cheap on every rank)."""
    ctx = ctx or DistContext()
    weight = getattr(table, "shard_weight", None) or table.n_fragments
    weights = {c: weight(c) for c in table.contigs}
    return lpt_pack(weights, ctx.world)[ctx.rank]


def genome_total_coverage(table, min_length=None, max_length=None, intersect_policy="midpoint",
                          quality_threshold=30, ctx: DistContext | None = None, device=None) -> int:
    """Whole-file fragment count (the ``normalize=True`` denominator) with one all-reduce."""
    from .device import interval_hist, require_cuda, torch
    ctx = ctx or DistContext()
    dev = require_cuda(device)
    t = torch()
    total = t.zeros(1, dtype=t.int64, device=dev)
    for c in owned_contigs(table, ctx):
        if table.n_fragments(c):
            cnt, _, _ = interval_hist(table.device(c, dev), [0], [None], intersect_policy, min_length, max_length,
                                      quality_threshold)
            total += cnt[:1]
    ctx.all_reduce_sum(total)
    return int(total.item())


def genome_length_distribution(table, min_length=0, max_length=None, intersect_policy="midpoint",
                               quality_threshold=30, ctx: DistContext | None = None, device=None,
                               region=(None, None)) -> dict:
    """The reference's genome-wide ``length -> count`` dict (first-seen order) on every rank."""
    from .device import interval_hist, require_cuda, torch
    ctx = ctx or DistContext()
    dev = require_cuda(device)
    t = torch()
    order = {c: i for i, c in enumerate(table.contigs)}
    mine = owned_contigs(table, ctx)
    # a common histogram width: the longest fragment over ALL contigs = MAX over the ranks' own contigs
    gmax_t = t.tensor([max([table.device(c, dev).max_len for c in mine if table.n_fragments(c)] + [0])],
                      dtype=t.int64, device=dev)
    ctx.all_reduce_max(gmax_t)
    gmax = int(gmax_t.item())
    n_bins = (gmax if max_length is None else min(gmax, int(max_length))) + 1
    parts = []
    for c in mine:
        if not table.n_fragments(c):
            continue
        _, h, f = interval_hist(table.device(c, dev), [region[0]], [region[1]], intersect_policy, min_length,
                                max_length, quality_threshold, n_bins=n_bins, pooled=True, first_seen=True)
        parts.append((order[c], h[0], f[0]))
    if not parts:  # rank without fragments still takes part in the collectives
        parts = [(0, t.zeros(n_bins, dtype=t.int64, device=dev), t.full((n_bins,), 2 ** 31 - 1, dtype=t.int32, device=dev))]
    return genome_length_dict(ctx, parts, n_bins)


def genome_end_motif_counts(table, ref, k=4, strand_mode=0, quality_threshold=30,
                            ctx: DistContext | None = None, device=None, breakpoint=False) -> np.ndarray:
    """int64[4**k] genome-wide end-motif (or, with ``breakpoint``, breakpoint-motif) counts over every
    contig's 1 Mb windows, one all-reduce."""
    from .device import end_motif_hist, require_cuda, torch
    from .frag._motif_common import genome_windows
    ctx = ctx or DistContext()
    dev = require_cuda(device)
    t = torch()
    total = t.zeros((1, 4 ** k), dtype=t.int64, device=dev)
    mine = set(owned_contigs(table, ctx))
    for chrom, chrom_length in ref.chroms.items():
        if chrom not in mine or not table.n_fragments(chrom):
            continue
        w = genome_windows(chrom_length)
        end_motif_hist(table.device(chrom, dev), ref.device_contig(chrom, dev), [a for a, _ in w], [b for _, b in w],
                       k=k, strand_mode=strand_mode, quality_threshold=quality_threshold, pooled=True, counts=total,
                       breakpoint=breakpoint)
    ctx.all_reduce_sum(total)
    return total[0].cpu().numpy()


def genome_delfi_windows(table, ref, bins_by_contig, blacklist_by_contig=None, gaps_by_contig=None,
                         quality_threshold=30, ctx: DistContext | None = None, device=None) -> dict:
    """DELFI bin counts of a whole genome: contigs are LPT-sharded over the ranks (bins of different
    contigs are independent, frag/_delfi.py:283-294 hands them to a process pool), each rank runs
    ``ftk_delfi_windows_u64`` on the contigs it owns and ONE all-reduce merges the packed table.

    ``bins_by_contig``: {contig: (starts, stops)}; ``blacklist_by_contig``: {contig: (starts, stops)}
    sorted by (start, stop); ``gaps_by_contig``: {contig: (centromere, telomeres)}.
    Returns {contig: int64[n_bins, 4]} = short, long, num_frags, G+C bases, identical on every rank."""
    from .device import delfi_windows, require_cuda, torch
    ctx = ctx or DistContext()
    dev = require_cuda(device)
    t = torch()
    names = [c for c in bins_by_contig if len(bins_by_contig[c][0])]
    sizes = [len(bins_by_contig[c][0]) for c in names]
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    packed = t.zeros((int(offs[-1]), 4), dtype=t.int64, device=dev)
    mine = set(owned_contigs(table, ctx))
    for i, chrom in enumerate(names):
        # a contig the fragment table does not know is owned by nobody: rank 0 counts its G+C bases
        owner_here = chrom in mine or (chrom not in table.contigs and ctx.rank == 0)
        if not owner_here:
            continue
        ws, we = bins_by_contig[chrom]
        got = delfi_windows(table.device(chrom, dev), ref.device_contig(chrom, dev) if chrom in ref.chroms else None,
                            np.asarray(ws, np.int64), np.asarray(we, np.int64),
                            blacklist=(blacklist_by_contig or {}).get(chrom), gaps=(gaps_by_contig or {}).get(chrom),
                            quality_threshold=quality_threshold)
        packed[int(offs[i]): int(offs[i + 1])] = got
    ctx.all_reduce_sum(packed)
    host = packed.cpu().numpy()
    return {c: host[int(offs[i]): int(offs[i + 1])] for i, c in enumerate(names)}


# ----------------------------------------------------------------------------------------------
# Genome-wide WPS (+ coverage + length histogram) and the device-resident WPS -> adjust_wps chain
# ----------------------------------------------------------------------------------------------
class ContigWps:
    """Per-contig result of ``multi_wps_genome`` on the rank that owns the contig (device tensors)."""

    __slots__ = ("contig", "starts", "stops", "offsets", "wps", "cov", "adjusted", "adj_offsets", "adj_segments")

    def __init__(self, contig, starts, stops, offsets, wps, cov=None):
        self.contig, self.starts, self.stops, self.offsets = contig, starts, stops, offsets
        self.wps, self.cov = wps, cov
        self.adjusted = self.adj_offsets = self.adj_segments = None


def tile_genome(chrom_sizes, interval_size: int = 5000) -> dict:
    """{contig: (starts, stops)}: every contig tiled by ``interval_size`` windows (the genome-wide
    site set of SURVEY.md §8d config 3)."""
    out = {}
    for contig, size in chrom_sizes:
        edges = np.arange(0, int(size) + int(interval_size), int(interval_size), dtype=np.int64).clip(max=int(size))
        out[contig] = (edges[:-1].copy(), edges[1:].copy())
    return out


def multi_wps_genome(table, chrom_sizes, sites: dict | None = None, interval_size=5000, window_size=120,
                     min_length=120, max_length=180, quality_threshold=30, coverage=False, length_hist=False,
                     adjust: dict | None = None, ctx: DistContext | None = None, device=None,
                     contigs: list | None = None, reduce: bool = True, plans: dict | None = None,
                     keep_adjusted: bool = True, n_bins: int | None = None):
    """Genome-wide L-WPS over the ranks of one box (reference drivers frag/_multi_wps.py:152-198 +,
    with ``adjust``, frag/_adjust_wps.py:229-291) - contigs LPT-sharded, every rank sweeps the contigs
    it owns, results stay on the owning rank's GPU, no data-path collective.

    ``sites``: {contig: (starts, stops)} sorted by start (``tile_genome`` when None).  With
    ``coverage`` / ``length_hist`` the sweep is the fused pass (WPS + per-interval midpoint coverage +
    pooled length histogram, one read of the fragments); the histogram and the coverage total are the
    only things combined across ranks (ONE packed all_reduce).  ``adjust`` = kwargs of
    ``device.adjust_segments`` (median_window_size, savgol, ...): each interval of at least
    ``median_window_size`` positions is adjusted straight from the int32 WPS in HBM - no bigWig
    round trip, no float32 copy.

    ``n_bins``: histogram width when the caller knows it (longest fragment of the job + 1) - saves the
    MAX all-reduce and its host synchronisation; ``plans``: {contig: WpsPlan} to reuse across calls
    (built and added when missing);
    ``contigs``: override the LPT assignment; ``reduce=False`` skips the collectives (single-rank
    checks inside a multi-rank job); ``keep_adjusted=False`` drops each contig's adjusted series once
    computed (timing runs).

    Returns ``(results {contig: ContigWps} of this rank, hist int64[n_bins] | None, total | None)``.
    """
    from .device import AdjustPlan, WpsPlan, adjust_segments, require_cuda, torch
    ctx = ctx or DistContext()
    dev = require_cuda(device)
    t = torch()
    sizes = dict(chrom_sizes)
    if sites is None:
        sites = tile_genome(chrom_sizes, interval_size)
    mine = contigs if contigs is not None else [c for c in owned_contigs(table, ctx) if c in sites]
    fused = bool(coverage or length_hist)
    if not length_hist:
        n_bins = 0
    elif n_bins is None:   # one histogram width for the whole job: MAX over the ranks' own contigs
        m = t.tensor([max([table.device(c, dev).max_len for c in mine if table.n_fragments(c)] + [0])],
                     dtype=t.int64, device=dev)
        if reduce:
            ctx.all_reduce_max(m)
        n_bins = int(m.item()) + 1
    n_bins = int(n_bins)
    packed = t.zeros(1 + n_bins, dtype=t.int64, device=dev)      # [coverage total, histogram...]
    results = {}
    # (Alternating the contigs between two streams so that a contig's kernel tail overlaps the next
    # contig's head was measured and lost - 14.4 ms against 6.5 ms per genome pass: outputs allocated on
    # side streams defeat the caching allocator's reuse.  One stream it is.)
    deferred = []
    for n_c, c in enumerate(mine):
        starts, stops = (np.asarray(a, dtype=np.int64) for a in sites[c])
        plan = plans.get(c) if plans is not None else None
        if plan is None:
            plan = WpsPlan(starts, stops, int(sizes[c]), int(max_length), dev)
            if plans is not None:
                plans[c] = plan
        frags = table.device(c, dev)
        if fused:
            cov = t.empty(max(plan.n_intervals, 1), dtype=t.int64, device=dev)   # cleared by the range prepass
            wps, cov, _ = plan.run_fused(frags, window_size, min_length, max_length, quality_threshold,
                                         None, None, quality_threshold, n_bins=n_bins, counts=cov,
                                         hist=packed[1:] if n_bins else None, zero_counts=True)
        else:
            wps, cov = plan.run(frags, window_size, min_length, max_length, quality_threshold), None
        res = ContigWps(c, starts, stops, plan.offsets, wps, cov)
        if adjust is not None:
            # every interval is a segment of the int32 WPS buffer as it lies in HBM (no gather, no float
            # copy); intervals shorter than the filters need produce no output, like the reference's driver
            aplan = plans.get((c, "adjust")) if plans is not None else None
            if aplan is None:
                aplan = AdjustPlan(np.diff(plan.offsets), int(adjust.get("median_window_size", 1000)),
                                   adjust.get("savgol", True), adjust.get("savgol_window_size", 21),
                                   adjust.get("savgol_poly_deg", 2), dev, skip_short=True)
                if plans is not None:
                    plans[(c, "adjust")] = aplan
            res.adj_offsets = aplan.out_off
            res.adj_segments = np.flatnonzero(aplan.n_out > 0)
            if aplan.n_total:
                res.adjusted, _, flag = adjust_segments(wps, None, plan=aplan, defer_check=True, **adjust)
                if flag is not None:
                    deferred.append((res, aplan, flag.any()))
                if not keep_adjusted:
                    res.adjusted = None
        results[c] = res
    if deferred and bool(t.stack([f for _, _, f in deferred]).any().item()):
        # a tile the rank kernel could not take (cannot happen for integer WPS of ordinary depth):
        # redo those contigs on the general path
        for res, aplan, f in deferred:
            if bool(f.item()):
                res.adjusted, _ = adjust_segments(res.wps, None, plan=aplan, impl="hist", **adjust)
                if not keep_adjusted:
                    res.adjusted = None
    if fused:
        covs = [r.cov for r in results.values() if r.cov is not None]
        if covs:
            packed[0] = t.stack([c_.sum() for c_ in covs]).sum()
        if reduce:
            ctx.all_reduce_sum(packed)
    hist = packed[1:] if n_bins else None
    return results, hist, (int(packed[0].item()) if fused else None)


def adjust_wps_genome(results: dict, **adjust):
    """Adjust already computed device-resident WPS (``multi_wps_genome`` results) in place of the
    reference's bigWig -> adjust_wps -> bigWig pass (frag/_adjust_wps.py:59-111).  Interval i's adjusted
    series is ``res.adjusted[res.adj_offsets[i]: res.adj_offsets[i + 1]]`` (empty for intervals shorter
    than the filters need)."""
    from .device import AdjustPlan, adjust_segments
    for res in results.values():
        aplan = AdjustPlan(np.diff(res.offsets), int(adjust.get("median_window_size", 1000)),
                           adjust.get("savgol", True), adjust.get("savgol_window_size", 21),
                           adjust.get("savgol_poly_deg", 2), res.wps.device, skip_short=True)
        res.adj_offsets, res.adj_segments = aplan.out_off, np.flatnonzero(aplan.n_out > 0)
        if aplan.n_total:
            res.adjusted, _ = adjust_segments(res.wps, None, plan=aplan, **adjust)
    return results


def gather_to_writer(obj, ctx: DistContext | None = None):
    """Collect a picklable per-rank object on rank 0 (list in rank order; None elsewhere) - used by the
    file-writing API mirrors, whose single output file is written by rank 0 only."""
    ctx = ctx or DistContext()
    if not (ctx.on and ctx.world > 1):
        return [obj]
    out = [None] * ctx.world if ctx.rank == 0 else None
    ctx.dist.gather_object(obj, out, dst=0)
    return out
