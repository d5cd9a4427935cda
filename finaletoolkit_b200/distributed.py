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
           "genome_delfi_windows"]


def owned_contigs(table, ctx: DistContext | None = None):
    """Contigs of ``table`` this rank owns (LPT by fragment count, deterministic on every rank)."""
    ctx = ctx or DistContext()
    weights = {c: table.n_fragments(c) for c in table.contigs}
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
                               quality_threshold=30, ctx: DistContext | None = None, device=None) -> dict:
    """The reference's genome-wide ``length -> count`` dict (first-seen order) on every rank."""
    from .device import interval_hist, require_cuda, torch
    ctx = ctx or DistContext()
    dev = require_cuda(device)
    t = torch()
    order = {c: i for i, c in enumerate(table.contigs)}
    mine = owned_contigs(table, ctx)
    # a common histogram width: the longest admissible fragment over ALL contigs (host-side max)
    gmax = max([int((table.host(c)[1].astype(np.int64) - table.host(c)[0]).max()) for c in table.contigs
                if table.n_fragments(c)] + [0])
    n_bins = (gmax if max_length is None else min(gmax, int(max_length))) + 1
    parts = []
    for c in mine:
        if not table.n_fragments(c):
            continue
        _, h, f = interval_hist(table.device(c, dev), [None], [None], intersect_policy, min_length, max_length,
                                quality_threshold, n_bins=n_bins, pooled=True, first_seen=True)
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
