"""Genome-wide features across the GPUs of one box: one process per GPU, contigs sharded by LPT.

Launch with ``python -m torch.distributed.run --nproc-per-node N ...`` (NCCL).  Per-position /
per-interval outputs (WPS, adjust_wps, interval coverage, interval statistics) stay on the rank
that owns the contig - no collective.  Genome-wide reductions use exactly one packed
``all_reduce(SUM)`` (+ one ``MIN`` for first-seen order), see ``sharding.py``.

Reference semantics reproduced: ``coverage(normalize=True)`` total = ``single_coverage`` over the
whole file (frag/_coverage.py:215-227, 254); ``frag_length_bins`` genome-wide dict in stream order
(frag/_frag_length.py:408-421); ``end_motifs`` 4^k counts summed over contigs' 1 Mb windows
(frag/_motif_common.py:599-609).
"""
from __future__ import annotations

import numpy as np

from .sharding import DistContext, genome_length_dict, lpt_pack

__all__ = ["owned_contigs", "genome_total_coverage", "genome_length_distribution", "genome_end_motif_counts"]


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
                            ctx: DistContext | None = None, device=None) -> np.ndarray:
    """int64[4**k] genome-wide end-motif counts over every contig's 1 Mb windows, one all-reduce."""
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
                       k=k, strand_mode=strand_mode, quality_threshold=quality_threshold, pooled=True, counts=total)
    ctx.all_reduce_sum(total)
    return total[0].cpu().numpy()
