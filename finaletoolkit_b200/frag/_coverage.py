"""Fragment coverage (per-interval fragment counts) - API mirror of reference frag/_coverage.py.

``single_coverage``'s generator loop (frag/_coverage.py:117-130) and ``coverage``'s Pool over
intervals (:244-248) become one ``ftk_interval_hist_u64`` launch per contig; the fp64
normalisation keeps the reference's operation order (scale_factor /= total, then cov * scale_factor).
"""
from __future__ import annotations

import gzip
import sys
import time
from typing import NamedTuple

from ..exceptions import InvalidInputError
from ..io.fragments import as_table
from ..utils import get_intervals
from ._common import dist_context, group_by_contig, is_writer, per_fetch

__all__ = ["coverage", "single_coverage", "CoverageResult"]


class CoverageResult(NamedTuple):
    contig: str | None
    start: int | None
    stop: int | None
    name: str
    coverage: float


def _count_intervals(table, intervals, min_length, max_length, intersect_policy, quality_threshold, ctx=None):
    """[(contig, start, stop)] -> list of int counts (input order); contig None = all contigs.

    With a multi-rank ``ctx`` every rank counts only on the contigs it owns (LPT sharding) and ONE
    all_reduce(SUM) of the count vector gives every rank the full answer (the reference's Pool over
    intervals, frag/_coverage.py:212-248, and its whole-file total, :215-227)."""
    from ..device import interval_hist, policy_code
    policy_code(intersect_policy)
    counts = [0] * len(intervals)
    mine = None
    if ctx is not None:
        from ..distributed import owned_contigs
        mine = set(owned_contigs(table, ctx))
    for contig, idx in group_by_contig([iv[0] for iv in intervals]).items():
        if contig is None:
            for i in idx:
                s, e = intervals[i][1], intervals[i][2]
                if not (s is None and e is None) and not (s == 0 and e is None):
                    raise InvalidInputError("contig should be specified if start or stop given.")
                tot = 0
                for c in table.contigs:
                    if mine is not None and c not in mine:
                        continue
                    cnt, _, _ = interval_hist(table.device(c), [s], [e], intersect_policy, min_length, max_length,
                                              quality_threshold)
                    tot += int(cnt[0])
                counts[i] = tot
            continue
        if mine is not None and contig not in mine:
            continue

        def run(tab, sel, contig=contig, idx=idx):
            if tab.n_fragments(contig) == 0:
                return [0] * len(sel)
            cnt, _, _ = interval_hist(tab.device(contig), [intervals[idx[k]][1] for k in sel],
                                      [intervals[idx[k]][2] for k in sel], intersect_policy, min_length, max_length,
                                      quality_threshold)
            return [int(v) for v in cnt.cpu().tolist()]

        # the region handed to the fetch is the interval itself (frag/_coverage.py:118-127)
        got = per_fetch(table, contig, [intervals[i][1] for i in idx], [intervals[i][2] for i in idx], run)
        for i, v in zip(idx, got):
            counts[i] = v
    if ctx is not None and counts:
        import torch
        from ..device import require_cuda
        buf = torch.tensor(counts, dtype=torch.int64, device=require_cuda())
        ctx.all_reduce_sum(buf)
        counts = [int(v) for v in buf.cpu().tolist()]
    return counts


def single_coverage(input_file, contig=None, start=0, stop=None, name=".", min_length=None, max_length=None,
                    intersect_policy="midpoint", quality_threshold=30, verbose=False, reference_file=None):
    """Fragment count over one region (frag/_coverage.py:53-137)."""
    if verbose:
        start_time = time.time()
    table = as_table(input_file, reference_file)
    cov = _count_intervals(table, [(contig, start, stop)], min_length, max_length, intersect_policy,
                           quality_threshold)[0]
    if verbose:
        sys.stderr.write(f"single_coverage took {time.time() - start_time} s to complete\n")
    return CoverageResult(contig, start, stop, "." if name is None else name, cov)


def coverage(input_file, interval_file, output_file, scale_factor=1.0, min_length=None, max_length=None,
             normalize=False, intersect_policy="midpoint", quality_threshold=30, workers=1, verbose=False,
             reference_file=None):
    """Scaled / normalised fragment counts of every BED interval (frag/_coverage.py:145-305)."""
    if verbose:
        start_time = time.time()
    table = as_table(input_file, reference_file)
    ctx = dist_context()   # torch.distributed initialised: contigs sharded over the ranks, rank 0 writes
    if normalize:
        total = _count_intervals(table, [(None, 0, None)], min_length, max_length, intersect_policy,
                                 quality_threshold, ctx)[0]
    intervals = get_intervals(interval_file)
    counts = _count_intervals(table, [iv[:3] for iv in intervals], min_length, max_length, intersect_policy,
                              quality_threshold, ctx)
    if not is_writer(ctx):
        output_file = None
    if normalize:
        scale_factor /= total  # frag/_coverage.py:254 (ZeroDivisionError for an empty file, like the reference)
    return_val = []
    if output_file is not None:
        output_is_file = False
        try:
            if output_file.endswith(".bed") or output_file.endswith(".bedgraph"):
                output_is_file = True
                output = open(output_file, "w")
            elif output_file.endswith(".bed.gz"):
                output = gzip.open(output_file, "wt")
                output_is_file = True
            elif output_file == "-":
                output = sys.stdout
            else:
                raise ValueError("output_file should have .bed or .bed.gz as suffix")
            bedgraph = output_file.endswith(".bedgraph")
            for (contig, start, stop, name), cov in zip(intervals, counts):
                if bedgraph:
                    output.write(f"{contig}\t{start}\t{stop}\t{cov * scale_factor}\n")
                else:
                    output.write(f"{contig}\t{start}\t{stop}\t{name}\t{cov * scale_factor}\n")
                return_val.append(CoverageResult(contig, start, stop, name, cov * scale_factor))
        finally:
            if output_is_file:
                output.close()
    else:
        return_val = [CoverageResult(c, s, e, n, cov * scale_factor) for (c, s, e, n), cov in zip(intervals, counts)]
    if verbose:
        sys.stderr.write(f"coverage took {time.time() - start_time} s to complete\n")
    return return_val
