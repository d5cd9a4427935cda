"""Breakpoint-motif features - API mirror of reference frag/_breakpoint_motifs.py.

k-mers centred on the two breakpoints of each fragment (``k//2`` bases either side).  The
per-fragment loop (frag/_breakpoint_motifs.py:120-186) is the CUDA kernel behind
``ftk_breakpoint_motif_hist_u64``; the 1 Mb-window and interval drivers are shared with end
motifs (frag/_motif_common.py:580-610, 633-687) and become one launch per contig.  Like the
reference, the length arguments are accepted and unused: membership is tabix overlap + mapq.
"""
from __future__ import annotations

from sys import stderr, stdout
from time import time

import numpy as np

from ..io.fragments import as_table
from ..utils import gen_kmers
from ._common import dist_context, group_by_contig, is_writer, resolve_length_aliases
from ._end_motifs import _ref, _strand_mode
from ._motif_common import (_BASES, _MotifFreqs, _MotifsIntervals, genome_windows, interval_motif_rows,
                            parse_intervals_arg, pooled_window_counts, write_motif_freqs)

__all__ = ["BreakpointMotifFreqs", "BreakpointMotifsIntervals", "region_breakpoint_motifs", "breakpoint_motifs",
           "interval_breakpoint_motifs"]


class BreakpointMotifFreqs(_MotifFreqs):
    """Genome-wide breakpoint-motif k-mer frequencies."""


class BreakpointMotifsIntervals(_MotifsIntervals):
    """Interval-stratified breakpoint-motif k-mer counts."""


def region_breakpoint_motifs(input_file, contig, start, stop, refseq_file, k=6, fraction_low=10, fraction_high=600,
                             both_strands=True, negative_strand=False, output_file=None, quality_threshold=30,
                             verbose=False) -> dict:
    """k-mer -> count of breakpoint motifs of fragments overlapping a region (frag/_breakpoint_motifs.py:53-196)."""
    from ..device import end_motif_hist
    if verbose:
        start_time = time()
    mode = _strand_mode(both_strands, negative_strand)
    table = as_table(input_file, refseq_file)
    ref = _ref(refseq_file)
    kmer_list = gen_kmers(k, "ACGT")
    widen = 0
    if table.has_read1(contig):
        # BAM: every fragment an indexed fetch of the region yields counts, with no fragment-level test
        # (:115-120) - the sub-table holds exactly those rows and the wider bounds let all of them through
        widen = table.fetch_reach(contig)
        table = table.fetched(contig, int(start), int(stop), fetch_only=True)
    if table.n_fragments(contig) == 0:
        return dict(zip(kmer_list, 4 ** k * [0]))
    counts = end_motif_hist(table.device(contig), ref.device_contig(contig), [int(start) - widen], [int(stop) + widen], k=k,
                            strand_mode=mode, quality_threshold=quality_threshold, breakpoint=True)
    if verbose:
        stderr.write(f"region_breakpoint_motifs took {time() - start_time} seconds to run\n")
    return dict(zip(kmer_list, counts[0].cpu().tolist()))


def breakpoint_motifs(input_file, refseq_file, k=6, min_length=50, max_length=None, both_strands=True,
                      negative_strand=False, output_file=None, quality_threshold=30, workers=1, verbose=False,
                      fraction_low=None, fraction_high=None) -> BreakpointMotifFreqs:
    """Genome-wide breakpoint-motif frequencies over 1 Mb windows (frag/_breakpoint_motifs.py:204-297)."""
    if verbose:
        start_time = time()
    resolve_length_aliases(min_length, max_length, fraction_low, fraction_high)
    mode = _strand_mode(both_strands, negative_strand)
    table = as_table(input_file, refseq_file)
    ref = _ref(refseq_file)
    ccounts = np.zeros((4 ** k,), np.float64)
    ctx = dist_context()
    if ctx is not None:   # contigs LPT-sharded over the ranks, one all_reduce(SUM) of the 4^k counts
        from ..distributed import genome_end_motif_counts
        ccounts = ccounts + genome_end_motif_counts(table, ref, k=k, strand_mode=mode, quality_threshold=quality_threshold,
                                                    ctx=ctx, breakpoint=True).astype(np.float64)
        if not is_writer(ctx):
            output_file = None
    else:
        total = None
        for chrom, chrom_length in ref.chroms.items():
            if table.n_fragments(chrom) == 0:
                continue
            total = pooled_window_counts(table, ref, chrom, genome_windows(chrom_length), k, mode, quality_threshold,
                                         breakpoint=True, total=total)
        if total is not None:
            ccounts = ccounts + total[0].cpu().numpy().astype(np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        frequencies = ccounts / np.sum(ccounts)
    results = BreakpointMotifFreqs(zip(gen_kmers(k, _BASES), frequencies), k, quality_threshold)
    write_motif_freqs(results, output_file)
    if verbose:
        stdout.write(f"breakpoint_motifs took {time() - start_time} seconds to run\n")
    return results


def interval_breakpoint_motifs(input_file, refseq_file, intervals, k=6, min_length=50, max_length=None,
                               both_strands=True, negative_strand=False, output_file=None, quality_threshold=30,
                               workers=1, verbose=False, fraction_low=None, fraction_high=None) -> BreakpointMotifsIntervals:
    """Interval-stratified breakpoint-motif counts (frag/_breakpoint_motifs.py:300-385)."""
    if verbose:
        start_time = time()
    resolve_length_aliases(min_length, max_length, fraction_low, fraction_high)
    mode = _strand_mode(both_strands, negative_strand)
    table = as_table(input_file, refseq_file)
    ref = _ref(refseq_file)
    intervals_tuples = parse_intervals_arg(intervals)
    kmer_list = gen_kmers(k, "ACGT")
    rows = np.zeros((len(intervals_tuples), 4 ** k), np.int64)
    for chrom, idx in group_by_contig([iv[0] for iv in intervals_tuples]).items():
        if table.n_fragments(chrom) == 0:
            continue
        rows[idx] = interval_motif_rows(table, ref, chrom, [intervals_tuples[i][1] for i in idx],
                                        [intervals_tuples[i][2] for i in idx], k, mode, quality_threshold,
                                        breakpoint=True)
    results = BreakpointMotifsIntervals([(iv, dict(zip(kmer_list, r.tolist()))) for iv, r in zip(intervals_tuples, rows)],
                                        k, quality_threshold)
    write_motif_freqs(results, output_file)
    if verbose:
        stdout.write(f"breakpoint_motifs took {time() - start_time} seconds to run\n")
    return results


def _cli_mds(file_path: str, sep: str = "\t", header: int = 0) -> None:
    """frag/_breakpoint_motifs.py:388-391."""
    stdout.write(f"{BreakpointMotifFreqs.from_file(file_path, 30, sep, header).motif_diversity_score()}\n")


def _cli_regional_mds(file_path: str, file_out: str, sep: str = ",", header: int = 0, miller_madow: bool = False) -> None:
    """frag/_breakpoint_motifs.py:394-404."""
    BreakpointMotifsIntervals.from_file(file_path, 30, sep, header).mds_bed(file_out, miller_madow=miller_madow)
