"""5' end-motif features - API mirror of reference frag/_end_motifs.py.

The per-fragment loop with two py2bit calls and a dict increment
(frag/_end_motifs.py:128-179) is the CUDA kernel behind ``ftk_end_motif_hist_u64``; the Pool
drivers over 1 Mb windows / intervals (frag/_motif_common.py:580-610, 633-687) become one
launch per contig (windows keep their overlap double-counting; genome-wide counts pool on the
device).  ``min_length``/``max_length`` are accepted and - like the reference - unused beyond
the ``< k`` clamp.
"""
from __future__ import annotations

import warnings
from sys import stderr, stdout
from time import time

import numpy as np

from ..io.fragments import as_table
from ..io.reference import ReferenceWrapper, open_reference
from ..utils import gen_kmers
from ._common import dist_context, group_by_contig, is_writer, resolve_length_aliases
from ._motif_common import (MIN_QUALITY, _MotifFreqs, _MotifsIntervals, _BASES, genome_windows,
                            interval_motif_rows, parse_intervals_arg, pooled_window_counts, write_motif_freqs)

__all__ = ["EndMotifFreqs", "EndMotifsIntervals", "region_end_motifs", "end_motifs", "interval_end_motifs",
           "MIN_QUALITY"]


class EndMotifFreqs(_MotifFreqs):
    """Genome-wide 5' end-motif k-mer frequencies (Zhou et al., 2023)."""


class EndMotifsIntervals(_MotifsIntervals):
    """Interval-stratified 5' end-motif k-mer counts."""


def _strand_mode(both_strands, negative_strand) -> int:
    if both_strands and negative_strand:
        raise ValueError("Cannot have both both_strands and negative_strand.")
    return 0 if both_strands else (2 if negative_strand else 1)


def _ref(refseq_file) -> ReferenceWrapper:
    return open_reference(refseq_file)


def region_end_motifs(input_file, contig, start, stop, refseq_file, k=4, fraction_low=50, fraction_high=None,
                      both_strands=True, negative_strand=False, output_file=None, quality_threshold=MIN_QUALITY,
                      verbose=False) -> dict:
    """k-mer -> count of 5' end motifs of fragments overlapping a region (frag/_end_motifs.py:51-187)."""
    from ..device import end_motif_hist
    if verbose:
        start_time = time()
    mode = _strand_mode(both_strands, negative_strand)
    if fraction_low < k:  # TypeError for None, like the reference
        warnings.warn(f"fraction_low={fraction_low} < k={k}, which may cause errors. Automatically setting "
                      "fraction_low=k.")
        fraction_low = k
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
                            strand_mode=mode, quality_threshold=quality_threshold)
    if verbose:
        stderr.write(f"region_end_motifs took {time() - start_time} seconds to run\n")
    return dict(zip(kmer_list, counts[0].cpu().tolist()))


def end_motifs(input_file, refseq_file, k=4, min_length=50, max_length=None, both_strands=True,
               negative_strand=False, output_file=None, quality_threshold=30, workers=1, verbose=False,
               fraction_low=None, fraction_high=None) -> EndMotifFreqs:
    """Genome-wide 5' end-motif frequencies over 1 Mb windows (frag/_end_motifs.py:198-293)."""
    if verbose:
        start_time = time()
    min_length, max_length = resolve_length_aliases(min_length, max_length, fraction_low, fraction_high)
    if min_length is not None and min_length < k:
        warnings.warn(f"min_length={min_length} < k={k}, which may cause errors. Automatically setting min_length=k.")
        min_length = k
    mode = _strand_mode(both_strands, negative_strand)
    if min_length is None:
        raise TypeError("'<' not supported between instances of 'NoneType' and 'int'")  # frag/_end_motifs.py:108
    table = as_table(input_file, refseq_file)
    ref = _ref(refseq_file)
    ctx = dist_context()
    ccounts = np.zeros((4 ** k,), np.float64)
    if ctx is not None:
        # contigs LPT-sharded over the ranks, one all_reduce(SUM) of the 4^k counts
        # (the reference pools its per-window dicts on the host, frag/_motif_common.py:599-609)
        from ..distributed import genome_end_motif_counts
        ccounts = ccounts + genome_end_motif_counts(table, ref, k=k, strand_mode=mode,
                                                    quality_threshold=quality_threshold, ctx=ctx).astype(np.float64)
        if not is_writer(ctx):
            output_file = None
    else:
        total = None
        for chrom, chrom_length in ref.chroms.items():
            if table.n_fragments(chrom) == 0:
                continue
            total = pooled_window_counts(table, ref, chrom, genome_windows(chrom_length), k, mode, quality_threshold,
                                         breakpoint=False, total=total)
        if total is not None:
            ccounts = ccounts + total[0].cpu().numpy().astype(np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        frequencies = ccounts / np.sum(ccounts)
    results = EndMotifFreqs(zip(gen_kmers(k, _BASES), frequencies), k, quality_threshold)
    write_motif_freqs(results, output_file)
    if verbose:
        stdout.write(f"end_motifs took {time() - start_time} seconds to run\n")
    return results


def interval_end_motifs(input_file, refseq_file, intervals, k=4, min_length=50, max_length=None, both_strands=True,
                        negative_strand=False, output_file=None, quality_threshold=30, workers=1, verbose=False,
                        fraction_low=None, fraction_high=None) -> EndMotifsIntervals:
    """Interval-stratified 5' end-motif counts (frag/_end_motifs.py:296-383)."""
    if verbose:
        start_time = time()
    min_length, max_length = resolve_length_aliases(min_length, max_length, fraction_low, fraction_high)
    if min_length is not None and min_length < k:
        warnings.warn(f"min_length={min_length} < k={k}, which may cause errors. Automatically setting min_length=k.")
        min_length = k
    mode = _strand_mode(both_strands, negative_strand)
    if min_length is None:
        raise TypeError("'<' not supported between instances of 'NoneType' and 'int'")
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
                                        breakpoint=False)
    results = EndMotifsIntervals([(iv, dict(zip(kmer_list, r.tolist()))) for iv, r in zip(intervals_tuples, rows)],
                                 k, quality_threshold)
    write_motif_freqs(results, output_file)
    if verbose:
        stdout.write(f"end_motifs took {time() - start_time} seconds to run\n")
    return results


def _cli_mds(file_path: str, sep: str = "\t", header: int = 0) -> None:
    """frag/_end_motifs.py:386-390."""
    stdout.write(f"{EndMotifFreqs.from_file(file_path, 30, sep, header).motif_diversity_score()}\n")


def _cli_regional_mds(file_path: str, file_out: str, sep: str = ",", header: int = 0, miller_madow: bool = False) -> None:
    """frag/_end_motifs.py:393-402."""
    EndMotifsIntervals.from_file(file_path, 30, sep, header).mds_bed(file_out, miller_madow=miller_madow)
