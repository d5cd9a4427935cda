"""Fragment-length features - API mirror of reference frag/_frag_length.py (plot_histogram excluded).

The reference builds a ``dict length -> count`` with a Python loop over the fragment stream
(frag/_frag_length.py:147-153) and does its statistics in the dict's insertion order.  Here the
CUDA kernel returns, per region, the length histogram plus the first fragment index of every
length; the dict is rebuilt on the host in first-seen order and the reference's own arithmetic
(mean / quirky median / stdev / short fraction, :156-238, :432-469) is applied to it unchanged,
so every float is bit-identical.
"""
from __future__ import annotations

import gzip
import time
import warnings
from sys import stderr, stdout
from typing import NamedTuple

import numpy as np

from ..exceptions import InvalidInputError
from ..io.fragments import as_table
from ..utils import get_intervals
from ._common import dist_context, group_by_contig, is_writer, per_fetch

__all__ = ["frag_length", "frag_length_bins", "frag_length_intervals", "FragLengthStats"]


class FragLengthStats(NamedTuple):
    contig: str
    start: int
    stop: int
    name: str
    mean: float
    median: float
    stdev: float
    minimum: int
    maximum: int
    count: int
    frac_short_reads: float


def _check_region(contig, start, stop):
    """utils/_frag_generator.py:105-110."""
    if contig is None and not (start is None and stop is None):
        if not (start == 0 and stop is None):
            raise InvalidInputError("contig should be specified if start or stop given.")


def _n_bins(frags, max_length):
    m = frags.max_len if max_length is None else min(frags.max_len, int(max_length))
    return max(m, 0) + 1


def _dict_from(hist_row: np.ndarray, first_row: np.ndarray) -> dict:
    nz = np.flatnonzero(hist_row)
    order = np.argsort(first_row[nz], kind="stable")
    return {int(nz[i]): int(hist_row[nz[i]]) for i in order}


def _region_dict(table, contig, start, stop, min_length, max_length, intersect_policy, quality_threshold) -> dict:
    """The reference's ``_distribution_from_gen`` dict (first-seen order) of one region; contig None = all."""
    from ..device import interval_hist
    merged: dict = {}
    if contig is not None and table.has_read1(contig):
        table = table.fetched(contig, start, stop)   # BAM: what an indexed fetch of the region yields
    for c in ([contig] if contig is not None else table.contigs):
        if table.n_fragments(c) == 0:
            continue
        frags = table.device(c)
        _, h, f = interval_hist(frags, [start], [stop], intersect_policy, min_length, max_length,
                                quality_threshold, n_bins=_n_bins(frags, max_length), pooled=True, first_seen=True)
        for k, v in _dict_from(h[0].cpu().numpy(), f[0].cpu().numpy()).items():
            merged[k] = merged.get(k, 0) + v
    return merged


def _find_median(val_freq_dict: dict) -> float:
    """frag/_frag_length.py:156-172, off-by-one for odd totals included."""
    val = np.array(list(val_freq_dict.keys()))
    freq = np.array(list(val_freq_dict.values()))
    order = np.argsort(val)
    val, freq = val[order], freq[order]
    cdf = np.cumsum(freq)
    total_count = cdf[-1]
    if total_count % 2 == 1:
        return float(val[np.searchsorted(cdf, total_count // 2)])
    median_indices = np.searchsorted(cdf, [total_count // 2, total_count // 2 + 1])
    return float(np.mean(val[median_indices]))


def frag_length(input_file, contig=None, start=None, stop=None, intersect_policy="midpoint", output_file=None,
                quality_threshold=30, verbose=False, reference_file=None) -> np.ndarray:
    """int32 lengths of the fragment stream, in stream order (frag/_frag_length.py:246-330)."""
    from ..device import frag_lengths, policy_code
    if verbose:
        start_time = time.time()
    policy_code(intersect_policy)
    _check_region(contig, start, stop)
    table = as_table(input_file, reference_file)
    if contig is not None and table.has_read1(contig):
        table = table.fetched(contig, start, stop)   # BAM: what an indexed fetch of the region yields
    parts = [frag_lengths(table.device(c), start, stop, intersect_policy, 0, 1000000000, quality_threshold).cpu().numpy()
             for c in ([contig] if contig is not None else table.contigs) if table.n_fragments(c)]
    lengths = np.concatenate(parts).astype(np.int32) if parts else np.array([], dtype=np.int32)
    if isinstance(output_file, str):
        if output_file.endswith(".bin"):
            with open(output_file, "wt") as out:
                lengths.tofile(out)
        elif output_file == "-":
            for line in lengths:
                stdout.write(f"{line}\n")
        else:
            raise ValueError("output_file can only have suffixes .wig or .wig.gz.")
    elif output_file is not None:
        raise TypeError(f'output_file is unsupported type "{type(input_file)}". output_file should be a string '
                        "specifying the path of the file to write output scores to.")
    if verbose:
        stderr.write(f"frag_length took {time.time() - start_time} s to complete\n")
    return lengths


def frag_length_bins(input_file, contig=None, start=None, stop=None, min_length=0, max_length=None, bin_size=1,
                     output_file=None, intersect_policy="midpoint", quality_threshold=30, summary_stats=False,
                     short_fraction=None, histogram_path=None, verbose=False, reference_file=None):
    """Binned fragment-length distribution (frag/_frag_length.py:333-508) -> (bins, counts)."""
    from ..device import policy_code
    if verbose:
        start_time = time.time()
    policy_code(intersect_policy)
    _check_region(contig, start, stop)
    table = as_table(input_file, reference_file)
    ctx = dist_context()
    if ctx is not None and contig is None:
        # genome-wide dict over the ranks of one box: contigs LPT-sharded, SUM of histograms + MIN of
        # first-seen keys (frag/_frag_length.py:408-421 streams every contig in file order)
        from ..distributed import genome_length_distribution
        frag_len_dict = genome_length_distribution(table, min_length, max_length, intersect_policy,
                                                   quality_threshold, ctx=ctx, region=(start, stop))
    else:
        frag_len_dict = _region_dict(table, contig, start, stop, min_length, max_length, intersect_policy,
                                     quality_threshold)
    if not is_writer(ctx):
        output_file = None
    total_count = sum(frag_len_dict.values())
    if total_count == 0:
        warnings.warn("No fragments found in the specified region. Returning empty result.", RuntimeWarning,
                      stacklevel=2)
        return np.array([]), np.array([])
    mean = sum(value * count for value, count in frag_len_dict.items()) / total_count
    variance = sum(count * ((value - mean) ** 2) for value, count in frag_len_dict.items()) / total_count
    stats = [("mean", mean), ("median", _find_median(frag_len_dict)), ("stdev", variance ** 0.5),
             ("min", min(frag_len_dict.keys())), ("max", max(frag_len_dict.keys())), ("total count", total_count)]
    if short_fraction is not None:
        short_coverage = sum(count for length, count in frag_len_dict.items() if length <= short_fraction)
        stats.append((f"short fraction (s{short_fraction})", short_coverage / total_count))
    bin_start, bin_stop = min(frag_len_dict.keys()), max(frag_len_dict.keys())
    n_bins = (bin_stop - bin_start) // bin_size
    bins = np.arange(bin_start, bin_stop + bin_size, bin_size)
    lengths_arr = np.fromiter(frag_len_dict.keys(), dtype=np.int64)
    freqs_arr = np.fromiter(frag_len_dict.values(), dtype=np.int64)
    counts_arr = np.zeros(n_bins + 1, dtype=np.int64)
    np.add.at(counts_arr, (lengths_arr - bin_start) // bin_size, freqs_arr)
    counts = counts_arr.tolist()
    if output_file is not None:
        out_is_file = False
        try:
            if output_file == "-":
                out = stdout
            elif output_file.endswith(".gz"):
                out_is_file, out = True, gzip.open(output_file, "wt")
            else:
                out_is_file, out = True, open(output_file, "w")
            out.write("min\tmax\tcount\n")
            for bin_val, count in zip(bins, counts):
                out.write(f"{bin_val}\t{bin_val + bin_size - 1}\t{count}\n")
            if summary_stats:
                for name, value in stats:
                    out.write(f"#{name}: {value}\n")
        finally:
            if out_is_file:
                out.close()
    if histogram_path is not None:
        raise NotImplementedError("histogram plotting (matplotlib) is outside the B200 hot path")
    if verbose:
        stderr.write(f"frag_length_bins took {time.time() - start_time} s to complete.\n")
    return bins, counts


def _stats_from_dict(contig, start, stop, name, frag_len_dict, short_reads) -> FragLengthStats:
    """frag/_frag_length.py:204-238."""
    total_count = sum(frag_len_dict.values())
    if total_count == 0:
        return FragLengthStats(contig, start, stop, name, -1, -1, -1, -1, -1, -1, -1)
    mean = sum(value * count for value, count in frag_len_dict.items()) / total_count
    median = _find_median(frag_len_dict)
    variance = sum(count * ((value - mean) ** 2) for value, count in frag_len_dict.items()) / total_count
    n_short = sum(count for length, count in frag_len_dict.items() if length <= short_reads)
    return FragLengthStats(contig, start, stop, name, mean, median, variance ** 0.5, min(frag_len_dict.keys()),
                           max(frag_len_dict.keys()), total_count, n_short / total_count)


def _length_stats(hist: np.ndarray, first: np.ndarray, short_reads: int):
    """(mean, median, stdev, min, max, count, frac_short) columns for the rows of ``hist`` / ``first``."""
    from ctypes import POINTER, c_double, c_int32, c_int64
    from .._lib import check, lib
    hist = np.ascontiguousarray(hist, dtype=np.int32); first = np.ascontiguousarray(first, dtype=np.int32)
    n, nb = hist.shape
    f64 = [np.empty(n, np.float64) for _ in range(4)]
    i64 = [np.empty(n, np.int64) for _ in range(3)]
    p = lambda a, t: a.ctypes.data_as(POINTER(t))   # noqa: E731
    check(lib().ftk_length_stats_host(p(hist, c_int32), p(first, c_int32), n, nb, int(short_reads), 0,
                                      p(f64[0], c_double), p(f64[1], c_double), p(f64[2], c_double),
                                      p(i64[0], c_int64), p(i64[1], c_int64), p(i64[2], c_int64), p(f64[3], c_double)),
          "ftk_length_stats_host")
    return (f64[0].tolist(), f64[1].tolist(), f64[2].tolist(), i64[0].tolist(), i64[1].tolist(), i64[2].tolist(),
            f64[3].tolist())


def _stats_row(contig, start, stop, name, cols, k) -> FragLengthStats:
    if cols[5][k] < 0:   # empty interval: the reference's all -1 row (ints)
        return FragLengthStats(contig, start, stop, name, -1, -1, -1, -1, -1, -1, -1)
    return FragLengthStats(contig, start, stop, name, cols[0][k], cols[1][k], cols[2][k], cols[3][k], cols[4][k],
                           cols[5][k], cols[6][k])


def frag_length_intervals(input_file, interval_file, output_file=None, min_length=0, max_length=None,
                          quality_threshold=30, intersect_policy="midpoint", short_reads=150, workers=1,
                          verbose=False, reference_file=None):
    """Per-interval fragment-length statistics (frag/_frag_length.py:511-640)."""
    from ..device import interval_hist, policy_code, torch
    if verbose:
        start_time = time.time()
    policy_code(intersect_policy)
    table = as_table(input_file, reference_file)
    intervals = get_intervals(interval_file)
    results: list = [None] * len(intervals)
    for contig, idx in group_by_contig([iv[0] for iv in intervals]).items():
        def run(tab, sel, contig=contig, idx=idx):
            if tab.n_fragments(contig) == 0:
                return [_stats_from_dict(*intervals[idx[k]], {}, short_reads) for k in sel]
            frags = tab.device(contig)
            nb = _n_bins(frags, max_length)
            batch = max(1, (256 << 20) // (12 * nb))  # bound the per-interval histogram block to ~256 MB
            rows = []
            for b0 in range(0, len(sel), batch):
                sub = [idx[k] for k in sel[b0: b0 + batch]]
                _, h, f = interval_hist(frags, [intervals[i][1] for i in sub], [intervals[i][2] for i in sub],
                                        intersect_policy, min_length, max_length, quality_threshold, n_bins=nb,
                                        first_seen=True)
                # the statistics of all intervals of the batch in one native call (csrc/ftk_hoststats.cu):
                # the reference's arithmetic and operation order, without a Python loop per interval
                cols = _length_stats(h.to(torch().int32).cpu().numpy(), f.cpu().numpy(), short_reads)
                rows += [_stats_row(*intervals[i], cols, k) for k, i in enumerate(sub)]
            return rows

        got = per_fetch(table, contig, [intervals[i][1] for i in idx], [intervals[i][2] for i in idx], run)
        for i, row in zip(idx, got):
            results[i] = row
    if output_file is not None:
        output_is_file = False
        try:
            if output_file.endswith(".bed") or output_file.endswith(".bedgraph"):
                output_is_file, output = True, open(output_file, "w")
            elif output_file.endswith(".bed.gz"):
                output_is_file, output = True, gzip.open(output_file, "wt")
            elif output_file == "-":
                output = stdout
            else:
                raise ValueError("The output file should have .bed or .bed.gz as as suffix.")
            output.write(f"contig\tstart\tstop\tname\tmean\tmedian\tstdev\tmin\tmax\tcount\ts{short_reads}\n")
            output.write("\n".join("\t".join(str(element) for element in item) for item in results))
            output.write("\n")
        finally:
            if output_is_file:
                output.close()
    if verbose:
        stderr.write(f"Calculating fragment length statistics for intervals took {time.time() - start_time} s\n")
    return results
