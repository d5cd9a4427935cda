"""agg_bw - API mirror of reference utils/_agg_bw.py:18-146.

Sums a bigWig signal over equal-length BED6 intervals, flipping '-' strand intervals, and writes
a fixedStep WIG.  The per-interval ``pyBigWig.values`` + running numpy sum of the reference
(utils/_agg_bw.py:84-123) becomes one multi-threaded section inflate (``BigWigReader.prefetch``)
and one CUDA call (``ftk_agg_signal_f64``) that adds the intervals in file order, so the fp64
result is bit-identical.
"""
from __future__ import annotations

import gzip
import time
from sys import stderr

import numpy as np

from ..io import bigwig as pbw

__all__ = ["agg_bw"]


def agg_bw(input_file, interval_file, output_file, median_window_size: int = 1, mean: bool = False,
           verbose: bool = False) -> np.ndarray:
    """Aggregate bigWig signal over BED intervals (strand-aware); returns the per-position signal."""
    from ..device import agg_signal
    if verbose:
        start_time = time.time()
        stderr.write("Reading intervals from bed...\n")
    if not (str(interval_file).endswith(".bed") or str(interval_file).endswith(".bed.gz")):
        raise ValueError("Invalid filetype for interval_file.")
    intervals = []
    opener = gzip.open if str(interval_file).endswith(".gz") else open
    with opener(interval_file, "rt") as file:
        for line in file:
            contents = line.split("\t")
            intervals.append((contents[0], int(contents[1]), int(contents[2]), contents[5].strip()))

    # interval length after trimming by the median-filter window (utils/_agg_bw.py:86)
    interval_size = intervals[0][2] - intervals[0][1] - median_window_size
    lo, hi = median_window_size // 2, -median_window_size // 2     # values[lo:hi], like the reference
    rows, strands = [], []
    with pbw.open(str(input_file), "r") as raw_wps:
        raw_wps.prefetch([(c, s, e) for c, s, e, _ in intervals])
        for contig, start, stop, strand in intervals:
            try:
                values = raw_wps.values(contig, start, stop)
            except RuntimeError as e:   # invalid bounds / unknown contig: printed and skipped
                print(e)
                continue
            trimmed_len = len(range(len(values))[lo:hi])
            if trimmed_len != interval_size:
                print(f"Trimmed size {trimmed_len} for {contig}:{start}-{stop} is not equal to "
                      f"interval size {interval_size}. Skipping.")
                continue
            if strand in ("+", "-"):
                rows.append(values)
                strands.append(1 if strand == "+" else -1)
            elif verbose:
                stderr.write("A segment without strand was encountered. Skipping.")

    if rows:
        agg_scores = agg_signal(np.stack(rows), strands, lo, interval_size).cpu().numpy()
    else:
        agg_scores = np.zeros(interval_size, dtype=np.int64)
    if mean:
        agg_scores = agg_scores / len(rows)

    if str(output_file).endswith("wig"):
        with open(output_file, "wt") as out:
            out.write(f"fixedStep\tchrom=.\tstart={-interval_size // 2}\tstep={1}\tspan={interval_size}\n")
            out.write("".join(f"{score}\n" for score in agg_scores))
    else:
        raise ValueError("The output_file is an unaccepted type. Must be a wiggle file ending in .wig")
    if verbose:
        stderr.write(f"Aggregating bigWig took {time.time() - start_time} s to run.\n")
    return agg_scores
