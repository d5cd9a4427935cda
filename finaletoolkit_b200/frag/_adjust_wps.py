"""adjust_wps - API mirror of reference frag/_adjust_wps.py:166-297.

Reads raw WPS from a bigWig, subtracts a running median/mean, applies Savitzky-Golay and
writes a bigWig.  The driver logic (BED parsing, the merge rule for intervals that would
overlap after trimming, contiguity check, error handling) follows the reference; the numeric
core of ``_single_adjust_wps`` (frag/_adjust_wps.py:119-140) runs in one batched CUDA call
for all intervals (``ftk_adjust_wps_f64``).
"""
from __future__ import annotations

import gzip
from sys import stderr
from time import time

import numpy as np

from ..io import bigwig as pbw
from ..utils import chrom_sizes_to_list

__all__ = ["adjust_wps"]


def _read_adjust_sites(interval_file, interval_size, median_window_size):
    """frag/_adjust_wps.py:219-263."""
    left_of_site = round(-interval_size / 2)
    right_of_site = round(interval_size / 2)
    assert right_of_site - left_of_site == interval_size
    if not (interval_file.endswith(".bed") or interval_file.endswith(".bed.gz")):
        raise ValueError("Invalid filetype for interval_file.")
    end_decrease = median_window_size // 2
    intervals = []
    opener = gzip.open if interval_file.endswith(".gz") else open
    with opener(interval_file, "rt") as file:
        for line in file:
            contents = line.split("\t")
            contig = contents[0].strip()
            midpoint = (int(contents[1]) + int(contents[2])) // 2
            start = max(0, midpoint + int(left_of_site))
            stop = midpoint + int(right_of_site)
            if (len(intervals) > 0 and intervals[-1][0] == contig
                    and intervals[-1][2] - end_decrease > start + end_decrease):
                start = intervals[-1][1]
                intervals.pop(-1)
            intervals.append((contig, int(start), int(stop)))
    return intervals


def adjust_wps(input_file, interval_file, output_file, chrom_sizes, interval_size=5000, median_window_size=1000,
               savgol_window_size=21, savgol_poly_deg=2, savgol=True, mean=False, subtract_edges=False,
               edge_size=500, workers=1, verbose=False) -> None:
    """Adjust raw WPS in a bigWig with median/mean and Savitzky-Golay filters."""
    from ..device import adjust_segments
    if verbose:
        start_time = time()
        stderr.write("Reading intervals from bed...\n")
    intervals = _read_adjust_sites(interval_file, interval_size, median_window_size)
    if not input_file.endswith(".bw"):
        raise ValueError("Invalid filetype for input_file.")
    raw_wps = pbw.open(input_file, "r")
    try:
        seg_vals, seg_pos, seg_contig = [], [], []
        for k, (contig, start, stop) in enumerate(intervals):
            if k == 0 or contig != intervals[k - 1][0]:
                # one multi-threaded inflate of every section this contig's intervals touch; the
                # inflated sections of the previous contig are dropped
                raw_wps.drop_cache()
                j = k
                while j < len(intervals) and intervals[j][0] == contig:
                    j += 1
                raw_wps.prefetch(intervals[k:j])
            try:
                # per-base WPS tracks (what multi_wps writes) come back as (first position, values); anything
                # else through the general query with its contiguity check (frag/_adjust_wps.py:85-117)
                run = raw_wps.per_base_run(contig, start, stop)
                rng = None if run is not None else raw_wps.intervals_arrays(contig, start, stop)
            except RuntimeError:  # frag/_adjust_wps.py:145-153: invalid interval -> skipped
                stderr.write(f"Invalid interval detected:\n{contig}:{start}-{stop}. This interval will be skipped.\n")
                continue
            if run is not None:
                first, scores = run
            elif rng is None:
                stderr.write(f"No entries in range: {contig}:{start}-{stop}. This interval will be skipped.\n")
                continue
            else:
                starts, _, scores = rng
                if not np.all(starts[:-1] + 1 == starts[1:]):
                    raise ValueError("BigWig was found to be nonsequential. There may be multiple entries for one "
                                     "position or gaps in the regions specified in the interval file.")
                first = int(starts[0])
            if median_window_size > scores.shape[0]:
                raise ValueError(f"median_window_size ({median_window_size}) cannot be greater than the length "
                                 f"of interval ({scores.shape[0]}).")
            seg_vals.append(scores.astype(np.float32)); seg_pos.append((first, scores.shape[0])); seg_contig.append(contig)
    finally:
        raw_wps.close()

    outputs = []
    if seg_vals:
        out, off = adjust_segments(np.concatenate(seg_vals), [len(v) for v in seg_vals],
                                   median_window_size=median_window_size, use_mean=mean, savgol=savgol,
                                   savgol_window_size=savgol_window_size, savgol_poly_deg=savgol_poly_deg,
                                   subtract_edges=subtract_edges, edge_size=edge_size)
        host = out.cpu().numpy()
        h = median_window_size // 2
        for k, (contig, (first, _)) in enumerate(zip(seg_contig, seg_pos)):
            outputs.append((contig, first + h, host[off[k]: off[k + 1]]))
    if verbose:
        stderr.write("Writing to output\n")
    with pbw.open(output_file, "w") as output_bw:
        output_bw.addHeader(chrom_sizes_to_list(chrom_sizes))
        for contig, first, values in outputs:
            n = len(values)
            if n == 0:
                continue
            try:
                # the positions are one checked-contiguous run (see above): the per-base entries of
                # frag/_adjust_wps.py:275-283 as one fixedStep block - same intervals for every reader, same
                # bounds / order errors from the writer
                output_bw.addEntries(contig, first, values=values, span=1, step=1)
            except RuntimeError as e:  # frag/_adjust_wps.py:285-291
                stderr.write(f"RuntimeError encountered while writing to {output_file} at interval "
                             f"{contig}:{first}-{first + n}: {e}\n")
    if verbose:
        stderr.write(f"Adjust-WPS took {time() - start_time} s to run.\n")
