"""Windowed Protection Score over one interval - API mirror of reference frag/_wps.py:56-229.

Same signature, return dtype, warnings and errors; the per-position numba loop
(frag/_wps.py:176-188) is replaced by the CUDA tile kernel behind
``ftk_wps_tiles_i32`` (finaletoolkit_b200/csrc/ftk_wps.cu).
"""
from __future__ import annotations

import gzip
import time
import warnings
from sys import stderr, stdout

import numpy as np

from ..io.fragments import as_table
from ._common import resolve_length_aliases

__all__ = ["wps"]

_WPS_DTYPE = [("contig", "U16"), ("start", "i8"), ("wps", "i8")]


def _wps_device(table, chrom, starts, stops, chrom_size, window_size, min_length, max_length,
                quality_threshold, device=None):
    """int32 CUDA tensor with the WPS of every interval back to back + host offsets."""
    from ..device import WpsPlan
    frags = table.device(chrom, device)
    plan = WpsPlan(starts, stops, chrom_size, max_length, frags.device)
    return plan.run(frags, window_size, min_length, max_length, quality_threshold), plan.offsets


def fetch_window(start, stop, max_length, chrom_size):
    """frag/_wps.py:156-157: the region ``wps`` fetches fragments from."""
    return max(round(start - max_length), 0), min(round(stop + max_length), chrom_size)


_STREAM_MIN_FRAGMENTS = 1 << 21     # below this a contig's columns are uploaded whole (nothing to overlap)


def _wps_streamed(table, chrom, starts, stops, chrom_size, window_size, min_length, max_length,
                  quality_threshold, device=None):
    """WPS of every interval of one large contig through the streamed pipeline: the contig's host columns are
    page-locked in place, chunks of intervals go H2D (double-buffered) || kernels || D2H on three streams
    (``pipeline.StreamedContig``) and the scores arrive in pinned host memory as int16 (a quarter of the int64
    the reference returns, half of int32 on PCIe).  Returns ``(host scores, offsets)`` or ``None`` when the
    contig is small, its intervals are not start-sorted, its pages cannot be locked or a score does not fit
    int16 (a > 32767-fold pile-up) - the caller then takes the resident path.  Same numbers either way."""
    from ..device import require_cuda
    from ..pipeline import StreamedContig
    n = table.n_fragments(chrom)
    s = np.asarray(starts, dtype=np.int64)
    if n < _STREAM_MIN_FRAGMENTS or len(s) < 16 or np.any(s[1:] < s[:-1]):
        return None
    cols = table.pinned(chrom)
    if cols is None:
        return None
    h_start, h_stop, h_mapq = cols[:3]
    try:
        pipe = StreamedContig(h_start, h_stop, h_mapq, s, np.asarray(stops, dtype=np.int64), int(chrom_size),
                              int(window_size), min_length, int(max_length), int(quality_threshold), n_chunks=8,
                              device=require_cuda(device), coverage=False, length_hist=False, wps_dtype="int16")
        w, _, _, _ = pipe.run()
    except OverflowError:
        return None
    return w.numpy()[: pipe.n_positions], pipe.offsets


def wps(input_file, chrom, start, stop, chrom_size, output_file=None, window_size=120, min_length=120,
        max_length=180, quality_threshold=30, verbose=0, fraction_low=None, fraction_high=None,
        reference_file=None) -> np.ndarray:
    """Raw WPS over ``chrom:[start, stop)``: structured array ('contig','start','wps')."""
    if verbose:
        start_time = time.time()
        stderr.write("[finaletoolkit-wps] Reading fragments\n")
        stderr.write(f"Region: {chrom}:{start}-{stop}\n")
    min_length, max_length = resolve_length_aliases(min_length, max_length, fraction_low, fraction_high)
    start, stop = int(start), int(stop)
    if stop <= start:
        warnings.warn(f"[wps] {chrom}:{start}-{stop} is a degenerate interval (stop <= start); skipping.",
                      UserWarning, stacklevel=2)
        return np.zeros(0, dtype=_WPS_DTYPE)
    table = as_table(input_file, reference_file)
    if table.has_read1(chrom):   # BAM: the rows an indexed fetch of the padded window yields (frag/_wps.py:156-168)
        table = table.fetched(chrom, *fetch_window(start, stop, max_length, chrom_size))
    out, _ = _wps_device(table, chrom, [start], [stop], int(chrom_size), int(window_size), min_length,
                         int(round(max_length)), quality_threshold)
    scores = np.zeros(stop - start, dtype=_WPS_DTYPE)
    scores["contig"] = chrom
    scores["start"] = np.arange(start, stop, dtype=np.int64)
    scores["wps"] = out.cpu().numpy()
    if isinstance(output_file, str):
        if verbose:
            stderr.write("Writing to output file.\n")
        _write_wig(output_file, chrom, start, stop, scores)
    elif output_file is not None:
        raise TypeError(f'output_file is unsupported type "{type(input_file)}". output_file should be a '
                        "string specifying the path of the file to output scores to.")
    if verbose:
        stderr.write(f"wps took {time.time() - start_time} s to complete\n")
    return scores


def _write_wig(output_file, chrom, start, stop, scores) -> None:
    """fixedStep WIG exactly as frag/_wps.py:208-229."""
    header = f"fixedStep\tchrom={chrom}\tstart={start}\tstep={1}\tspan={stop - start}\n"
    body = "".join(f"{s}\n" for s in scores["wps"])
    if output_file.endswith(".wig.gz"):
        with gzip.open(output_file, "wt") as out:
            out.write(header + body)
    elif output_file.endswith(".wig"):
        with open(output_file, "wt") as out:
            out.write(header + body)
    elif output_file == "-":
        stdout.write(header + body)
        stdout.flush()
    else:
        raise ValueError("output_file can only have suffixes .wig or .wig.gz.")
