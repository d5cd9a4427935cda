"""WPS over many BED sites - API mirror of reference frag/_multi_wps.py:31-341.

The reference fans ``wps()`` out over a ``multiprocessing.Pool`` (one file re-open and
one O(len x n_frag) loop per interval); here the intervals of each contig go to the
GPU in ONE launch and ``workers`` is accepted for compatibility only.
"""
from __future__ import annotations

import time
import warnings
from sys import stderr, stdin

import numpy as np

from ..io import bigwig as pbw
from ..io.fragments import FragmentTable, as_table
from ..io.textout import GzipTextWriter, bedgraph_text
from ..utils import chrom_sizes_to_list
from ._common import group_by_contig, per_fetch, resolve_length_aliases
from ._wps import _wps_device, _wps_streamed, fetch_window

__all__ = ["multi_wps", "LAST_TIMINGS"]

# wall-clock split of the most recent multi_wps call (seconds): decode = file -> columns, compute = upload +
# kernels + download, write = output file; read by bench.py's api_wall block
LAST_TIMINGS: dict = {}


def _read_header(input_file, table: FragmentTable, chrom_sizes):
    """frag/_multi_wps.py:226-237: BAM/CRAM header, else chrom.sizes is mandatory."""
    if table.is_sam and table.contig_lengths:
        return list(table.contig_lengths.items())
    if chrom_sizes is None:
        raise ValueError("chrom_sizes must be specified for BED/Fragment files")
    return chrom_sizes_to_list(chrom_sizes)


def _read_sites(site_bed, interval_size, references, chrom_sizes_dict):
    """frag/_multi_wps.py:240-297: centred windows, previous window truncated on overlap."""
    contigs, starts, stops = [], [], []
    left_of_site = round(-interval_size / 2)
    right_of_site = round(interval_size / 2)
    assert right_of_site - left_of_site == interval_size
    bed = stdin if site_bed == "-" else open(site_bed)
    try:
        prev_contig, prev_start, prev_stop = None, 0, 0
        for line in bed:
            contents = line.split()
            contig = contents[0].strip()
            if int(contents[1]) > int(contents[2]):
                raise ValueError(
                    f"[multi_wps] {contig}:{contents[1]}-{contents[2]} is invalid. Please be sure start "
                    f"coordinate occurs before stop for all intervals in {site_bed}.")
            if contig not in references:
                warnings.warn(f"Skipping site {contig}:{int(contents[1])} from site_bed (chrom not in chrom_sizes)",
                              UserWarning)
                continue
            midpoint = (int(contents[1]) + int(contents[2])) // 2
            start = max(0, midpoint + int(left_of_site))
            stop = min(midpoint + int(right_of_site), chrom_sizes_dict[contig])
            if contig == prev_contig and start < prev_stop:
                prev_stop = start
            if prev_contig is not None and prev_stop > prev_start:
                contigs.append(prev_contig); starts.append(prev_start); stops.append(prev_stop)
            prev_contig, prev_start, prev_stop = contig, start, stop
        if prev_stop > prev_start:
            contigs.append(prev_contig); starts.append(prev_start); stops.append(prev_stop)
    finally:
        if site_bed != "-":
            bed.close()
    return contigs, starts, stops


def multi_wps(input_file, site_bed, chrom_sizes=None, output_file=None, window_size=120, interval_size=5000,
              min_length=120, max_length=180, quality_threshold=30, workers=1, verbose=0, fraction_low=None,
              fraction_high=None, reference_file=None):
    """Aggregate WPS over the sites of a BED file; writes ``.bw`` / ``.bed.gz`` / ``bedGraph.gz``."""
    if verbose:
        start_time = time.time()
        stderr.write(f"Calculating aggregate WPS\ninput_file: {input_file}\nsite_bed: {site_bed}\n"
                     f"output_file: {output_file}\nwindow_size: {window_size}\ninterval_size: {interval_size}\n"
                     f"quality_threshold: {quality_threshold}\nworkers: {workers}\n")
    if input_file == "-" and site_bed == "-":
        raise ValueError("input_file and site_bed cannot both read from stdin")
    min_length, max_length = resolve_length_aliases(min_length, max_length, fraction_low, fraction_high)
    t_begin = time.perf_counter()
    table = as_table(input_file, reference_file)
    header = _read_header(input_file, table, chrom_sizes)
    references = [chrom for chrom, _ in header]
    chrom_sizes_dict = dict(header)
    contigs, starts, stops = _read_sites(site_bed, interval_size, references, chrom_sizes_dict)
    if header and contigs:  # bigWig wants header order (frag/_multi_wps.py:149-160); sorted() is stable
        chrom_order = {chrom: idx for idx, (chrom, _) in enumerate(header)}
        order = sorted(range(len(contigs)), key=lambda i: (chrom_order.get(contigs[i], len(header)), starts[i]))
        contigs = [contigs[i] for i in order]; starts = [starts[i] for i in order]; stops = [stops[i] for i in order]

    if isinstance(output_file, str):
        if not (output_file.endswith(".bw") or output_file.endswith(".bed.gz") or output_file.endswith("bedGraph.gz")):
            raise ValueError("output_file can only have suffix .bw")
    elif output_file is not None:
        raise TypeError(f'output_file is unsupported type "{type(input_file)}". output_file should be a '
                        "string specifying the path of the file to output scores to.")

    # one GPU launch per contig, results kept in interval order
    results: list = [None] * len(contigs)
    n_streamed = 0
    t_decoded = time.perf_counter()
    for contig, idx in group_by_contig(contigs).items():
        table.host(contig)      # lazy tables decode here: keep it out of the compute share
        size, width = chrom_sizes_dict[contig], int(round(max_length))

        def run(tab, sel, contig=contig, idx=idx, size=size, width=width):
            nonlocal n_streamed
            args = (tab, contig, [starts[idx[k]] for k in sel], [stops[idx[k]] for k in sel], size,
                    int(window_size), min_length, width, quality_threshold)
            # large contigs stream (pinned columns, chunked H2D || kernels || D2H, int16 scores); small ones upload whole
            streamed = _wps_streamed(*args)
            if streamed is not None:
                host, off = streamed
                n_streamed += 1
            else:
                out, off = _wps_device(*args)
                host = out.cpu().numpy()
            return [host[off[k]: off[k + 1]] for k in range(len(sel))]

        # every interval is one wps() call of the reference, which fetches from its own padded window
        bam = table.has_read1(contig)
        windows = [fetch_window(starts[i], stops[i], max_length, size) if bam else (None, None) for i in idx]
        got = per_fetch(table, contig, [w[0] for w in windows], [w[1] for w in windows], run)
        for i, scores in zip(idx, got):
            results[i] = scores

    t_computed = time.perf_counter()
    if isinstance(output_file, str):
        if output_file.endswith(".bw"):
            with pbw.open(output_file, "w") as bigwig:
                bigwig.addHeader(header)
                for contig, start, scores in zip(contigs, starts, results):
                    if scores.shape == (0,):
                        continue
                    try:
                        bigwig.addEntries(contig, start, values=scores.astype(np.float64), step=1, span=1)
                    except RuntimeError:  # frag/_multi_wps.py:319-325
                        stderr.write(f"{contig}:{start}-{stops[-1]}\n")
                        stderr.write("/n invalid or out of order interval encountered. Skipping to next.\n")
                        continue
        else:
            # frag/_multi_wps.py:328-341, same text.  The lines of the intervals are formatted by a pool of
            # host threads (the native formatter releases the GIL; one 5-kb interval is too short to split) and
            # handed to the writer in interval order
            import os
            from concurrent.futures import ThreadPoolExecutor
            with GzipTextWriter(output_file) as bedgraph, ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as pool:
                fmt = lambda i: bedgraph_text(contigs[i], starts[i], results[i])          # noqa: E731
                for b0 in range(0, len(contigs), 512):        # bounded look-ahead: 512 intervals of text in memory
                    for text in pool.map(fmt, range(b0, min(b0 + 512, len(contigs)))):
                        bedgraph.write(text)
    t_end = time.perf_counter()
    LAST_TIMINGS.clear()
    LAST_TIMINGS.update(decode=t_decoded - t_begin, compute=t_computed - t_decoded, write=t_end - t_computed,
                        total=t_end - t_begin, streamed_contigs=n_streamed)
    if verbose:
        stderr.write(f"multi_wps took {time.time() - start_time} s to complete\n")
    return output_file
