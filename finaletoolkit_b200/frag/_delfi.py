"""DELFI short/long fragment counts per bin - API mirror of reference frag/_delfi.py.

The hot part of ``delfi`` is ``_delfi_single_window`` (frag/_delfi.py:404-511): for each 100 kb
bin a Pool worker streams the bin's fragments, applies length / midpoint / blacklist / gap tests
and counts G+C bases of the bin.  Here every bin of a contig goes through one CUDA call
(``ftk_delfi_windows_u64``); the table post-processing (ratio, hg19 no-coverage rows, 100 kb ->
5 Mb merge) stays pandas on the host and follows the reference's conventions, including its
column dtypes.  GC correction needs the third-party ``loess`` package exactly like the reference
(frag/_delfi_gc_correct.py:12); it is imported only when requested.
"""
from __future__ import annotations

import gzip
import time
import warnings
from collections import defaultdict
from sys import stderr, stdout

import numpy as np
import pandas

from ..genome.gaps import GenomeGaps
from ..io.fragments import as_table
from ..io.reference import open_reference
from ..utils import chrom_sizes_to_list
from ._common import dist_context, is_writer

__all__ = ["delfi", "delfi_gc_correct", "delfi_merge_bins", "trim_coverage"]

_COLUMNS = ["contig", "start", "stop", "arm", "short", "long", "gc", "num_frags"]
_BINS_PER_WINDOW = 50
_GC_CORRECT_COLUMNS = ["short", "long", "num_frags", "ratio"]


def trim_coverage(window_data: np.ndarray, trim_percentile: int = 10):
    """Blank the bins below the ``trim_percentile`` of num_frags (frag/_delfi.py:32-45)."""
    threshold = np.percentile(window_data["num_frags"], trim_percentile)
    trimmed = window_data.copy()
    low = window_data["num_frags"] < threshold
    for col in ("short", "long", "gc"):
        trimmed[col][low] = np.nan
    trimmed["num_frags"][low] = 0
    return trimmed


def _load_blacklist_indexed(blacklist_file) -> dict:
    """contig -> (starts, stops) sorted by (start, stop) (frag/_delfi.py:85-107)."""
    if blacklist_file is None:
        return {}
    by_contig = defaultdict(list)
    with open(blacklist_file) as fh:
        for line in fh:
            parts = line.split()
            if len(parts) >= 3:
                by_contig[parts[0]].append((int(parts[1]), int(parts[2])))
    out = {}
    for contig, regions in by_contig.items():
        regions.sort()
        out[contig] = (np.array([r[0] for r in regions], np.int64), np.array([r[1] for r in regions], np.int64))
    return out


def _resolve_gaps(gap_file):
    """frag/_delfi.py:372-381."""
    if gap_file is None:
        return None
    if isinstance(gap_file, str):
        return GenomeGaps(gap_file)
    if isinstance(gap_file, GenomeGaps):
        return gap_file
    raise TypeError(f"{type(gap_file)} is not accepted type for gap_file")


def _bins_overlapping_gaps(bins: pandas.DataFrame, gaps: GenomeGaps) -> np.ndarray:
    """utils.overlaps (utils/utils.py:346-382) per contig instead of an all-pairs matrix."""
    hit = np.zeros(bins.shape[0], dtype=bool)
    contigs = bins["contig"].to_numpy()
    starts, stops = bins["start"].to_numpy().astype(np.int64), bins["stop"].to_numpy().astype(np.int64)
    for contig in np.unique(gaps.gaps["contig"]):
        rows = np.flatnonzero(contigs == contig)
        g = gaps.gaps[gaps.gaps["contig"] == contig]
        if rows.size:
            hit[rows] = np.any((starts[rows, None] < g["stop"][None]) & (stops[rows, None] > g["start"][None]), axis=1)
    return hit


def delfi_rows(table, ref, contig, starts, stops, blacklist=None, gaps=None, quality_threshold=30, device=None):
    """int64[n, 4] (host) = short, long, num_frags, G+C bases of one contig's bins: one ``ftk_delfi_windows_u64``
    launch; for BAM input every bin is its own read-level fetch (frag/_delfi.py:443, ``_common.per_fetch``)."""
    from ..device import delfi_windows
    from ._common import per_fetch
    dev_ref = ref.device_contig(contig, device) if (ref is not None and contig in ref.chroms) else None

    def run(tab, sel):
        got = delfi_windows(tab.device(contig, device), dev_ref, [starts[k] for k in sel], [stops[k] for k in sel],
                            blacklist=blacklist, gaps=gaps, quality_threshold=quality_threshold)
        return list(got.cpu().numpy())

    rows = per_fetch(table, contig, list(starts), list(stops), run)
    return np.stack(rows) if rows else np.zeros((0, 4), np.int64)


def _contig_counts(table, ref, contig, starts, stops, blacklist, contig_gaps, quality_threshold, count_here=True):
    """Arm labels, live mask and int64[n, 4] counts (short, long, num_frags, G+C bases) of one contig's bins
    (the loop body of ``_delfi_single_window``, frag/_delfi.py:404-511).  ``count_here=False`` (a contig
    another rank owns) skips the CUDA call and leaves the counts at zero for the all-reduce."""
    n = len(starts)
    arms = [contig] * n
    live = np.ones(n, dtype=bool)
    if contig_gaps is not None:
        for i, (s, e) in enumerate(zip(starts, stops)):
            arm = "NOARM" if contig_gaps.in_tcmere(s, e) else contig_gaps.get_arm(s, e)
            arms[i] = arm
            live[i] = arm != "NOARM"
    idx = np.flatnonzero(live)
    counts = np.zeros((n, 4), np.int64)
    if idx.size:
        if contig not in table.contigs:   # pysam.TabixFile.fetch on a contig the index does not know
            raise ValueError(f"could not create iterator for region '{contig}:{starts[idx[0]] + 1}-{stops[idx[0]]}'")
        if count_here:
            gaps = None if contig_gaps is None else (contig_gaps.centromere, contig_gaps.telomeres)
            counts[idx] = delfi_rows(table, ref, contig, [starts[i] for i in idx], [stops[i] for i in idx],
                                     blacklist, gaps, quality_threshold)
    return arms, live, counts


def _rows_from_counts(contig, starts, stops, arms, live, counts):
    """The tuples ``_delfi_single_window`` returns (frag/_delfi.py:498-511)."""
    rows = []
    for i in range(len(starts)):
        s, e = starts[i], stops[i]
        if not live[i]:
            rows.append((contig, s, e, "NOARM", np.nan, np.nan, np.nan, 0))
            continue
        short, long_, num, gc = (int(x) for x in counts[i])
        rows.append((contig, s, e, arms[i], short, long_, gc / (e - s) if num > 0 else np.nan, num))
    return rows


def _contig_windows(table, ref, contig, starts, stops, blacklist, contig_gaps, quality_threshold):
    arms, live, counts = _contig_counts(table, ref, contig, starts, stops, blacklist, contig_gaps, quality_threshold)
    return _rows_from_counts(contig, starts, stops, arms, live, counts)


def delfi_gc_correct(windows: pandas.DataFrame, alpha: float = 0.75, it: int = 8, verbose: bool = False) -> pandas.DataFrame:
    """LOESS GC correction of short / long / num_frags / ratio (frag/_delfi_gc_correct.py:20-94)."""
    try:
        from loess.loess_1d import loess_1d
    except ImportError as e:  # the reference fails at import time without it
        raise ImportError("DELFI GC correction needs the `loess` package (a dependency of the reference); "
                          "install it or call delfi(..., no_gc_correct=True)") from e
    out = windows.copy()
    out.replace([np.inf, -np.inf], np.nan, inplace=True)
    valid = out.dropna()
    grid = np.arange(valid["gc"].min(), valid["gc"].max() + 0.01, 0.01)
    for col in _GC_CORRECT_COLUMNS:
        _, fit, _ = loess_1d(valid["gc"].to_numpy(), valid[col].to_numpy(), xnew=grid, degree=2, frac=alpha)
        out[f"{col}_corrected"] = out[col] - np.interp(out["gc"], grid, fit) + valid[col].median()
    return out


def delfi_merge_bins(hundred_kb_bins: pandas.DataFrame, gc_corrected: bool = True, verbose: bool = False) -> pandas.DataFrame:
    """Merge runs of 50 bins per chromosome arm (frag/_delfi_merge_bins.py:40-92).

    p arms are cut from their first bin, q arms from their last bin (partial runs dropped), arms
    appear in order of first occurrence and q-arm records are emitted in ascending order."""
    columns = hundred_kb_bins.columns[hundred_kb_bins.columns != "index"]
    records = []

    def record(chunk, arm):
        r = [arm[:-1], chunk["start"].min(), chunk["stop"].max(), arm, chunk["short"].sum(), chunk["long"].sum(),
             chunk["gc"].mean(), chunk["num_frags"].sum(), chunk["ratio"].mean()]
        if gc_corrected:
            r += [chunk["short_corrected"].sum(), chunk["long_corrected"].sum(), chunk["num_frags_corrected"].sum(),
                  chunk["ratio_corrected"].mean()]
        return tuple(r)

    for arm in hundred_kb_bins["arm"].unique():
        rows = hundred_kb_bins[hundred_kb_bins["arm"] == arm]
        n = rows.shape[0]
        if "p" in arm:
            firsts = range(0, n - _BINS_PER_WINDOW + 1, _BINS_PER_WINDOW)
        elif "q" in arm:
            # last rows n-1, n-51, ... > 0; a run is kept when it has all 50 rows
            firsts = sorted(last - (_BINS_PER_WINDOW - 1) for last in range(n - 1, 0, -_BINS_PER_WINDOW)
                            if last - (_BINS_PER_WINDOW - 1) >= 0)
        else:
            continue
        records.extend(record(rows.iloc[a: a + _BINS_PER_WINDOW], arm) for a in firsts)
    return pandas.DataFrame(records, columns=columns)


def _write_delfi(final_bins: pandas.DataFrame, output_file: str) -> None:
    """frag/_delfi.py:384-401.  (``.bed.gz`` is written as real gzip here; the reference passes
    ``encoding="gzip"`` to ``to_csv``, which raises LookupError.)"""
    named = final_bins.rename(columns={"contig": "#contig"})
    if output_file.endswith(".bed") or output_file.endswith(".tsv"):
        named.to_csv(output_file, sep="\t", index=False)
    elif output_file.endswith(".csv"):
        final_bins.to_csv(output_file, sep=",", index=False)
    elif output_file.endswith(".bed.gz"):
        with gzip.open(output_file, "wt") as fh:
            named.to_csv(fh, sep="\t", index=False)
    elif output_file == "-":
        for window in final_bins.itertuples(index=False, name=None):
            stdout.write("\t".join(str(field) for field in window) + "\n")
    else:
        raise ValueError("Invalid file type! Only .bed, .bed.gz, and .tsv suffixes allowed.")


def delfi(input_file, chrom_sizes, bins_file, reference_file, blacklist_file=None, gap_file=None, output_file=None,
          no_gc_correct=False, gc_correct=None, remove_nocov=True, merge_bins=True, window_size=5000000,
          quality_threshold=30, workers=1, verbose=False) -> pandas.DataFrame:
    """DELFI features (Cristiano et al., 2019) with the reference's column names (frag/_delfi.py:129-370)."""
    if verbose:
        start_time = time.time()
    contigs = chrom_sizes_to_list(chrom_sizes)
    if gc_correct is None:
        gc_correct = not no_gc_correct
    else:
        warnings.warn("Warning: gc_correct is deprecated and may be removed in future releases. "
                      "Use no_gc_correct instead")
    gaps = _resolve_gaps(gap_file)
    bins = pandas.read_csv(bins_file, names=["contig", "start", "stop"], usecols=[0, 1, 2],
                           dtype={"contig": str, "start": np.int32, "stop": np.int32}, delimiter="\t", comment="#")
    gapless = bins.loc[~_bins_overlapping_gaps(bins, gaps)] if gaps is not None else bins
    if verbose:
        stderr.write(f"{bins.shape[0]} bins read from file, {bins.shape[0] - gapless.shape[0]} removed by gaps.\n")

    blacklist_by_contig = _load_blacklist_indexed(blacklist_file)
    table = as_table(input_file, reference_file)
    ref = open_reference(reference_file)
    windows = []
    bin_contigs = gapless["contig"].to_numpy()
    ctx = dist_context()
    mine = None
    if ctx is not None:   # bins of different contigs are independent (the reference's Pool, frag/_delfi.py:283-294):
        from ..distributed import owned_contigs   # each rank counts the contigs it owns, one all_reduce merges
        mine = set(owned_contigs(table, ctx))
    parts = []
    for contig, _size in contigs:
        sel = gapless.loc[bin_contigs == contig]
        if not sel.shape[0]:
            continue
        starts, stops = sel["start"].tolist(), sel["stop"].tolist()
        arms, live, counts = _contig_counts(table, ref, contig, starts, stops, blacklist_by_contig.get(contig),
                                            gaps.get_contig_gaps(contig) if gaps is not None else None,
                                            quality_threshold, count_here=(mine is None or contig in mine))
        parts.append((contig, starts, stops, arms, live, counts))
    if ctx is not None and parts:
        import torch
        from ..device import require_cuda
        buf = torch.from_numpy(np.concatenate([p[5] for p in parts])).to(require_cuda())
        ctx.all_reduce_sum(buf)
        host, o = buf.cpu().numpy(), 0
        for p in parts:
            p[5][:] = host[o: o + len(p[5])]
            o += len(p[5])
    for p in parts:
        windows.extend(_rows_from_counts(*p))
    if not is_writer(ctx):
        output_file = None
    if verbose:
        stderr.write(f"{len(windows)} windows counted.\n")

    window_df = pandas.DataFrame(windows, columns=_COLUMNS)
    trimmed = window_df.loc[window_df["arm"] != "NOARM", :].copy()
    with np.errstate(divide="ignore", invalid="ignore"):
        trimmed["ratio"] = np.where(trimmed["long"] == 0, np.nan, trimmed["short"] / trimmed["long"])
    if remove_nocov:   # the two hg19 no-coverage bins of Cristiano et al., by position (frag/_delfi.py:332-340)
        pos = np.arange(trimmed.shape[0])
        table_out = trimmed.loc[(pos != 8779) & (pos != 13664)].reset_index()
    else:
        table_out = trimmed
    if gc_correct:
        table_out = delfi_gc_correct(table_out, 0.75, 8, verbose)
    final_bins = delfi_merge_bins(table_out, gc_correct, verbose=verbose) if merge_bins else table_out
    if output_file is not None:
        _write_delfi(final_bins, output_file)
    if verbose:
        stderr.write(f"{sum(w[7] for w in windows)} fragments included.\n")
        stderr.write(f"delfi took {time.time() - start_time} s to complete\n")
    return final_bins
