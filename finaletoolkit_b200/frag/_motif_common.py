"""Motif containers, MDS and window helpers behind the end- / breakpoint-motif features.

Same public surface and byte-identical text as reference frag/_motif_common.py (the reference's own
golden TSVs round-trip through these classes in tests/), but built around the arrays the CUDA kernels
return - one float64[4^k] vector / one [n_intervals, 4^k] table - instead of dicts of Python scalars."""
from __future__ import annotations

import gzip
from pathlib import Path
from sys import stdin, stdout

import numpy as np

from ..utils import gen_kmers

MIN_QUALITY: int = 20   # Jiang et al. (2020), frag/_motif_common.py:30
_BASES = "ACGT"
_WINDOW_SIZE = 1_000_000


def _normalized_shannon_mds(counts, k: int, miller_madow: bool = False, n=None) -> float:
    """frag/_motif_common.py:38-94: -sum f ln f / ln 4^k, optional Miller-Madow (m-1)/(2N)."""
    num_kmers = 4 ** k
    freq = np.asarray(counts, dtype=np.float64)
    entropy = -np.sum(freq * np.log(freq, out=np.zeros_like(freq, dtype=np.float64), where=(freq != 0)))
    if miller_madow:
        if n is None:
            raise ValueError("n is required when miller_madow is True.")
        if not n > 0:
            return float("nan")
        occupied = int(np.count_nonzero(np.nan_to_num(freq)))
        entropy = entropy + (occupied - 1) / (2 * n)
    return float(entropy / np.log(num_kmers))


def _open_text(path, mode: str):
    """(handle, close?) for a path, ``"-"`` (stdin / stdout) or a ``.gz`` path (read side only)."""
    name = str(path)
    if name == "-":
        return (stdin if "r" in mode else stdout), False
    if "r" in mode and name.endswith("gz"):
        return gzip.open(name, "rt"), True
    return open(name, mode), True


def _require_path(output_file, what: str) -> None:
    if not isinstance(output_file, (str, Path)):
        raise TypeError(what)


class _MotifFreqs:
    """Genome-wide k-mer frequencies (API of reference frag/_motif_common.py:141-262).

    Array-backed: the labels and ONE float64 vector (what the kernel's 4^k counts normalise to);
    ``freq_dict`` and the iteration protocol are views built on demand."""

    def __init__(self, kmer_frequencies, k: int, quality_threshold: int = MIN_QUALITY) -> None:
        pairs = list(kmer_frequencies.items()) if isinstance(kmer_frequencies, dict) else list(kmer_frequencies)
        self._labels = [p[0] for p in pairs]
        self._values = np.array([p[1] for p in pairs], dtype=np.float64)
        self.k = k
        self.quality_threshold = quality_threshold
        if any(len(label) != k for label in self._labels):
            raise ValueError("kmer_frequencies contains a kmer with length not equal to k.")

    @classmethod
    def from_counts(cls, counts, k: int, quality_threshold: int = MIN_QUALITY):
        """From the kernel's int64[4^k] counts (lexicographic ACGT order): frequencies = counts / total."""
        c = np.asarray(counts, dtype=np.float64)
        with np.errstate(invalid="ignore", divide="ignore"):
            f = c / np.sum(c)
        return cls(zip(gen_kmers(k, _BASES), f), k, quality_threshold)

    @property
    def freq_dict(self) -> dict:
        return dict(zip(self._labels, self._values.tolist()))

    def __iter__(self):
        return iter(zip(self._labels, self._values.tolist()))

    def __len__(self) -> int:
        return len(self._labels)

    def __str__(self) -> str:
        return "".join(f"{kmer}: {freq}\n" for kmer, freq in self)

    def kmers(self) -> list:
        return list(self._labels)

    def frequencies(self) -> list:
        return self._values.tolist()

    def freq(self, kmer: str) -> float:
        try:
            return float(self._values[self._labels.index(kmer)])
        except ValueError:
            raise KeyError(kmer) from None

    def to_tsv(self, output_file, sep: str = "\t") -> None:
        _require_path(output_file, "output_file must be a string or path.")
        out, close = _open_text(output_file, "w")
        try:
            out.write("".join(f"{kmer}{sep}{freq}\n" for kmer, freq in self))
        finally:
            if close:
                out.close()

    def motif_diversity_score(self) -> float:
        return _normalized_shannon_mds(self._values, self.k)

    @classmethod
    def from_file(cls, file_path, quality_threshold: int, sep: str = "\t", header: int = 0):
        """Two-column ``kmer<sep>frequency`` text; reading stops at the first row that is not two
        columns; the table must hold exactly 4^k rows of equal-length k-mers (RuntimeError otherwise)."""
        fh, close = _open_text(file_path, "r")
        try:
            rows = fh.read().splitlines(keepends=True)[header:]
        finally:
            if close:
                fh.close()
        # k comes from the row at index `header` of what is left after skipping (reference :230)
        k = len(rows[header].split(sep)[0])
        labels, values = [], []
        for row in rows:
            cells = row.split(sep)
            if len(cells) != 2:
                break
            labels.append(cells[0]); values.append(float(cells[1]))
            if len(cells[0]) != k:
                raise RuntimeError("File contains k-mers of inconsistent length.")
        if len(labels) != 4 ** k:
            raise RuntimeError(f"File contains {len(labels)} {k}-mers instead of the expected {4**k} {k}-mers.")
        return cls(zip(labels, values), k, quality_threshold)


class _MotifsIntervals:
    """Interval-stratified k-mer counts (API of reference frag/_motif_common.py:265-521).

    Array-backed: interval tuples + ONE ``[n_intervals, 4^k]`` table (int64 straight from the kernel,
    float64 when read back from a file) + the column labels; the reference's
    ``[(interval, {kmer: value})]`` list is materialised only when ``.intervals`` is asked for.
    Text is formatted from table rows through ``tolist()``, so ints print as ints and floats as
    floats exactly like the dict-of-Python-scalars version."""

    def __init__(self, intervals, k: int, quality_threshold: int = MIN_QUALITY, total_counts=None,
                 table=None, labels=None) -> None:
        self.k = k
        self.quality_threshold = quality_threshold
        self.total_counts = total_counts
        if table is not None:
            self._keys = list(intervals)
            self._table = np.asarray(table)
            self._labels = list(labels) if labels is not None else gen_kmers(k, _BASES)
            widths_ok = self._table.ndim == 2 and (self._table.shape[1] == 4 ** k or not self._keys)
            self._cache = None
        else:
            intervals = list(intervals)
            self._keys = [iv for iv, _ in intervals]
            widths_ok = all(len(d) == 4 ** k for _, d in intervals)
            self._labels = list(intervals[0][1].keys()) if intervals else gen_kmers(k, _BASES)
            rows = [list(d.values()) for _, d in intervals]
            self._table = np.array(rows) if (rows and widths_ok) else np.zeros((0, 4 ** k), np.int64)
            self._cache = intervals
        if not widths_ok:
            raise ValueError("bins contains results for kmer with length not equal to k.")
        if total_counts is not None and len(total_counts) != len(self._keys):
            raise ValueError("total_counts must have one entry per interval.")

    @property
    def intervals(self) -> list:
        if self._cache is None:
            self._cache = [(key, dict(zip(self._labels, row))) for key, row in zip(self._keys, self._table.tolist())]
        return self._cache

    def __iter__(self):
        return iter(self.intervals)

    def __len__(self) -> int:
        return len(self._keys)

    def __str__(self) -> str:
        return f"{type(self).__name__} over {len(self._keys)} intervals."

    @classmethod
    def from_file(cls, file_path: str, quality_threshold: int, sep: str = ",", header: int = 0):
        """``contig,start,stop,name,count,<4^k values>`` with one header row naming the k-mers."""
        fh, close = _open_text(file_path, "r")
        try:
            rows = fh.read().splitlines(keepends=True)[header:]
        finally:
            if close:
                fh.close()
        labels = rows[0].split(sep)[5:]
        k = round(np.log(len(labels)) / np.log(4))
        assert 4 ** k == len(labels), f"k={k} but should be {len(labels)}."
        keys, totals, table = [], [], np.empty((len(rows) - 1, len(labels)), np.float64)
        for i, row in enumerate(rows[1:]):
            contig, start, stop, name, count, *cells = row.split(sep)
            keys.append((contig, int(start), int(stop), name))
            totals.append(float(count))
            table[i] = [float(c) for c in cells]
        return cls(keys, k, quality_threshold, totals, table=table, labels=labels)

    def freq(self, kmer: str):
        col = self._labels.index(kmer) if kmer in self._labels else None
        if col is None:
            raise KeyError(kmer)
        return dict((*key, v) for key, v in zip(self._keys, self._table[:, col].tolist()))

    def motif_diversity_score(self, miller_madow: bool = False):
        scores = []
        for i, key in enumerate(self._keys):
            counts = self._table[i]
            total = np.sum(counts)
            n = self.total_counts[i] if self.total_counts is not None else total
            with np.errstate(invalid="ignore", divide="ignore"):
                scores.append((key, _normalized_shannon_mds(counts / total, self.k, miller_madow, n)))
        return scores

    def mds_bed(self, output_file, sep: str = "\t", miller_madow: bool = False) -> None:
        with open(output_file, "w") as out:
            out.write("".join(sep.join([key[0], str(key[1]), str(key[2]), key[3], str(score)]) + "\n"
                              for key, score in self.motif_diversity_score(miller_madow)))

    @staticmethod
    def _cell(value, count, calc_freq: bool):
        if not calc_freq:
            return value
        return f"{(value / count):.6f}" if count != 0 else "NaN"

    def _row_values(self, i):
        """Python scalars of row i and their left-to-right sum (the reference sums dict values in order)."""
        row = self._table[i].tolist()
        return row, sum(row)

    def to_tsv(self, output_file, calc_freq: bool = True, sep: str = "\t") -> None:
        _require_path(output_file, "output_file must be a string or path.")
        out, close = _open_text(output_file, "w")
        try:
            out.write(sep.join(["contig", "start", "stop", "name", "count", *gen_kmers(self.k, _BASES)]) + "\n")
            for i, key in enumerate(self._keys):
                row, count = self._row_values(i)
                cells = [str(self._cell(v, count, calc_freq)) for v in row]
                out.write(sep.join([key[0], str(key[1]), str(key[2]), str(key[3]), str(count), *cells]) + "\n")
        finally:
            if close:
                out.close()

    def _one_kmer_track(self, kmer, output_file, calc_freq, sep, with_name) -> None:
        _require_path(output_file, "output_file must be a string.")
        if kmer not in self._labels:
            raise KeyError(kmer)
        col = self._labels.index(kmer)
        out, close = _open_text(output_file, "w")
        try:
            for i, key in enumerate(self._keys):
                row, count = self._row_values(i)
                fields = [key[0], str(key[1]), str(key[2])] + ([key[3]] if with_name else [])
                fields.append(self._cell(row[col], count, calc_freq))
                out.write(sep.join(fields) + "\n")      # a raw (non-string) count raises TypeError like the reference
        finally:
            if close:
                out.close()

    def to_bedgraph(self, kmer, output_file, calc_freq: bool = True, sep: str = "\t") -> None:
        self._one_kmer_track(kmer, output_file, calc_freq, sep, with_name=False)

    def to_bed(self, kmer, output_file, calc_freq: bool = True, sep: str = "\t") -> None:
        self._one_kmer_track(kmer, output_file, calc_freq, sep, with_name=True)


def genome_windows(chrom_length: int):
    """frag/_motif_common.py:527-577: 1 Mb windows + the trailing partial window (quirks kept:
    the last FULL window is never visited when chrom_length is a multiple of 1 Mb)."""
    w = [(s, s + _WINDOW_SIZE) for s in range(0, chrom_length - _WINDOW_SIZE, _WINDOW_SIZE)]
    w.append((chrom_length - chrom_length % _WINDOW_SIZE, chrom_length))
    return w


def interval_motif_rows(table, ref, chrom, starts, stops, k, strand_mode, quality_threshold, breakpoint=False,
                        device=None):
    """int64[len(starts), 4^k] (host): end- or breakpoint-motif counts of the fragments each region's fetch
    yields (frag/_end_motifs.py:115-120, frag/_breakpoint_motifs.py:114-119 - tabix: fragments overlapping the
    region, one launch for all regions; BAM: read 1 overlapping it, ``_common.per_fetch``)."""
    from ..device import end_motif_hist
    from ._common import per_fetch

    def run(tab, sel, widen=0):
        if tab.n_fragments(chrom) == 0:
            return [np.zeros(4 ** k, np.int64) for _ in sel]
        # widen > 0: ``tab`` holds exactly the fetched rows; the wider bounds let every one of them through
        got = end_motif_hist(tab.device(chrom, device), ref.device_contig(chrom, device),
                             [int(starts[j]) - widen for j in sel], [int(stops[j]) + widen for j in sel], k=k,
                             strand_mode=strand_mode, quality_threshold=quality_threshold, breakpoint=breakpoint)
        return list(got.cpu().numpy())

    rows = per_fetch(table, chrom, starts, stops, run, fetch_only=True)
    return np.stack(rows) if rows else np.zeros((0, 4 ** k), np.int64)


def pooled_window_counts(table, ref, chrom, windows, k, strand_mode, quality_threshold, breakpoint=False, total=None,
                         device=None):
    """Adds the k-mer counts of a contig's 1 Mb windows (frag/_motif_common.py:580-610) to the device tensor
    ``total`` int64[1, 4^k] (created when None) and returns it.  Fragment files: one pooled launch.  BAM input:
    every window is its own read-level fetch, so the per-window rows are summed instead."""
    from ..device import end_motif_hist
    ws, we = [a for a, _ in windows], [b for _, b in windows]
    if not table.has_read1(chrom):
        return end_motif_hist(table.device(chrom, device), ref.device_contig(chrom, device), ws, we, k=k,
                              strand_mode=strand_mode, quality_threshold=quality_threshold, pooled=True, counts=total,
                              breakpoint=breakpoint)
    import torch
    rows = interval_motif_rows(table, ref, chrom, ws, we, k, strand_mode, quality_threshold, breakpoint, device)
    add = torch.from_numpy(rows.sum(axis=0, keepdims=True)).to(table.device(chrom, device).device)
    return add if total is None else total.add_(add)


def parse_intervals_arg(intervals):
    """frag/_motif_common.py:613-630: whitespace-split BED or list of tuples."""
    if type(intervals) is str:
        with open(intervals, "r") as interval_file:
            return [(chrom, int(start), int(stop), name[0] if len(name) > 0 else ".")
                    for chrom, start, stop, *name in (line.split() for line in interval_file.readlines())]
    if isinstance(intervals, list):
        return intervals
    raise TypeError("Intervals should be string or list.")


def write_motif_freqs(results, output_file) -> None:
    """frag/_motif_common.py:690-697."""
    if output_file is None:
        return
    if output_file.endswith(".csv"):
        results.to_tsv(output_file, sep=",")
    else:
        results.to_tsv(output_file)
