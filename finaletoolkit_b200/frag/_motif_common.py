"""Motif containers, MDS and drivers - mirror of reference frag/_motif_common.py for end motifs."""
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


class _MotifFreqs:
    """Genome-wide k-mer frequencies (frag/_motif_common.py:141-262)."""

    def __init__(self, kmer_frequencies, k: int, quality_threshold: int = MIN_QUALITY) -> None:
        self.freq_dict = dict(kmer_frequencies)
        self.k = k
        self.quality_threshold = quality_threshold
        if not all(len(kmer) == k for kmer in self.freq_dict):
            raise ValueError("kmer_frequencies contains a kmer with length not equal to k.")

    def __iter__(self):
        return ((kmer, frequency) for kmer, frequency in self.freq_dict.items())

    def __len__(self) -> int:
        return len(self.freq_dict)

    def __str__(self) -> str:
        return "".join(f"{kmer}: {freq}\n" for kmer, freq in self)

    def kmers(self) -> list:
        return list(self.freq_dict.keys())

    def frequencies(self) -> list:
        return list(self.freq_dict.values())

    def freq(self, kmer: str) -> float:
        return self.freq_dict[kmer]

    def to_tsv(self, output_file, sep: str = "\t") -> None:
        if not isinstance(output_file, (str, Path)):
            raise TypeError("output_file must be a string or path.")
        output_is_file = False
        try:
            if str(output_file) == "-":
                output = stdout
            else:
                output_is_file, output = True, open(output_file, "w")
            for kmer, freq in self:
                output.write(f"{kmer}{sep}{freq}\n")
        finally:
            if output_is_file:
                output.close()

    def motif_diversity_score(self) -> float:
        return _normalized_shannon_mds(np.array(self.frequencies()), self.k)

    @classmethod
    def from_file(cls, file_path, quality_threshold: int, sep: str = "\t", header: int = 0):
        file, is_file = None, False
        try:
            if str(file_path).endswith("gz"):
                is_file, file = True, gzip.open(file_path, "rt")
            elif str(file_path) == "-":
                file = stdin
            else:
                is_file, file = True, open(file_path, "rt")
            for _ in range(header):
                file.readline()
            freq_list = []
            lines = file.readlines()
            k = len(lines[header].split(sep)[0])
            for line in lines:
                line_data = line.split(sep)
                if len(line_data) != 2:
                    break
                freq_list.append((line_data[0], float(line_data[1])))
                if k != len(line_data[0]):
                    raise RuntimeError("File contains k-mers of inconsistent length.")
            if (length := len(freq_list)) != 4 ** k:
                raise RuntimeError(f"File contains {length} {k}-mers instead of the expected {4**k} {k}-mers.")
        finally:
            if is_file and file is not None:
                file.close()
        return cls(freq_list, k, quality_threshold)


class _MotifsIntervals:
    """Interval-stratified k-mer counts (frag/_motif_common.py:265-521)."""

    def __init__(self, intervals, k: int, quality_threshold: int = MIN_QUALITY, total_counts=None) -> None:
        self.intervals = intervals
        self.k = k
        self.quality_threshold = quality_threshold
        self.total_counts = total_counts
        if not all(len(freqs) == 4 ** k for _, freqs in intervals):
            raise ValueError("bins contains results for kmer with length not equal to k.")
        if total_counts is not None and len(total_counts) != len(intervals):
            raise ValueError("total_counts must have one entry per interval.")

    def __iter__(self):
        return (interval for interval in self.intervals)

    def __len__(self) -> int:
        return len(self.intervals)

    def __str__(self) -> str:
        return f"{type(self).__name__} over {len(self.intervals)} intervals."

    @classmethod
    def from_file(cls, file_path: str, quality_threshold: int, sep: str = ",", header: int = 0):
        file, is_file = None, False
        try:
            if file_path.endswith("gz"):
                is_file, file = True, gzip.open(file_path, "rt")
            elif file_path == "-":
                file = stdin
            else:
                is_file, file = True, open(file_path)
            for _ in range(header):
                file.readline()
            intervals, total_counts = [], []
            lines = file.readlines()
            _, _, _, _, _, *kmers = lines[0].split(sep)
            k = round(np.log(len(kmers)) / np.log(4))
            assert 4 ** k == len(kmers), f"k={k} but should be {len(kmers)}."
            for line in lines[1:]:
                contig, start, stop, name, count, *freqs = line.split(sep)
                intervals.append(((contig, int(start), int(stop), name), dict(zip(kmers, [float(f) for f in freqs]))))
                total_counts.append(float(count))
        finally:
            if is_file and file is not None:
                file.close()
        return cls(intervals, k, quality_threshold, total_counts)

    def freq(self, kmer: str):
        return dict((*interval, freq[kmer]) for interval, freq in self.intervals)

    def motif_diversity_score(self, miller_madow: bool = False):
        mds = []
        for index, (interval, kmers) in enumerate(self.intervals):
            counts = np.array(list(kmers.values()))
            total = np.sum(counts)
            n = self.total_counts[index] if self.total_counts is not None else total
            with np.errstate(invalid="ignore", divide="ignore"):
                region_mds = _normalized_shannon_mds(counts / total, self.k, miller_madow, n)
            mds.append((interval, region_mds))
        return mds

    def mds_bed(self, output_file, sep: str = "\t", miller_madow: bool = False) -> None:
        with open(output_file, "w") as out:
            for interval, region_mds in self.motif_diversity_score(miller_madow):
                contig, start, stop, name = interval
                out.write(sep.join([contig, str(start), str(stop), name, str(region_mds)]) + "\n")

    def to_tsv(self, output_file, calc_freq: bool = True, sep: str = "\t") -> None:
        if not isinstance(output_file, (str, Path)):
            raise TypeError("output_file must be a string or path.")
        output_is_file = False
        try:
            if str(output_file) == "-":
                output = stdout
            else:
                output_is_file, output = True, open(output_file, "w")
            output.write(sep.join(["contig", "start", "stop", "name", "count", *gen_kmers(self.k, _BASES)]) + "\n")
            for interval, freqs in self.intervals:
                count = sum(freqs.values())
                if calc_freq:
                    values = [f"{(freq / count):.6f}" if count != 0 else "NaN" for freq in freqs.values()]
                else:
                    values = [str(freq) for freq in freqs.values()]
                output.write(sep.join([interval[0], str(interval[1]), str(interval[2]), str(interval[3]),
                                       str(count), *values]) + "\n")
        finally:
            if output_is_file:
                output.close()

    def _to_record(self, kmer, output_file, calc_freq, sep, include_name) -> None:
        if not isinstance(output_file, (str, Path)):
            raise TypeError("output_file must be a string.")
        output_is_file = False
        try:
            if str(output_file) == "-":
                output = stdout
            else:
                output_is_file, output = True, open(output_file, "w")
            for interval, freqs in self.intervals:
                count = sum(freqs.values())
                if calc_freq:
                    value = f"{(freqs[kmer] / count):.6f}" if count != 0 else "NaN"
                else:
                    value = freqs[kmer]
                fields = [interval[0], str(interval[1]), str(interval[2])]
                if include_name:
                    fields.append(interval[3])
                fields.append(value)
                output.write(sep.join(fields) + "\n")
        finally:
            if output_is_file:
                output.close()

    def to_bedgraph(self, kmer, output_file, calc_freq: bool = True, sep: str = "\t") -> None:
        self._to_record(kmer, output_file, calc_freq, sep, include_name=False)

    def to_bed(self, kmer, output_file, calc_freq: bool = True, sep: str = "\t") -> None:
        self._to_record(kmer, output_file, calc_freq, sep, include_name=True)


def genome_windows(chrom_length: int):
    """frag/_motif_common.py:527-577: 1 Mb windows + the trailing partial window (quirks kept:
    the last FULL window is never visited when chrom_length is a multiple of 1 Mb)."""
    w = [(s, s + _WINDOW_SIZE) for s in range(0, chrom_length - _WINDOW_SIZE, _WINDOW_SIZE)]
    w.append((chrom_length - chrom_length % _WINDOW_SIZE, chrom_length))
    return w


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
