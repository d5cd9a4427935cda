"""CPU oracle for the FinaleToolkit hot path (TEST INFRASTRUCTURE).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  The product package
``finaletoolkit_b200`` never does; its compute path is CUDA only.

Thin numpy/ctypes front-end over ``oracle/ftk_oracle.c`` (the literal C
restatement of the reference's loops) plus small pure-Python restatements of
the reference's host-side arithmetic (statistics in dict order, binning, MDS,
site/window construction).  Every function cites the reference file:line it
follows (relative to /root/reference/src/finaletoolkit).

Parity pin: ``tests/test_oracle_golden.py`` checks every function here against
``tests/golden/*`` - outputs of the UNMODIFIED reference produced in the build
container by ``oracle/make_golden.py``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_char_p, c_double, c_int, c_int32, c_int64, c_uint8

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libftk_oracle.so")
_SRC = os.path.join(_HERE, "ftk_oracle.c")
NONE = -(2 ** 63)  # ORC_NONE: Python None (unbounded)


def build(force: bool = False) -> str:
    """Compile the C restatement with gcc (seconds)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(
            ["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", _SO, _SRC, "-lm"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        i32p, u8p, i64p, dp = POINTER(c_int32), POINTER(c_uint8), POINTER(c_int64), POINTER(c_double)
        L = _lib
        L.orc_wps_interval.restype = c_int64
        L.orc_wps_interval.argtypes = [i32p, i32p, u8p, c_int64, c_int64, c_int64, c_int64, c_int64,
                                       c_int64, c_int64, c_int64, c_int64, i64p]
        L.orc_wps_intervals.restype = None
        L.orc_wps_intervals.argtypes = [i32p, i32p, u8p, c_int64, c_int64, i64p, i64p, i64p, c_int64, c_int64,
                                        c_int64, c_int64, c_int64, c_int64, i64p, c_int]
        L.orc_single_coverage.restype = c_int64
        L.orc_single_coverage.argtypes = [i32p, i32p, u8p, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int, c_int64]
        L.orc_interval_coverage.restype = None
        L.orc_interval_coverage.argtypes = [i32p, i32p, u8p, c_int64, c_int64, i64p, i64p, c_int64, c_int64, c_int64, c_int, c_int64, i64p, c_int]
        L.orc_length_dist.restype = c_int64
        L.orc_length_dist.argtypes = [i32p, i32p, u8p, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int, c_int64, i64p, i64p, c_int64]
        L.orc_frag_lengths.restype = c_int64
        L.orc_frag_lengths.argtypes = [i32p, i32p, u8p, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int, c_int64, i32p, c_int64]
        L.orc_region_end_motifs.restype = c_int
        L.orc_region_end_motifs.argtypes = [i32p, i32p, u8p, u8p, c_int64, c_int64, c_int64, c_int64, c_char_p, c_int64, c_int, c_int, c_int64, i64p]
        L.orc_delfi_window.restype = None
        L.orc_delfi_window.argtypes = [i32p, i32p, u8p, c_int64, c_int64, c_int64, c_int64, i64p, i64p, c_int64,
                                       c_int, c_int64, c_int64, i64p, c_int64, c_int64, c_char_p, c_int64, i64p]
        L.orc_region_breakpoint_motifs.restype = c_int
        L.orc_region_breakpoint_motifs.argtypes = L.orc_region_end_motifs.argtypes
        L.orc_cleavage_interval.restype = c_int64
        L.orc_cleavage_interval.argtypes = [i32p, i32p, u8p, u8p, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64,
                                            c_int64, c_int64, c_int64, c_int64, dp]
        L.orc_local_filter.restype = c_int64
        L.orc_local_filter.argtypes = [dp, c_int64, c_int64, c_int, dp]
    return _lib


def _p(a, t):
    return a.ctypes.data_as(POINTER(t))


def _n(v):
    return NONE if v is None else int(v)


class Frags:
    """Start-sorted fragments of one contig as C-contiguous columns."""

    def __init__(self, start, stop, mapq, strand=None):
        self.start = np.ascontiguousarray(start, dtype=np.int32)
        self.stop = np.ascontiguousarray(stop, dtype=np.int32)
        self.mapq = np.ascontiguousarray(mapq, dtype=np.uint8)
        self.strand = (np.ones_like(self.mapq) if strand is None
                       else np.ascontiguousarray(strand, dtype=np.uint8))
        self.n = int(self.start.shape[0])
        self.max_len = int((self.stop.astype(np.int64) - self.start).max()) if self.n else 0
        self.max_len = max(self.max_len, 0)

    def _args(self):
        return (_p(self.start, c_int32), _p(self.stop, c_int32), _p(self.mapq, c_uint8), self.n, self.max_len)


# ------------------------------------------------------------------- WPS
def wps_interval(fr: Frags, start, stop, chrom_size, window_size=120, min_length=120,
                 max_length=180, quality_threshold=30) -> np.ndarray:
    """frag/_wps.py:56-205 -> int64[stop-start]."""
    n = max(int(stop) - int(start), 0)
    out = np.zeros(n, dtype=np.int64)
    lib().orc_wps_interval(*fr._args(), int(start), int(stop), int(chrom_size), int(window_size),
                           int(min_length), int(max_length), int(quality_threshold), _p(out, c_int64))
    return out


def wps_intervals(fr: Frags, ivl_start, ivl_stop, chrom_size, window_size=120, min_length=120,
                  max_length=180, quality_threshold=30, threads=1) -> tuple[np.ndarray, np.ndarray]:
    """frag/_multi_wps.py:196-198 : the Pool over intervals. Returns (out int64, offsets)."""
    s = np.ascontiguousarray(ivl_start, dtype=np.int64)
    e = np.ascontiguousarray(ivl_stop, dtype=np.int64)
    ln = np.maximum(e - s, 0)
    off = np.zeros(len(s) + 1, dtype=np.int64)
    np.cumsum(ln, out=off[1:])
    out = np.zeros(int(off[-1]), dtype=np.int64)
    lib().orc_wps_intervals(*fr._args(), _p(s, c_int64), _p(e, c_int64), _p(off, c_int64), len(s),
                            int(chrom_size), int(window_size), int(min_length), int(max_length),
                            int(quality_threshold), _p(out, c_int64), int(threads))
    return out, off


def read_sites(lines, interval_size, chrom_sizes: dict):
    """frag/_multi_wps.py:240-297 (_read_sites) + :152-160 (header-order sort).

    ``lines`` is an iterable of BED lines; ``chrom_sizes`` an ordered dict.
    Returns a list of (contig, start, stop).  Raises ValueError like the reference.
    """
    left = round(-interval_size / 2)
    right = round(interval_size / 2)
    assert right - left == interval_size
    out = []
    prev_contig, prev_start, prev_stop = None, 0, 0
    for line in lines:
        c = line.split()
        contig = c[0].strip()
        if int(c[1]) > int(c[2]):
            raise ValueError("start after stop")
        if contig not in chrom_sizes:
            continue
        mid = (int(c[1]) + int(c[2])) // 2
        start = max(0, mid + int(left))
        stop = min(mid + int(right), chrom_sizes[contig])
        if contig == prev_contig and start < prev_stop:
            prev_stop = start
        if prev_contig is not None and prev_stop > prev_start:
            out.append((prev_contig, prev_start, prev_stop))
        prev_contig, prev_start, prev_stop = contig, start, stop
    if prev_stop > prev_start:
        out.append((prev_contig, prev_start, prev_stop))
    order = {c: i for i, c in enumerate(chrom_sizes)}
    out.sort(key=lambda t: (order.get(t[0], len(order)), t[1]))  # stable, like sorted()
    return out


# ------------------------------------------------------ coverage / lengths
_POLICY = {"midpoint": 0, "any": 1}


def single_coverage(fr: Frags, start=0, stop=None, min_length=None, max_length=None,
                    intersect_policy="midpoint", quality_threshold=30) -> int:
    """frag/_coverage.py:117-130."""
    return int(lib().orc_single_coverage(*fr._args(), _n(start), _n(stop), _n(min_length), _n(max_length),
                                         _POLICY[intersect_policy], int(quality_threshold)))


def interval_coverage(fr: Frags, ivl_start, ivl_stop, min_length=None, max_length=None,
                      intersect_policy="midpoint", quality_threshold=30, threads=1) -> np.ndarray:
    """frag/_coverage.py:244-248."""
    s = np.ascontiguousarray(ivl_start, dtype=np.int64)
    e = np.ascontiguousarray(ivl_stop, dtype=np.int64)
    out = np.zeros(len(s), dtype=np.int64)
    lib().orc_interval_coverage(*fr._args(), _p(s, c_int64), _p(e, c_int64), len(s), _n(min_length), _n(max_length),
                                _POLICY[intersect_policy], int(quality_threshold), _p(out, c_int64), int(threads))
    return out


def length_dist(fr: Frags, start=None, stop=None, min_length=None, max_length=None,
                intersect_policy="midpoint", quality_threshold=30) -> dict:
    """frag/_frag_length.py:147-153 : dict length->count in first-seen order."""
    cap = 1 << 16
    while True:
        k = np.zeros(cap, np.int64)
        v = np.zeros(cap, np.int64)
        r = lib().orc_length_dist(*fr._args(), _n(start), _n(stop), _n(min_length), _n(max_length),
                                  _POLICY[intersect_policy], int(quality_threshold), _p(k, c_int64), _p(v, c_int64), cap)
        if r >= 0:
            return dict(zip(k[:r].tolist(), v[:r].tolist()))
        cap *= 16


def frag_lengths(fr: Frags, start=None, stop=None, intersect_policy="midpoint", quality_threshold=30) -> np.ndarray:
    """frag/_frag_length.py:290-308 (length filter hard-wired 0..1e9)."""
    out = np.zeros(max(fr.n, 1), np.int32)
    r = lib().orc_frag_lengths(*fr._args(), _n(start), _n(stop), 0, 1000000000, _POLICY[intersect_policy],
                               int(quality_threshold), _p(out, c_int32), fr.n)
    assert r >= 0
    return out[:r].copy()


def merge_dists(dists) -> dict:
    """Genome-wide stream = contigs in file order; merge keeps first-seen order."""
    out: dict = {}
    for d in dists:
        for k, v in d.items():
            out[k] = out.get(k, 0) + v
    return out


def find_median(d: dict) -> float:
    """frag/_frag_length.py:156-172 (_find_median), quirks included."""
    val = np.array(list(d.keys()))
    freq = np.array(list(d.values()))
    order = np.argsort(val)
    val, freq = val[order], freq[order]
    cdf = np.cumsum(freq)
    total = cdf[-1]
    if total % 2 == 1:
        return float(val[np.searchsorted(cdf, total // 2)])
    idx = np.searchsorted(cdf, [total // 2, total // 2 + 1])
    return float(np.mean(val[idx]))


def length_stats(d: dict, short_reads: int):
    """frag/_frag_length.py:204-238 : (mean, median, stdev, min, max, count, frac_short) or -1s."""
    total = sum(d.values())
    if total == 0:
        return (-1, -1, -1, -1, -1, -1, -1)
    mean = sum(v * c for v, c in d.items()) / total
    median = find_median(d)
    var = sum(c * ((v - mean) ** 2) for v, c in d.items()) / total
    n_short = sum(c for v, c in d.items() if v <= short_reads)
    return (mean, median, var ** 0.5, min(d), max(d), total, n_short / total)


def length_bins(d: dict, bin_size: int):
    """frag/_frag_length.py:458-469 : (bins ndarray, counts list)."""
    lo, hi = min(d), max(d)
    n_bins = (hi - lo) // bin_size
    bins = np.arange(lo, hi + bin_size, bin_size)
    counts = np.zeros(n_bins + 1, dtype=np.int64)
    for v, c in d.items():
        counts[(v - lo) // bin_size] += c
    return bins, counts.tolist()


# --------------------------------------------------------------- end motifs
def region_end_motifs(fr: Frags, seq_ascii: bytes, start, stop, k=4, both_strands=True,
                      negative_strand=False, quality_threshold=20) -> np.ndarray:
    """frag/_end_motifs.py:51-187 -> int64[4**k] in gen_kmers order. RuntimeError like the reference."""
    if both_strands and negative_strand:
        raise ValueError("Cannot have both both_strands and negative_strand.")
    mode = 0 if both_strands else (2 if negative_strand else 1)
    counts = np.zeros(4 ** k, np.int64)
    err = lib().orc_region_end_motifs(*fr._args()[:3], _p(fr.strand, c_uint8), fr.n, fr.max_len, int(start), int(stop),
                                      seq_ascii, len(seq_ascii), int(k), mode, int(quality_threshold), _p(counts, c_int64))
    if err:
        raise RuntimeError("Error querying sequence (reverse k-mer out of bounds)")
    return counts


def region_breakpoint_motifs(fr: Frags, seq_ascii: bytes, start, stop, k=6, both_strands=True,
                             negative_strand=False, quality_threshold=30) -> np.ndarray:
    """frag/_breakpoint_motifs.py:53-196 -> int64[4**k] in gen_kmers order."""
    if both_strands and negative_strand:
        raise ValueError("Cannot have both both_strands and negative_strand.")
    mode = 0 if both_strands else (2 if negative_strand else 1)
    counts = np.zeros(4 ** k, np.int64)
    lib().orc_region_breakpoint_motifs(*fr._args()[:3], _p(fr.strand, c_uint8), fr.n, fr.max_len, int(start),
                                       int(stop), seq_ascii, len(seq_ascii), int(k), mode, int(quality_threshold),
                                       _p(counts, c_int64))
    return counts


def genome_windows(chrom_len: int, window: int = 1_000_000):
    """frag/_motif_common.py:527-577 (_genome_window_args) for one contig."""
    w = [(s, s + window) for s in range(0, chrom_len - window, window)]
    w.append((chrom_len - chrom_len % window, chrom_len))
    return w


def mds(freq, k: int, miller_madow=False, n=None) -> float:
    """frag/_motif_common.py:38-94 (_normalized_shannon_mds)."""
    freq = np.asarray(freq, dtype=np.float64)
    ent = -np.sum(freq * np.log(freq, out=np.zeros_like(freq), where=(freq != 0)))
    if miller_madow:
        if not n > 0:
            return float("nan")
        ent = ent + (int(np.count_nonzero(np.nan_to_num(freq))) - 1) / (2 * n)
    return float(ent / np.log(4 ** k))


# ------------------------------------------------------------ DELFI windows
def contig_arm(contig: str, gaps, start: int, stop: int) -> str:
    """ContigGaps.get_arm, genome/gaps.py:250-268.  ``gaps`` = (centromere, telomeres, has_short_arm)."""
    (c0, c1), _, has_short_arm = gaps
    if stop < start:
        raise ValueError("start must be less than stop")
    if stop < c0:
        return "NOARM" if has_short_arm else f"{contig.replace('chr', '')}p"
    if start > c1:
        return f"{contig.replace('chr', '')}q"
    return "NOARM"


def in_tcmere(gaps, start: int, stop: int) -> bool:
    """ContigGaps.in_tcmere, genome/gaps.py:226-248 (``all`` over telomeres, like the reference)."""
    (c0, c1), telomeres, _ = gaps
    in_c = stop > c0 and start < c1
    in_t = bool(telomeres) and all(stop > t0 and start < t1 for t0, t1 in telomeres)
    return in_c or in_t


def delfi_window(fr: Frags, seq_ascii, contig: str, start: int, stop: int, blacklist=None, gaps=None,
                 quality_threshold=30):
    """frag/_delfi.py:404-511 -> (contig, start, stop, arm, short, long, gc, num_frags).

    ``blacklist`` = (starts, stops) of the contig sorted by (start, stop) (frag/_delfi.py:85-107);
    ``gaps`` = (centromere (start, stop), [telomere (start, stop)...], has_short_arm) or None."""
    if gaps is not None:
        if in_tcmere(gaps, start, stop) or contig_arm(contig, gaps, start, stop) == "NOARM":
            return (contig, start, stop, "NOARM", np.nan, np.nan, np.nan, 0)
        arm = contig_arm(contig, gaps, start, stop)
    else:
        arm = contig
    short, long_, num, gc = delfi_counts(fr, seq_ascii, start, stop, blacklist,
                                         None if gaps is None else gaps[:2], quality_threshold)
    gc_content = gc / (stop - start) if num > 0 else np.nan
    return (contig, start, stop, arm, short, long_, gc_content, num)


def delfi_counts(fr: Frags, seq_ascii, start: int, stop: int, blacklist=None, gaps=None, quality_threshold=30):
    """(short, long, num_frags, G+C bases) of one bin - the loop of frag/_delfi.py:437-484.

    ``gaps`` = (centromere (start, stop), [telomere (start, stop)...]) or None."""
    bs = np.ascontiguousarray([] if blacklist is None else blacklist[0], np.int64)
    be = np.ascontiguousarray([] if blacklist is None else blacklist[1], np.int64)
    telo = np.ascontiguousarray([] if gaps is None else [x for t in gaps[1] for x in t], np.int64)
    out = np.zeros(4, np.int64)
    lib().orc_delfi_window(*fr._args()[:3], fr.n, fr.max_len, int(start), int(stop), _p(bs, c_int64), _p(be, c_int64),
                           len(bs), int(gaps is not None), 0 if gaps is None else int(gaps[0][0]),
                           0 if gaps is None else int(gaps[0][1]), _p(telo, c_int64), len(telo) // 2,
                           int(quality_threshold), seq_ascii, 0 if seq_ascii is None else len(seq_ascii), _p(out, c_int64))
    return tuple(int(x) for x in out)


# ---------------------------------------------------------- BAM -> fragments
def bam_fragments(bam_bytes: bytes, with_read1: bool = False):
    """io/alignment.py:60-71 (_read_is_low_quality without the mapq test) + :242-268 (_fetch_sam) on the
    records of an uncompressed-by-gzip BAM: returns (references [(name, length)], rows) with rows =
    (contig, start, stop, mapq, is_forward) in file order.  ``with_read1`` appends the reference span
    ``(pos, bam_endpos)`` of the read itself - what an indexed fetch tests against the region (htslib:
    ``pos < stop and bam_endpos > start``, bam_endpos = pos + max(reference bases of the CIGAR, 1))."""
    import gzip
    import struct
    raw = gzip.decompress(bam_bytes)
    assert raw[:4] == b"BAM\x01"
    (l_text,) = struct.unpack_from("<i", raw, 4); off = 8 + l_text
    (n_ref,) = struct.unpack_from("<i", raw, off); off += 4
    refs = []
    for _ in range(n_ref):
        (ln,) = struct.unpack_from("<i", raw, off)
        refs.append((raw[off + 4: off + 4 + ln - 1].decode(), struct.unpack_from("<i", raw, off + 4 + ln)[0])); off += 8 + ln
    rows = []
    while off < len(raw):
        (bs,) = struct.unpack_from("<i", raw, off)
        ref_id, pos, l_rn, mapq, _b, n_cig, flag, _ls, _nr, _np, tlen = struct.unpack_from("<iiBBHHHiiii", raw, off + 4)
        cig = struct.unpack_from(f"<{n_cig}I", raw, off + 36 + l_rn)
        off += 4 + bs
        unmapped, secondary, paired, mate_unmapped = flag & 0x4, flag & 0x100, flag & 0x1, flag & 0x8
        dup, qcfail, supp, proper, read2 = flag & 0x400, flag & 0x200, flag & 0x800, flag & 0x2, flag & 0x80
        if unmapped or secondary or not paired or mate_unmapped or dup or qcfail or supp or not proper:   # :60-71
            continue
        if read2:                                                                                          # :248
            continue
        ref_end = pos + sum(c >> 4 for c in cig if (c & 15) in (0, 2, 3, 7, 8))
        if tlen > 0:                                                                                       # :252-260
            f_start, f_stop = pos, pos + tlen
        elif tlen < 0:
            if not cig:
                continue   # pysam: reference_end is None
            f_start, f_stop = ref_end + tlen, ref_end
        else:
            continue
        row = (refs[ref_id][0], f_start, f_stop, mapq, not (flag & 0x10))
        rows.append(row + (pos, ref_end if ref_end > pos else pos + 1) if with_read1 else row)
    return refs, rows


def bam_fetch(bam_bytes: bytes, contig=None, start=None, stop=None):
    """``AlignmentWrapper._fetch_sam`` (io/alignment.py:242-268) for one region: the fragments of the READS an
    indexed ``fetch(contig, start, stop)`` returns, file order, mapq filter not applied."""
    _, rows = bam_fragments(bam_bytes, with_read1=True)
    out = []
    for c, f_start, f_stop, mapq, fwd, r_start, r_end in rows:
        if contig is not None:
            if c != contig or (stop is not None and r_start >= stop) or (start is not None and r_end <= start):
                continue
        out.append((c, f_start, f_stop, mapq, fwd))
    return out


def frag_stream(rows, quality_threshold=30, start=None, stop=None, min_length=None, max_length=None,
                intersect_policy="midpoint"):
    """utils/_frag_generator.py:112-130 over already fetched rows (contig, start, stop, mapq, is_forward):
    mapq (io/alignment.py:60-71), the inclusive length window and the intersect policy."""
    out = []
    for c, s, e, q, fwd in rows:
        ln = e - s
        if q < quality_threshold or (min_length is not None and ln < min_length) or (max_length is not None and ln > max_length):
            continue
        if intersect_policy == "midpoint":
            mid = (s + e) // 2
            ok = (start is None or mid >= start) and (stop is None or mid < stop)
        else:
            ok = (start is None or e > start) and (stop is None or s < stop)
        if ok:
            out.append((c, s, e, q, fwd))
    return out


# ------------------------------------------------------------------ agg_bw
def agg_bw_core(signals, strands, median_window_size=1, mean=False):
    """utils/_agg_bw.py:84-126: ``signals[i]`` = pyBigWig ``values`` of interval i (float32, NaN where
    uncovered) or None when the query raised; ``strands[i]`` = BED column 6.  The first interval
    defines the size.  Returns the aggregate exactly as the reference builds it (running sum)."""
    first = next(len(v) for v in signals if v is not None) if signals[0] is None else len(signals[0])
    interval_size = first - median_window_size
    agg = np.zeros(interval_size, dtype=np.int64)
    added = 0
    for v, strand in zip(signals, strands):
        if v is None:
            continue
        values = np.nan_to_num(np.array([float(x) for x in v]), nan=0)
        trimmed = values[median_window_size // 2: -median_window_size // 2]
        if trimmed.shape[0] != interval_size:
            continue
        if strand == "+":
            agg = agg + trimmed; added += 1
        elif strand == "-":
            agg = agg + np.flip(trimmed); added += 1
    if mean:
        agg = agg / added
    return agg


# --------------------------------------------------------- cleavage profile
def cleavage_profile(fr: Frags, chrom_size, start, stop, left=0, right=0, min_length=None, max_length=None,
                     quality_threshold=30):
    """frag/_cleavage_profile.py:93-224 -> (positions int64, proportion float64)."""
    a = max(int(start) - int(left), 0)
    b = min(int(stop) + int(right), int(chrom_size))
    out = np.zeros(max(b - a, 0), np.float64)
    lib().orc_cleavage_interval(*fr._args()[:3], _p(fr.strand, c_uint8), fr.n, fr.max_len, int(start), int(stop),
                                int(left), int(right), int(chrom_size), _n(min_length), _n(max_length),
                                int(quality_threshold), _p(out, c_double))
    return np.arange(a, max(b, a), dtype=np.int64), out


def cleavage_intervals(lines, left, right, chrom_sizes: dict):
    """frag/_cleavage_profile.py:411-449 (_read_intervals): expand, clamp, merge overlapping neighbours."""
    contigs, starts, stops = [], [], []
    prev_contig, prev_start, prev_stop = None, 0, 0
    for line in lines:
        c = line.split()
        contig = c[0].strip()
        start, stop = int(c[1]), int(c[2])
        if contig not in chrom_sizes:
            continue
        start = max(0, start - left)
        stop = min(stop + right, chrom_sizes[contig])
        if prev_contig == contig and start < prev_stop:
            prev_stop = max(prev_stop, stop)
        else:
            contigs.append(prev_contig); starts.append(prev_start); stops.append(prev_stop)
            prev_contig, prev_start, prev_stop = contig, start, stop
    contigs.append(prev_contig); starts.append(prev_start); stops.append(prev_stop)
    return list(zip(contigs[1:], starts[1:], stops[1:]))


# --------------------------------------------------------------- adjust_wps
def local_filter(data, w: int, use_mean=False) -> np.ndarray:
    """frag/_adjust_wps.py:25-45 via the C restatement (sort-based median)."""
    x = np.ascontiguousarray(data, dtype=np.float64)
    out = np.zeros(max(len(x) - w, 0), np.float64)
    lib().orc_local_filter(_p(x, c_double), len(x), int(w), int(bool(use_mean)), _p(out, c_double))
    return out


def adjust_core(data, w=1000, use_mean=False, savgol=True, sg_window=21, sg_deg=2) -> np.ndarray:
    """frag/_adjust_wps.py:131-140 : the same numpy + scipy calls the reference makes.

    Third-party arithmetic: numpy.median/mean over sliding_window_view and
    scipy.signal.savgol_filter (reference pins scipy 1.15.3/1.17.0; this image
    has 1.18.1 - same algorithm).
    """
    from numpy.lib.stride_tricks import sliding_window_view
    from scipy.signal import savgol_filter
    x = np.asarray(data, dtype=np.float64)
    nw = len(x) - w
    if nw <= 0:
        run = np.array([], dtype=np.float64)
    else:
        win = sliding_window_view(x, w)[:nw]
        run = np.mean(win, axis=1) if use_mean else np.median(win, axis=1)
    adj = x[w // 2: -(w // 2)] - run
    return savgol_filter(adj, sg_window, sg_deg) if savgol else adj


def adjust_sites(lines, interval_size, median_window_size):
    """frag/_adjust_wps.py:219-263 : BED -> merged (contig, start, stop) intervals."""
    left = round(-interval_size / 2)
    right = round(interval_size / 2)
    assert right - left == interval_size
    dec = median_window_size // 2
    out = []
    for line in lines:
        c = line.split("\t")
        contig = c[0].strip()
        mid = (int(c[1]) + int(c[2])) // 2
        start = max(0, mid + int(left))
        stop = mid + int(right)
        if out and out[-1][0] == contig and out[-1][2] - dec > start + dec:
            start = out[-1][1]
            out.pop()
        out.append((contig, int(start), int(stop)))
    return out
