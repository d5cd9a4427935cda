"""Host-side helpers with the reference's names and semantics.

``frag_generator`` / ``frag_array`` exist for API compatibility (they yield Python
tuples / a structured array, which is host work by definition); the feature
functions in ``finaletoolkit_b200.frag`` never call them - their predicates live
in the CUDA kernels.
"""
from __future__ import annotations

import itertools

import numpy as np

from ..exceptions import InvalidInputError
from ..io.fragments import as_table

__all__ = ["chrom_sizes_to_list", "chrom_sizes_to_dict", "get_intervals", "gen_kmers",
           "reverse_complement", "frag_generator", "frag_array", "frags_in_region", "overlaps"]


# utils/_comparison.py:13-31
def _none_leq(a, b) -> bool:
    return True if a is None or b is None else a <= b


def _none_geq(a, b) -> bool:
    return True if a is None or b is None else a >= b


def _none_eq(a, b) -> bool:
    return True if a is None or b is None else a == b


def chrom_sizes_to_list(chrom_sizes_file) -> list[tuple[str, int]]:
    """utils/utils.py:53-72."""
    out = []
    with open(chrom_sizes_file, "r") as fh:
        for line in fh:
            if line != "\n":
                chrom, size = line.strip().split("\t")
                out.append((chrom, int(size)))
    return out


def chrom_sizes_to_dict(chrom_sizes_file) -> dict[str, int]:
    """utils/utils.py:75-94."""
    return dict(chrom_sizes_to_list(chrom_sizes_file))


def get_intervals(interval_file) -> list[tuple[str, int, int, str]]:
    """utils/utils.py:310-343: BED -> (contig, start, stop, name); tab-split, name default '.'."""
    intervals = []
    with open(interval_file, "r") as bed:
        for line in bed:
            if line.startswith(("#", "track", "browser")) or not line.strip():
                continue
            parts = line.strip().split("\t")
            if len(parts) < 3:
                continue
            intervals.append((parts[0], int(parts[1]), int(parts[2]), parts[3] if len(parts) > 3 else "."))
    return intervals


def overlaps(contigs_1, starts_1, stops_1, contigs_2, starts_2, stops_2) -> np.ndarray:
    """utils/utils.py:346-383: for every interval of set 1, does it overlap (``start_1 < stop_2`` and
    ``stop_1 > start_2``) any interval of set 2 on the same contig?  The reference compares all pairs
    (an n1 x n2 boolean matrix); here set 2 is sorted by start per contig and a running maximum of its stops
    answers a query with one binary search - same booleans, O((n1 + n2) log n2)."""
    c1, s1, e1 = np.asarray(contigs_1), np.asarray(starts_1), np.asarray(stops_1)
    c2, s2, e2 = np.asarray(contigs_2), np.asarray(starts_2), np.asarray(stops_2)
    out = np.zeros(c1.shape[0], dtype=bool)
    for contig in np.unique(c2):
        q = np.flatnonzero(c1 == contig)
        m = c2 == contig
        if not q.size:
            continue
        order = np.argsort(s2[m], kind="stable")
        starts, stops = s2[m][order], e2[m][order]
        # intervals of set 2 starting before the query ends; one of them must end after the query starts
        k = np.searchsorted(starts, e1[q], side="left")
        best = np.maximum.accumulate(stops)
        hit = k > 0
        hit[hit] = best[k[hit] - 1] > s1[q][hit]
        out[q] = hit
    return out


def gen_kmers(k: int, bases: str = "ACGT") -> list[str]:
    """utils/utils.py:388-410."""
    if k < 0:
        raise ValueError("k must be non-negative")
    return ["".join(p) for p in itertools.product(bases, repeat=k)]


_COMP = bytes.maketrans(b"ACGTacgt", b"TGCATGCA")


def reverse_complement(kmer: str) -> str:
    """utils/utils.py:413-437 (N and other symbols preserved; lower case -> upper-case complement)."""
    return kmer.encode("ascii").translate(_COMP)[::-1].decode("ascii")


def _stream_mask(st, sp, mq, quality_threshold, start, stop, min_length, max_length, intersect_policy):
    """The reference's stream predicate, vectorised (io/alignment.py:270-302, utils/_frag_generator.py:21-55,117-123)."""
    if intersect_policy not in ("midpoint", "any"):
        raise InvalidInputError(f"{intersect_policy} is not a valid policy")
    s64, e64 = st.astype(np.int64), sp.astype(np.int64)
    ln = e64 - s64
    m = mq >= quality_threshold
    m &= e64 > (0 if start is None else start)           # tabix overlap
    if stop is not None:
        m &= s64 < stop
    if min_length is not None:
        m &= ln >= min_length
    if max_length is not None:
        m &= ln <= max_length
    if intersect_policy == "midpoint":
        mid = (s64 + e64) // 2
        if start is not None:
            m &= mid >= start
        if stop is not None:
            m &= mid < stop
    return m


def frag_generator(input_file, contig, quality_threshold=30, start=None, stop=None, min_length=None,
                   max_length=None, intersect_policy="midpoint", verbose=False, reference_file=None):
    """utils/_frag_generator.py:58-141: yields (contig, start, stop, mapq, is_forward)."""
    if intersect_policy not in ("midpoint", "any"):
        raise InvalidInputError(f"{intersect_policy} is not a valid policy")
    if contig is None and not (start is None and stop is None):
        if not (start == 0 and stop is None):
            raise InvalidInputError("contig should be specified if start or stop given.")
    table = as_table(input_file, reference_file)
    if contig is not None and table.has_read1(contig):
        table = table.fetched(contig, start, stop)   # BAM: the fetch selects by read 1 (io/alignment.py:245)
    for c in ([contig] if contig is not None else table.contigs):
        st, sp, mq, sd = table.host(c)
        m = _stream_mask(st, sp, mq, quality_threshold, start, stop, min_length, max_length, intersect_policy)
        for i in np.flatnonzero(m).tolist():
            yield (c, int(st[i]), int(sp[i]), int(mq[i]), bool(sd[i]))


def frag_array(input_file, contig, quality_threshold=30, start=None, stop=None, min_length=None,
               max_length=None, intersect_policy="midpoint", verbose=False, reference_file=None):
    """utils/utils.py:186-255: structured [('start','i8'),('stop','i8'),('strand','?')]."""
    table = as_table(input_file, reference_file)
    if contig is not None and table.has_read1(contig):
        table = table.fetched(contig, start, stop)   # BAM: the fetch selects by read 1 (io/alignment.py:245)
    st, sp, mq, sd = table.host(contig)
    m = _stream_mask(st, sp, mq, quality_threshold, start, stop, min_length, max_length, intersect_policy)
    out = np.zeros(int(m.sum()), dtype=[("start", "i8"), ("stop", "i8"), ("strand", "?")])
    out["start"], out["stop"], out["strand"] = st[m], sp[m], sd[m].astype(bool)
    return out


def frags_in_region(frag_array, start: int, stop: int):
    """utils/utils.py:160-183: frag.start < stop and frag.stop >= start."""
    return frag_array[np.logical_and(frag_array["start"] < stop, frag_array["stop"] >= start)]
