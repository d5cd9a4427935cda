"""Host-side fragment decode: file -> per-contig columnar int32 arrays -> HBM.

Replaces the per-interval ``AlignmentWrapper(...).fetch(...)`` text stream of the
reference (io/alignment.py:74-302, re-opened for EVERY interval by
utils/_frag_generator.py:112) with "decode once, keep columns resident": a
``FragmentTable`` holds, per contig and in file order (start-sorted for a
tabix-indexable file), ``start, stop:int32`` and ``mapq, strand:uint8``; its
``device()`` method uploads a contig once and caches the ``ContigFragments``.

Formats (io/alignment.py:143-156, 158-203, 270-302):
  * FinaleDB ``.frag.gz`` 5 columns ``chrom start stop mapq strand``;
  * BED6+ ``.bed.gz`` (more than 5 columns on the first line: mapq = col 5,
    strand = col 6, ``UserWarning`` like the reference);
  * malformed rows are skipped (``ValueError``/``IndexError`` -> continue);
  * a ``.gz``/``.bgz`` path needs its ``.tbi`` sibling (``MissingIndexError``),
    other extensions raise ``UnsupportedFormatError``; a missing file
    ``FileNotFoundError``.
Whole contigs are decoded, so small files are inflated and parsed in one parallel pass; for a
large BGZF file with a real ``.tbi`` the index's per-contig virtual-offset ranges are used to
decode a contig on first use and nothing else (``io/tabix.py``, ``ftk_fragfile_open_slice``).
BAM is decoded natively (``ftk_bamfile_open``: streamed BGZF inflate + record walk with the read
filter and fragment reconstruction of io/alignment.py:60-71, 242-268); CRAM / SAM text need htslib
and go through ``pysam`` when it is importable, ``UnsupportedFormatError`` otherwise.  For BAM input
the reference selects READS overlapping the query region and only then builds fragments
(io/alignment.py:245 ``self._handle.fetch(contig, start, stop)``), so a fragment whose read 1 lies
outside the region is dropped even when the fragment itself reaches into it.  A BAM-derived table
therefore keeps the reference span of read 1 next to every fragment (``FragmentTable.read1``) and
offers the same selection as an index query would make: ``read1_affected`` tells which query regions
hold a fragment that the fragment-level predicate of the kernels would keep and the read-level fetch
drops, ``fetched`` returns the rows an indexed fetch of one region yields (``frag/_common.per_fetch``
routes exactly those regions through it).  The mapq filter is NOT applied at load time - it is a
kernel predicate.
"""
from __future__ import annotations

import gzip
import os
import warnings
from os import PathLike
from typing import Dict, Iterable, Tuple

import numpy as np

from ..exceptions import MissingIndexError, UnsupportedFormatError

__all__ = ["FragmentTable", "load_fragments", "as_table"]

Columns = Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]


class FragmentTable:
    """Per-contig fragment columns (host) with a device cache.

    ``lazy`` = (contig names in file order, loader): contigs are decoded by ``loader(contig)`` the
    first time they are asked for (tabix-indexed files); ``columns`` then holds the loaded ones."""

    def __init__(self, columns: Dict[str, Columns], source: str | None = None, is_sam: bool = False,
                 contig_lengths: Dict[str, int] | None = None, lazy=None, read1: Dict[str, tuple] | None = None):
        self.source = source
        self.is_sam = is_sam
        self.contig_lengths = contig_lengths  # BAM/CRAM header only
        self.columns: Dict[str, Columns] = {}
        # BAM/CRAM only: {contig: (r1_start, r1_end) int32}, the reference span of read 1 of every fragment
        # (rows as in ``columns``), clipped to the fragment
        self.read1: Dict[str, tuple] = {}
        self.read1_raw: Dict[str, tuple] = {}      # the spans as aligned (a dovetailed read runs past its template)
        self._read1_in = read1 or {}
        self._lazy_names, self._loader = (list(lazy[0]), lazy[1]) if lazy is not None else (None, None)
        self._weights = lazy[2] if (lazy is not None and len(lazy) > 2) else None
        for contig, cols in columns.items():
            self._put(contig, cols)
        self._device: dict = {}

    def _put(self, contig: str, cols: Columns) -> None:
        st, sp, mq, sd = cols
        st = np.ascontiguousarray(st, dtype=np.int32)
        sp = np.ascontiguousarray(sp, dtype=np.int32)
        mq = np.ascontiguousarray(mq, dtype=np.uint8)
        sd = np.ascontiguousarray(sd, dtype=np.uint8)
        r1 = self._read1_in.pop(contig, None)
        if r1 is not None:
            r1 = (np.asarray(r1[0], dtype=np.int64), np.asarray(r1[1], dtype=np.int64))
            if r1[0].shape != st.shape or r1[1].shape != st.shape:
                raise ValueError(f"read-1 columns of contig {contig!r} do not match its fragment columns")
        if st.size and not np.all(st[1:] >= st[:-1]):
            order = np.argsort(st, kind="stable")  # keeps file order among equal starts
            st, sp, mq, sd = st[order], sp[order], mq[order], sd[order]
            if r1 is not None:
                r1 = (r1[0][order], r1[1][order])
        self.columns[contig] = (st, sp, mq, sd)
        if r1 is not None:
            # only the part of the read inside its own fragment can decide a query: a fragment that does not
            # overlap the region fails the intersect policy whatever its read does (utils/_frag_generator.py:36-50)
            lo = np.maximum(r1[0], st).astype(np.int32)
            hi = np.maximum(np.minimum(r1[1], sp), lo + 1).astype(np.int32)
            self.read1[contig] = (lo, hi)
            # the motif counters take every fetched fragment WITHOUT a fragment-level test
            # (frag/_end_motifs.py:115-120), so for them the whole read decides: ``fetch_only`` below
            self.read1_raw[contig] = (r1[0].astype(np.int32), np.maximum(r1[1], r1[0] + 1).astype(np.int32))

    def _ensure(self, contig) -> None:
        if self._loader is not None and contig not in self.columns and contig in self._lazy_names:
            self._put(contig, self._loader(contig))

    @property
    def contigs(self):
        return list(self._lazy_names) if self._lazy_names is not None else list(self.columns.keys())

    def n_fragments(self, contig=None) -> int:
        if contig is None:
            for c in self.contigs:
                self._ensure(c)
            return sum(c[0].size for c in self.columns.values())
        self._ensure(contig)
        return self.columns[contig][0].size if contig in self.columns else 0

    def shard_weight(self, contig: str) -> int:
        """Load estimate of one contig for LPT sharding, identical on every rank: the fragment count for
        an in-memory table, the compressed byte span of the contig's index range for a lazy tabix-backed
        one (no rank has to inflate a contig it does not own)."""
        if self._weights is None:
            return self.n_fragments(contig)
        return int(self._weights.get(contig, 0))

    def host(self, contig: str) -> Columns:
        self._ensure(contig)
        if contig not in self.columns:
            z = np.zeros(0, np.int32)
            return z, z.copy(), np.zeros(0, np.uint8), np.zeros(0, np.uint8)
        return self.columns[contig]

    # ---- read-level fetch of BAM input (io/alignment.py:242-247)
    def has_read1(self, contig=None) -> bool:
        """True when the rows of ``contig`` (any contig for None) carry read-1 spans, i.e. when a region query
        has to select by READ overlap like ``pysam.AlignmentFile.fetch`` does."""
        if contig is None:
            return bool(self.read1) or bool(self._read1_in)
        self._ensure(contig)
        return contig in self.read1

    def _read1_sorted(self, contig: str):
        cache = self.__dict__.setdefault("_r1_sorted", {})
        if contig not in cache:
            st, sp = self.columns[contig][:2]
            lo, hi = self.read1[contig]
            raw_lo, raw_hi = self.read1_raw[contig]
            max_len = int((sp.astype(np.int64) - st).max()) if st.size else 0
            reach = max(max_len, int((raw_hi.astype(np.int64) - raw_lo).max()) if st.size else 0)
            cache[contig] = (np.sort(sp), np.sort(lo), np.sort(hi), max_len, np.sort(raw_lo), np.sort(raw_hi), reach)
        return cache[contig]

    def fetch_reach(self, contig: str) -> int:
        """How far a fetched row can lie from its query region: the longest fragment or read-1 span of the contig."""
        return self._read1_sorted(contig)[6] if self.has_read1(contig) else 0

    @staticmethod
    def _bounds(starts, stops):
        big = np.int64(1) << 40
        lo = np.array([-big if v is None else int(v) for v in starts], dtype=np.int64)
        hi = np.array([big if v is None else int(v) for v in stops], dtype=np.int64)
        return lo, hi

    def read1_affected(self, contig: str, starts, stops, fetch_only: bool = False) -> np.ndarray:
        """bool per query region ``[starts[i], stops[i])`` (None = unbounded): can the read-level fetch of the
        reference (io/alignment.py:245) and a fragment-level predicate disagree on it?  Default (the fetch is
        followed by an intersect policy): the region holds a fragment that overlaps it while its read 1 does not.
        ``fetch_only`` (the motif counters take every fetched fragment as it comes): additionally a read that
        overlaps the region while its own - shorter, dovetailed - template does not.  Two binary searches per
        region and coordinate column over sorted copies."""
        lo, hi = self._bounds(starts, stops)
        if not self.has_read1(contig) or not len(lo):
            return np.zeros(len(lo), dtype=bool)
        st = self.columns[contig][0]
        sp_sorted, r1s_sorted, r1e_sorted, _, raw_s, raw_e, _ = self._read1_sorted(contig)
        # reads and fragments have positive length, so  #overlapping = #(begin < hi) - #(end <= lo)
        frag = np.searchsorted(st, hi, side="left") - np.searchsorted(sp_sorted, lo, side="right")
        both = np.searchsorted(r1s_sorted, hi, side="left") - np.searchsorted(r1e_sorted, lo, side="right")
        differ = frag != both         # ``read1`` = read AND fragment: its overlaps are a subset of either's
        if fetch_only:
            read = np.searchsorted(raw_s, hi, side="left") - np.searchsorted(raw_e, lo, side="right")
            differ |= read != both
        return differ & (hi > lo)

    def _subtable(self, contig: str, keep: np.ndarray) -> "FragmentTable":
        st, sp, mq, sd = self.columns[contig]
        sub = FragmentTable({contig: (st[keep], sp[keep], mq[keep], sd[keep])}, source=self.source,
                            is_sam=self.is_sam, contig_lengths=self.contig_lengths)
        sub._transient = True     # one query's rows: uploaded through staging, not page-locked in place
        return sub

    def fetched(self, contig: str, start=None, stop=None, fetch_only: bool = False) -> "FragmentTable":
        """The rows an indexed BAM fetch of ``contig:[start, stop)`` yields (read 1 overlaps the region,
        io/alignment.py:245; None = unbounded), as a one-contig table without read-1 columns: the fragment-level
        predicates of the kernels then finish the job exactly like ``frag_generator`` does after the fetch.
        ``fetch_only``: select by the whole read (see ``read1_affected``); a caller that applies no intersect
        policy then queries the sub-table with bounds widened by ``fetch_reach`` so that every row passes."""
        if not self.has_read1(contig):
            return self
        (lo,), (hi,) = self._bounds([start], [stop])
        st = self.columns[contig][0]
        r1s, r1e = (self.read1_raw if fetch_only else self.read1)[contig]
        reach = self._read1_sorted(contig)[6]
        a = int(np.searchsorted(st, lo - reach, side="left"))     # rows further out cannot reach the region
        b = int(np.searchsorted(st, hi + reach, side="left"))
        return self._subtable(contig, np.flatnonzero((r1s[a:b] < hi) & (r1e[a:b] > lo)) + a)

    def fetch_groups(self, contig: str, starts, stops, fetch_only: bool = False) -> list:
        """Partition query regions into groups that can share ONE fetched table: inside a group any two
        regions lie at least one maximal fragment length apart, so a row fetched for one of them is entirely
        outside every other and cannot pass the other's fragment predicate (midpoint or overlap).  A tiling of
        5-kb windows needs two groups (even / odd) instead of one query per window.  ``fetch_only``: the
        regions are going to be queried widened by ``fetch_reach`` on either side, so they keep twice that
        distance plus one.  Returns index arrays (ascending) into ``starts``; first-fit over the regions sorted
        by start."""
        lo, hi = self._bounds(starts, stops)
        if not len(lo):
            return []
        gap = 0
        if self.has_read1(contig):
            gap = 2 * self._read1_sorted(contig)[6] + 1 if fetch_only else self._read1_sorted(contig)[3]
        group_end: list = []           # the right-most region end of every group so far
        members: list = []
        for j in np.argsort(lo, kind="stable").tolist():
            for g, end in enumerate(group_end):
                if lo[j] >= end + gap:
                    group_end[g] = max(end, int(hi[j])); members[g].append(j)
                    break
            else:
                group_end.append(int(hi[j])); members.append([j])
        return [np.sort(np.asarray(m, dtype=np.int64)) for m in members]

    def fetched_union(self, contig: str, starts, stops, fetch_only: bool = False) -> "FragmentTable":
        """The rows whose read 1 overlaps ANY of the regions of one ``fetch_groups`` group (pairwise disjoint,
        at least a fragment length apart), as a one-contig table without read-1 columns - the rows of
        ``fetched`` for every region of the group in one table."""
        if not self.has_read1(contig):
            return self
        lo, hi = self._bounds(starts, stops)
        order = np.argsort(lo, kind="stable")
        lo, hi = lo[order], hi[order]
        if np.any(lo[1:] < hi[:-1]):
            raise ValueError("fetched_union needs pairwise disjoint regions (one group of fetch_groups)")
        r1s, r1e = (self.read1_raw if fetch_only else self.read1)[contig]
        # the regions are disjoint and sorted: the only candidate of a read is the first region ending after its start
        k = np.searchsorted(hi, r1s, side="right")
        ok = k < len(lo)
        return self._subtable(contig, np.flatnonzero(ok & (lo[np.minimum(k, len(lo) - 1)] < r1e)))

    def pinned(self, contig: str):
        """``(start, stop, mapq, strand)`` of one contig as CPU torch tensors that SHARE the host columns' memory,
        page-locked in place with ``cudaHostRegister`` (once per contig, released with the table) - what
        the streamed pipeline (``pipeline.StreamedContig``) copies from without an intermediate staging
        buffer.  ``None`` when the pages cannot be locked (the caller falls back to the resident upload)."""
        import weakref
        import torch
        cache = self.__dict__.setdefault("_pinned", {})
        if contig in cache:
            return cache[contig]
        st, sp, mq, sd = self.host(contig)
        tensors, ok = [], True
        rt = torch.cuda.cudart()
        for a in (st, sp, mq, sd):
            t = torch.from_numpy(a)
            if a.nbytes and not t.is_pinned():
                rc = rt.cudaHostRegister(t.data_ptr(), a.nbytes, 0)
                if int(rc) != 0:
                    ok = False
                    break
                weakref.finalize(self, _unregister, t.data_ptr())
            tensors.append(t)
        cache[contig] = tuple(tensors) if ok and all(t.is_pinned() or t.numel() == 0 for t in tensors) else None
        return cache[contig]

    def device(self, contig: str, device=None):
        """``ContigFragments`` of one contig in HBM (uploaded once, then cached)."""
        from ..device import ContigFragments, require_cuda
        dev = require_cuda(device)
        key = (contig, str(dev))
        if key not in self._device:
            cols = self.pinned(contig) if (self.n_fragments(contig) and not getattr(self, "_transient", False)) else None
            if cols is not None:
                # DMA straight out of the page-locked host columns (80 M fragments: 56 ms to lock the pages +
                # 21 ms of copies, against 0.8 s through freshly allocated staging buffers); the longest
                # fragment is reduced on the device
                import torch
                with torch.cuda.device(dev):
                    d = [c.to(dev, non_blocking=True) for c in cols]
                self._device[key] = ContigFragments(d[0], d[1], d[2], d[3], device=dev, contig=contig)
            else:
                st, sp, mq, sd = self.host(contig)
                self._device[key] = ContigFragments(st, sp, mq, sd, device=dev, contig=contig)
        return self._device[key]


def _unregister(ptr: int) -> None:
    try:
        import torch
        torch.cuda.cudart().cudaHostUnregister(ptr)
    except Exception:  # noqa: BLE001 - interpreter shutdown / context already gone
        pass


_CACHE: dict = {}


def _check_path(path: str) -> None:
    if not os.path.exists(path):
        raise FileNotFoundError(f"Alignment file not found: {path}")
    lower = path.lower()
    if lower.endswith((".bam", ".cram", ".sam")):
        if lower.endswith(".bam") and not (os.path.exists(path + ".bai") or os.path.exists(path[:-4] + ".bai")):
            raise MissingIndexError(f"BAM file {path} missing index (.bai)")
        if lower.endswith(".cram") and not (os.path.exists(path + ".crai") or os.path.exists(path[:-5] + ".crai")):
            raise MissingIndexError(f"CRAM file {path} missing index (.crai)")
    elif lower.endswith((".gz", ".bgz")):
        if not os.path.exists(path + ".tbi"):
            raise MissingIndexError(f"Compressed file {path} missing tabix index (.tbi)")
    else:
        raise UnsupportedFormatError(f"Unsupported file format: {path}")


def _parse_text_rows(lines: Iterable[str]) -> Dict[str, Columns]:
    """Tab-separated fragment rows -> columns per contig (file order)."""
    bed_format = None
    acc: Dict[str, list] = {}
    for line in lines:
        if not line or line[0] == "#" or not line.strip():
            continue
        f = line.rstrip("\n").split("\t")
        if bed_format is None:
            bed_format = len(f) > 5
            if bed_format:
                warnings.warn(
                    "input_file does not follow Fragmentation file format accepted by FinaleToolkit. "
                    "Attempting to read as a BED6 file.", UserWarning)
        try:
            st, sp = int(f[1]), int(f[2])
            if bed_format:
                mq, fw = int(f[4]), "+" in f[5]
            else:
                mq, fw = int(f[3]), "+" in f[4]
        except (ValueError, IndexError):
            continue  # io/alignment.py:301-302
        a = acc.get(f[0])
        if a is None:
            a = acc[f[0]] = ([], [], [], [])
        a[0].append(st); a[1].append(sp); a[2].append(min(max(mq, 0), 255)); a[3].append(1 if fw else 0)
    return {c: (np.array(a[0], np.int64), np.array(a[1], np.int64), np.array(a[2], np.uint8), np.array(a[3], np.uint8))
            for c, a in acc.items()}


def _decode_native(path: str, threads: int = 0) -> Dict[str, Columns] | None:
    """Multi-threaded C++ decoder in libftk_b200.so (BGZF blocks inflated + parsed in parallel).

    Returns None when the library is unavailable so the pure-Python parsers can take over
    (decode is host-side I/O, not the compute path)."""
    import ctypes
    try:
        from .._lib import lib
        L = lib()
    except Exception:  # noqa: BLE001
        return None
    err = ctypes.c_int32(0)
    h = L.ftk_fragfile_open(path.encode(), int(threads), ctypes.byref(err))
    if not h:
        return None
    try:
        if L.ftk_fragfile_is_bed6(h):
            warnings.warn(
                "input_file does not follow Fragmentation file format accepted by FinaleToolkit. "
                "Attempting to read as a BED6 file.", UserWarning)
        return _handle_columns(L, h)
    finally:
        L.ftk_fragfile_close(h)


def _handle_read1(L, h) -> Dict[str, tuple] | None:
    """Read-1 reference spans of a BAM-derived handle, per contig (rows as in ``_handle_columns``)."""
    import ctypes
    out: Dict[str, tuple] = {}
    p32 = ctypes.POINTER(ctypes.c_int32)
    for i in range(L.ftk_fragfile_n_contigs(h)):
        n = L.ftk_fragfile_contig_count(h, i)
        lo = np.empty(n, np.int32); hi = np.empty(n, np.int32)
        if L.ftk_fragfile_copy_read1(h, i, lo.ctypes.data_as(p32), hi.ctypes.data_as(p32)) != 0:
            return None
        out[L.ftk_fragfile_contig_name(h, i).decode()] = (lo, hi)
    return out


def _handle_columns(L, h) -> Dict[str, Columns] | None:
    import ctypes
    cols: Dict[str, Columns] = {}
    for i in range(L.ftk_fragfile_n_contigs(h)):
        n = L.ftk_fragfile_contig_count(h, i)
        st = np.empty(n, np.int32); sp = np.empty(n, np.int32)
        mq = np.empty(n, np.uint8); sd = np.empty(n, np.uint8)
        p32, p8 = ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_uint8)
        if L.ftk_fragfile_copy(h, i, st.ctypes.data_as(p32), sp.ctypes.data_as(p32), mq.ctypes.data_as(p8),
                               sd.ctypes.data_as(p8)) != 0:
            return None
        cols[L.ftk_fragfile_contig_name(h, i).decode()] = (st, sp, mq, sd)
    return cols


_LAZY_MIN_BYTES = 32 << 20   # below this a whole-file decode is cheaper than thinking about it


def _lazy_table(path: str) -> "FragmentTable | None":
    """A ``FragmentTable`` that decodes contigs on demand through the file's tabix index, or None when
    that does not apply (small file, placeholder / foreign .tbi, not BGZF, library unavailable)."""
    import ctypes
    if os.path.getsize(path) < int(os.environ.get("FTK_LAZY_MIN_BYTES", _LAZY_MIN_BYTES)):
        return None
    from .tabix import read_tbi
    index = read_tbi(path + ".tbi")
    if index is None or not index.ranges or (index.col_seq, index.col_beg, index.col_end) != (1, 2, 3):
        return None
    try:
        from .._lib import lib
        L = lib()
        with open(path, "rb") as fh:
            head = fh.read(18)
        if len(head) < 18 or head[:4] != b"\x1f\x8b\x08\x04" or head[12:14] != b"BC":
            return None
        with gzip.open(path, "rt") as fh:            # column layout from the first data line
            first = next((ln for ln in fh if ln.strip() and not ln.startswith(index.meta)), "")
    except Exception:  # noqa: BLE001
        return None
    bed6 = len(first.rstrip("\n").split("\t")) > 5
    if bed6:
        warnings.warn(
            "input_file does not follow Fragmentation file format accepted by FinaleToolkit. "
            "Attempting to read as a BED6 file.", UserWarning)

    def load(contig: str) -> Columns:
        cb, ub, ce, ue = index.ranges[contig]
        err = ctypes.c_int32(0)
        h = L.ftk_fragfile_open_slice(path.encode(), cb, ub, ce, ue, int(bed6), 0, ctypes.byref(err))
        if not h:
            raise OSError(f"{path}: cannot decode the tabix range of contig {contig!r} (error {err.value})")
        try:
            cols = _handle_columns(L, h)
        finally:
            L.ftk_fragfile_close(h)
        z = np.zeros(0, np.int32)
        return (cols or {}).get(contig, (z, z.copy(), np.zeros(0, np.uint8), np.zeros(0, np.uint8)))

    names = [n for n in index.names if n in index.ranges]
    weights = {n: max(index.ranges[n][2] - index.ranges[n][0], 1) for n in names}
    return FragmentTable({}, source=path, lazy=(names, load, weights))


def _parse_text_fast(path: str) -> Dict[str, Columns] | None:
    """pandas C parser for well-formed files; None -> caller falls back to the row parser."""
    try:
        import pandas as pd
        with gzip.open(path, "rt") as fh:
            first = fh.readline()
        n_cols = len(first.rstrip("\n").split("\t"))
        if n_cols < 5 or first.startswith("#"):
            return None
        bed = n_cols > 5
        use = [0, 1, 2, 4, 5] if bed else [0, 1, 2, 3, 4]
        df = pd.read_csv(path, sep="\t", header=None, usecols=use, compression="gzip", comment="#",
                         dtype={0: str, use[1]: np.int64, use[2]: np.int64, use[3]: np.int64, use[4]: str},
                         engine="c", na_filter=False)
        if bed:
            warnings.warn(
                "input_file does not follow Fragmentation file format accepted by FinaleToolkit. "
                "Attempting to read as a BED6 file.", UserWarning)
        cols = {}
        contig = df[0].to_numpy()
        st = df[use[1]].to_numpy(); sp = df[use[2]].to_numpy()
        mq = np.clip(df[use[3]].to_numpy(), 0, 255).astype(np.uint8)
        fw = df[use[4]].str.contains("+", regex=False).to_numpy().astype(np.uint8)
        # contigs in order of first appearance, rows in file order
        uniq, first_idx = np.unique(contig, return_index=True)
        for c in uniq[np.argsort(first_idx)]:
            m = contig == c
            cols[str(c)] = (st[m], sp[m], mq[m], fw[m])
        return cols
    except Exception:  # noqa: BLE001 - any irregularity: use the tolerant row parser
        return None


def _load_bam_native(path: str) -> "FragmentTable | None":
    """BAM through the C ABI's streamed decoder (``ftk_bamfile_open``): same read filter and fragment
    reconstruction as the reference (io/alignment.py:60-71,242-268), no htslib.  None = unavailable."""
    import ctypes
    try:
        from .._lib import lib
        L = lib()
    except Exception:  # noqa: BLE001
        return None
    err = ctypes.c_int32(0)
    h = L.ftk_bamfile_open(path.encode(), 0, ctypes.byref(err))
    if not h:
        return None
    try:
        cols = _handle_columns(L, L.ftk_bamfile_fragments(h))
        read1 = _handle_read1(L, L.ftk_bamfile_fragments(h)) if cols is not None else None
        lengths = {L.ftk_bamfile_ref_name(h, i).decode(): int(L.ftk_bamfile_ref_length(h, i))
                   for i in range(L.ftk_bamfile_n_refs(h))}
    finally:
        L.ftk_bamfile_close(h)
    if cols is None or read1 is None:
        return None
    return FragmentTable(cols, source=path, is_sam=True, contig_lengths=lengths, read1=read1)


def _load_sam(path: str, reference_file=None) -> FragmentTable:
    if path.lower().endswith(".bam"):
        tab = _load_bam_native(path)
        if tab is not None:
            return tab
    try:
        import pysam
    except ImportError as e:
        raise UnsupportedFormatError(
            f"{path}: BAM/CRAM/SAM decoding needs htslib (pysam), which is not installed; "
            "convert to a tabix-indexed .frag.gz") from e
    acc: Dict[str, list] = {}
    with pysam.AlignmentFile(path, "r", reference_filename=str(reference_file) if reference_file else None) as bam:
        lengths = dict(zip(bam.references, bam.lengths))
        for read in bam.fetch():
            # io/alignment.py:60-71 with quality_threshold deferred to the kernels, :248 read1 only
            if (read.is_unmapped or read.is_secondary or not read.is_paired or read.mate_is_unmapped
                    or read.is_duplicate or read.is_qcfail or read.is_supplementary
                    or not read.is_proper_pair or read.is_read2):
                continue
            tlen = read.template_length
            if tlen > 0:
                st, sp = read.reference_start, read.reference_start + tlen
            elif tlen < 0:
                st, sp = read.reference_end + tlen, read.reference_end
            else:
                continue
            a = acc.setdefault(read.reference_name, ([], [], [], [], [], []))
            a[0].append(st); a[1].append(sp); a[2].append(read.mapping_quality); a[3].append(1 if read.is_forward else 0)
            r_end = read.reference_end
            a[4].append(read.reference_start)
            a[5].append(r_end if (r_end is not None and r_end > read.reference_start) else read.reference_start + 1)
    cols = {c: (np.array(a[0], np.int64), np.array(a[1], np.int64), np.array(a[2], np.uint8), np.array(a[3], np.uint8))
            for c, a in acc.items()}
    read1 = {c: (np.array(a[4], np.int64), np.array(a[5], np.int64)) for c, a in acc.items()}
    return FragmentTable(cols, source=path, is_sam=True, contig_lengths=lengths, read1=read1)


def load_fragments(input_file, reference_file=None) -> FragmentTable:
    """Decode ``input_file`` once into a cached ``FragmentTable``."""
    path = str(input_file)
    _check_path(path)
    st = os.stat(path)
    key = (os.path.abspath(path), st.st_mtime_ns, st.st_size)
    tab = _CACHE.get(key)
    if tab is not None:
        return tab
    if path.lower().endswith((".bam", ".cram", ".sam")):
        tab = _load_sam(path, reference_file)
    else:
        tab = _lazy_table(path)
        if tab is None:
            cols = _decode_native(path)
            if cols is None:
                cols = _parse_text_fast(path)
            if cols is None:
                with gzip.open(path, "rt") as fh:
                    cols = _parse_text_rows(fh)
            tab = FragmentTable(cols, source=path)
    if len(_CACHE) > 8:
        _CACHE.pop(next(iter(_CACHE)))
    _CACHE[key] = tab
    return tab


def as_table(input_file, reference_file=None) -> FragmentTable:
    """Accept what the reference's ``FragFile`` accepts (utils/typing.py:13) plus a ``FragmentTable``.

    A path (str / PathLike) is decoded (and cached); an open ``pysam.TabixFile`` /
    ``pysam.AlignmentFile`` is re-opened by its filename; a ``FragmentTable`` or a dict
    ``{contig: (start, stop, mapq, strand)}`` is used as is.
    """
    if isinstance(input_file, FragmentTable):
        return input_file
    if isinstance(input_file, dict):
        return FragmentTable(input_file)
    if isinstance(input_file, (str, PathLike)):
        return load_fragments(input_file, reference_file)
    name = getattr(input_file, "filename", None)
    if name is not None:  # open pysam handle
        if isinstance(name, bytes):
            name = name.decode()
        return load_fragments(name, reference_file)
    raise TypeError(f"unsupported input_file type {type(input_file)!r}")
