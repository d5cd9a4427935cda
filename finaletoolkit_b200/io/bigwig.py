"""bigWig reader / writer (numpy + zlib) with the slice of the pyBigWig API the hot path uses.

pyBigWig (libBigWig) is not available in this image, and ``.bw`` is the primary
WPS format of the reference CLI: ``multi_wps`` writes raw WPS with
``addEntries(chrom, start, values=..., span=1, step=1)`` (frag/_multi_wps.py:300-325)
and ``adjust_wps`` reads it back with ``intervals(contig, start, stop)``
(frag/_adjust_wps.py:80-105) and writes ``addEntries(chroms, starts, ends=, values=)``
(frag/_adjust_wps.py:275-291).  This module implements the UCSC bigWig container
(header, chromosome B+ tree, zlib-compressed fixedStep / bedGraph sections, R-tree
index, total summary; zoom levels are optional in the format and omitted) so files
interoperate with pyBigWig / UCSC tools.  Values are stored as float32 like libBigWig.

Sections are independent zlib streams; the writer queues them and deflates a batch at a
time on all host threads through the C ABI (``ftk_zlib_compress_batch``), and the reader
can inflate every section a list of queries touches in one call (``prefetch`` ->
``ftk_zlib_uncompress_batch``).  Both need ``libftk_b200.so`` (host code, no GPU).

    bw = open(path, "w"); bw.addHeader([(chrom, size), ...]); bw.addEntries(...); bw.close()
    bw = open(path);      bw.chroms(); bw.intervals(chrom, start, end); bw.values(chrom, s, e)
"""
from __future__ import annotations

import builtins
import ctypes
import os
import struct
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .._lib import check, lib

__all__ = ["open", "BigWigReader", "BigWigWriter"]

_BW_MAGIC = 0x888FFC26
_CHROM_TREE_MAGIC = 0x78CA8C91
_RTREE_MAGIC = 0x2468ACE0
_ITEMS_PER_SECTION = 16384   # <= 65535 (itemCount is u16)
_RTREE_BLOCK = 256
_FLUSH_BYTES = 32 << 20      # queued raw section bytes before a batch is deflated
_SECTION_HDR = 24


def _ptr(a: np.ndarray, ctype):
    return a.ctypes.data_as(ctypes.POINTER(ctype))


def _deflate_sections(raws: list[bytes], level: int | None = None) -> list[bytes]:
    """zlib-compress independent sections on all host threads (ftk_zlib_compress_batch).  Level 6 like
    libBigWig unless ``FTK_BIGWIG_LEVEL`` says otherwise (1 is ~3x faster for ~20 % larger files)."""
    if level is None:
        level = int(os.environ.get("FTK_BIGWIG_LEVEL", 4))   # 3x the speed of level 6 for a ~14 % larger file
    n = len(raws)
    lens = np.fromiter((len(r) for r in raws), dtype=np.int64, count=n)
    in_off = np.zeros(n + 1, dtype=np.int64); np.cumsum(lens, out=in_off[1:])
    out_off = np.zeros(n + 1, dtype=np.int64); np.cumsum(lens + lens // 1000 + 64, out=out_off[1:])
    src = np.frombuffer(b"".join(raws), dtype=np.uint8)
    dst = np.empty(int(out_off[-1]), dtype=np.uint8)
    sizes = np.zeros(n, dtype=np.int64)
    check(lib().ftk_zlib_compress_batch(_ptr(src, ctypes.c_uint8), _ptr(in_off, ctypes.c_int64), n, level, 0,
                                        _ptr(dst, ctypes.c_uint8), _ptr(out_off, ctypes.c_int64),
                                        _ptr(sizes, ctypes.c_int64)), "ftk_zlib_compress_batch")
    return [dst[int(o): int(o) + int(z)].tobytes() for o, z in zip(out_off[:-1], sizes)]


def _inflate_sections(buf: bytes, blocks: list[tuple[int, int]], max_uncomp: int) -> list[np.ndarray]:
    """Inflate independent sections on all host threads (ftk_zlib_uncompress_batch); returns uint8 views
    into one batch buffer (no per-section copy)."""
    n = len(blocks)
    in_off = np.fromiter((b[0] for b in blocks), dtype=np.int64, count=n)
    in_size = np.fromiter((b[1] for b in blocks), dtype=np.int64, count=n)
    if n and (int(in_off.min()) < 0 or int(in_size.min()) < 0 or int((in_off + in_size).max()) > len(buf)):
        raise RuntimeError("corrupt bigWig index: a data section lies outside the file")   # never hand C a bad pointer
    out_off = np.arange(n + 1, dtype=np.int64) * int(max_uncomp)
    src = np.frombuffer(buf, dtype=np.uint8)
    dst = np.empty(int(out_off[-1]), dtype=np.uint8)
    sizes = np.zeros(n, dtype=np.int64)
    check(lib().ftk_zlib_uncompress_batch(_ptr(src, ctypes.c_uint8), _ptr(in_off, ctypes.c_int64),
                                          _ptr(in_size, ctypes.c_int64), n, 0, _ptr(dst, ctypes.c_uint8),
                                          _ptr(out_off, ctypes.c_int64), _ptr(sizes, ctypes.c_int64)),
          "ftk_zlib_uncompress_batch")
    return [dst[int(o): int(o) + int(z)] for o, z in zip(out_off[:-1], sizes)]


class BigWigWriter:
    def __init__(self, path: str):
        self.path = str(path)
        self._fh = builtins.open(self.path, "wb")
        self._chroms: list[tuple[str, int]] | None = None
        self._ids: dict[str, int] = {}
        self._sections: list[tuple[int, int, int, int, int]] = []  # chromId, start, end, offset, size
        self._pending: list[tuple[int, int, int, bytes]] = []      # chromId, start, end, raw section
        self._pending_bytes = 0
        self._last = (-1, -1)   # (chromId, end) of the previous entry: entries must be sorted
        self._max_uncomp = 0
        self._n_cov, self._min, self._max, self._sum, self._sumsq = 0, np.inf, -np.inf, 0.0, 0.0
        self._closed = False
        # batches are deflated and written by one background thread (ctypes releases the GIL), in
        # submission order, while the caller keeps queueing the next intervals
        self._worker = ThreadPoolExecutor(max_workers=1)
        self._inflight: list = []

    # -- pyBigWig-compatible surface ------------------------------------
    def addHeader(self, header, maxZooms: int = 10) -> None:
        self._chroms = [(str(c), int(n)) for c, n in header]
        self._ids = {c: i for i, (c, _) in enumerate(self._chroms)}
        n = len(self._chroms)
        key_size = max((len(c.encode()) for c, _ in self._chroms), default=1)
        # header (64) + total summary (40) + chrom tree; data follows
        self._fh.write(b"\0" * 64)
        self._summary_off = self._fh.tell()
        self._fh.write(b"\0" * 40)
        self._chrom_tree_off = self._fh.tell()
        self._fh.write(struct.pack("<IIIIQQ", _CHROM_TREE_MAGIC, max(n, 1), key_size, 8, n, 0))
        self._fh.write(struct.pack("<BBH", 1, 0, n))
        for c, i in sorted(self._ids.items(), key=lambda kv: kv[0].encode()):   # keys sorted bytewise
            self._fh.write(c.encode().ljust(key_size, b"\0") + struct.pack("<II", i, self._chroms[i][1]))
        self._data_off = self._fh.tell()
        self._fh.write(struct.pack("<Q", 0))  # section count, patched on close

    def addEntries(self, chroms, starts, ends=None, values=None, span=None, step=None, validate=True) -> None:
        if self._chroms is None:
            raise RuntimeError("The bigWig file handle is not opened for writing or has no header.")
        vals = np.ascontiguousarray(np.asarray(values, dtype=np.float64).astype(np.float32))
        if vals.size == 0:
            return
        if isinstance(chroms, str):
            cid = self._chrom_id(chroms)
            if span is not None and step is not None and np.ndim(starts) == 0:
                self._add_fixed(cid, int(starts), int(span), int(step), vals)
                return
            st = np.asarray(starts, dtype=np.int64)
            en = st + int(span) if ends is None else np.asarray(ends, dtype=np.int64)
            self._add_bedgraph(cid, st, en, vals)
            return
        chroms = list(chroms)
        st = np.asarray(starts, dtype=np.int64)
        en = np.asarray(ends, dtype=np.int64)
        if not (len(chroms) == st.size == en.size == vals.size):
            raise RuntimeError("chroms, starts, ends and values must have the same length")
        if chroms.count(chroms[0]) == len(chroms):       # the usual call: one contig repeated
            self._add_bedgraph(self._chrom_id(chroms[0]), st, en, vals)
            return
        # runs of identical contigs
        i = 0
        while i < len(chroms):
            j = i
            while j < len(chroms) and chroms[j] == chroms[i]:
                j += 1
            self._add_bedgraph(self._chrom_id(chroms[i]), st[i:j], en[i:j], vals[i:j])
            i = j

    def close(self) -> None:
        if self._closed:
            return
        self._closed = True
        fh = self._fh
        if self._chroms is None:
            self._worker.shutdown()
            fh.close()
            return
        try:
            self._flush(wait=True)
        finally:
            self._worker.shutdown()
        index_off = fh.tell()
        self._write_rtree(index_off)
        end = fh.tell()
        fh.seek(self._data_off)
        fh.write(struct.pack("<Q", len(self._sections)))
        fh.seek(self._summary_off)
        if self._n_cov:
            fh.write(struct.pack("<Qdddd", self._n_cov, self._min, self._max, self._sum, self._sumsq))
        fh.seek(0)
        fh.write(struct.pack("<IHHQQQHHQQIQ", _BW_MAGIC, 4, 0, self._chrom_tree_off, self._data_off, index_off,
                             0, 0, 0, self._summary_off, self._max_uncomp, 0))
        fh.seek(end)
        fh.write(struct.pack("<I", _BW_MAGIC))
        fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc) -> None:
        self.close()

    # -- internals ------------------------------------------------------
    def _chrom_id(self, chrom: str) -> int:
        try:
            return self._ids[chrom]
        except KeyError:
            raise RuntimeError(f"Invalid chromosome {chrom!r}: not in the bigWig header") from None

    def _check_order(self, cid: int, start: int, end: int) -> None:
        size = self._chroms[cid][1]
        if start < 0 or end > size or end <= start:
            raise RuntimeError("The entries you tried to add are out of bounds or empty")
        if (cid, start) < self._last:
            raise RuntimeError("The entries you tried to add are out of order")

    def _stats(self, vals: np.ndarray, bases: np.ndarray | int) -> None:
        v = vals.astype(np.float64)
        self._min = min(self._min, float(vals.min())); self._max = max(self._max, float(vals.max()))
        if np.ndim(bases) == 0:      # fixed span: two reductions, no weighted temporaries
            self._n_cov += int(bases) * v.size
            self._sum += float(bases) * float(v.sum()); self._sumsq += float(bases) * float(np.dot(v, v))
            return
        b = np.asarray(bases, dtype=np.float64)
        self._n_cov += int(b.sum())
        self._sum += float(np.dot(v, b)); self._sumsq += float(np.dot(v * v, b))

    def _emit(self, cid, start, end, step, span, typ, payload: bytes, n_items: int) -> None:
        raw = struct.pack("<IIIIIBBH", cid, start, end, step, span, typ, 0, n_items) + payload
        self._max_uncomp = max(self._max_uncomp, len(raw))
        self._pending.append((cid, start, end, raw))
        self._pending_bytes += len(raw)
        if self._pending_bytes >= _FLUSH_BYTES:
            self._flush()

    def _flush(self, wait: bool = False) -> None:
        if self._pending:
            pend, self._pending, self._pending_bytes = self._pending, [], 0
            self._inflight.append(self._worker.submit(self._write_batch, pend))
        while self._inflight and (wait or len(self._inflight) > 2 or self._inflight[0].done()):
            try:
                self._inflight.pop(0).result()
            except Exception as e:   # noqa: BLE001 - a failed background batch (deflate / write) is an I/O
                # failure of the file, not an "out of order interval": never a RuntimeError
                from ..exceptions import BigWigWriteError
                raise BigWigWriteError(f"bigWig section batch failed: {e}") from e

    def _write_batch(self, pend) -> None:
        for (cid, start, end, _), comp in zip(pend, _deflate_sections([p[3] for p in pend])):
            off = self._fh.tell()
            self._fh.write(comp)
            self._sections.append((cid, start, end, off, len(comp)))

    def _add_fixed(self, cid, start, span, step, vals) -> None:
        n = vals.size
        end = start + (n - 1) * step + span
        self._check_order(cid, start, end)
        for i in range(0, n, _ITEMS_PER_SECTION):
            v = vals[i: i + _ITEMS_PER_SECTION]
            s = start + i * step
            self._emit(cid, s, s + (v.size - 1) * step + span, step, span, 3, v.astype("<f4").tobytes(), v.size)
        self._stats(vals, span)
        self._last = (cid, end)

    def _add_bedgraph(self, cid, st, en, vals) -> None:
        if st.size == 0:
            return
        if np.any(st[1:] < en[:-1]) or np.any(en <= st):
            raise RuntimeError("The entries you tried to add are out of order or overlap")
        self._check_order(cid, int(st[0]), int(en[-1]))
        if st.size > 1:
            # per-base tracks (adjust_wps output): constant span and step -> fixedStep sections, 4 bytes
            # per item instead of 12; readers see the same (start, end, value) intervals
            span, step = int(en[0] - st[0]), int(st[1] - st[0])
            if step >= span and np.all(en - st == span) and np.all(st[1:] - st[:-1] == step):
                self._add_fixed(cid, int(st[0]), span, step, vals)
                return
        for i in range(0, st.size, _ITEMS_PER_SECTION):
            s, e, v = st[i: i + _ITEMS_PER_SECTION], en[i: i + _ITEMS_PER_SECTION], vals[i: i + _ITEMS_PER_SECTION]
            rec = np.empty(s.size, dtype=[("s", "<u4"), ("e", "<u4"), ("v", "<f4")])
            rec["s"], rec["e"], rec["v"] = s, e, v
            self._emit(cid, int(s[0]), int(e[-1]), 0, 0, 1, rec.tobytes(), s.size)
        self._stats(vals, en - st)
        self._last = (cid, int(en[-1]))

    def _write_rtree(self, index_off: int) -> None:
        fh = self._fh
        secs = self._sections
        n = len(secs)
        if n == 0:
            fh.write(struct.pack("<IIQIIIIQII", _RTREE_MAGIC, _RTREE_BLOCK, 0, 0, 0, 0, 0, index_off, 1, 0))
            fh.write(struct.pack("<BBH", 1, 0, 0))
            return
        # level 0 = leaves over sections; upper levels group _RTREE_BLOCK children
        levels = [[(s[0], s[1], s[0], s[2], i) for i, s in enumerate(secs)]]
        while len(levels[-1]) > _RTREE_BLOCK:
            prev, cur = levels[-1], []
            for i in range(0, len(prev), _RTREE_BLOCK):
                grp = prev[i: i + _RTREE_BLOCK]
                cur.append((grp[0][0], grp[0][1], grp[-1][2], max(g[3] for g in grp if g[2] == grp[-1][2]), i))
            levels.append(cur)
        fh.write(struct.pack("<IIQIIIIQII", _RTREE_MAGIC, _RTREE_BLOCK, n, secs[0][0], secs[0][1],
                             secs[-1][0], secs[-1][2], index_off, 1, 0))
        # node sizes per level, top-down layout
        top = len(levels) - 1
        node_counts = [(-(-len(levels[k]) // _RTREE_BLOCK)) for k in range(len(levels))]
        offsets = {}
        pos = fh.tell()
        for k in range(top, -1, -1):
            item = 32 if k == 0 else 24
            for j in range(node_counts[k]):
                cnt = min(_RTREE_BLOCK, len(levels[k]) - j * _RTREE_BLOCK)
                offsets[(k, j)] = pos
                pos += 4 + cnt * item
        for k in range(top, -1, -1):
            for j in range(node_counts[k]):
                items = levels[k][j * _RTREE_BLOCK: (j + 1) * _RTREE_BLOCK]
                fh.write(struct.pack("<BBH", 1 if k == 0 else 0, 0, len(items)))
                for it in items:
                    if k == 0:
                        s = secs[it[4]]
                        fh.write(struct.pack("<IIIIQQ", s[0], s[1], s[0], s[2], s[3], s[4]))
                    else:
                        child = offsets[(k - 1, it[4] // _RTREE_BLOCK)]
                        fh.write(struct.pack("<IIIIQ", it[0], it[1], it[2], it[3], child))


class BigWigReader:
    def __init__(self, path: str):
        self.path = str(path)
        with builtins.open(self.path, "rb") as fh:
            self._buf = fh.read()
        b = self._buf
        magic = struct.unpack_from("<I", b, 0)[0]
        self._e = "<"
        if magic != _BW_MAGIC:
            if struct.unpack_from(">I", b, 0)[0] != _BW_MAGIC:
                raise RuntimeError(f"{path} is not a bigWig file")
            self._e = ">"
        e = self._e
        (_, self.version, self.n_zoom, ct_off, self._data_off, self._index_off, _, _, _, self._summary_off,
         self._uncomp, _) = struct.unpack_from(e + "IHHQQQHHQQIQ", b, 0)
        self._chroms: dict[str, tuple[int, int]] = {}
        _, _, key_size, _, _, _ = struct.unpack_from(e + "IIIIQQ", b, ct_off)
        self._walk_chrom_tree(ct_off + 32, key_size)
        self._by_id = {cid: (name, size) for name, (cid, size) in self._chroms.items()}
        self._cache: dict = {}               # data offset -> inflated section (uint8 view)
        self._sections = None                # flat R-tree leaves, built on the first query
        self._spanning = ()

    def _walk_chrom_tree(self, off: int, key_size: int) -> None:
        b, e = self._buf, self._e
        is_leaf, _, count = struct.unpack_from(e + "BBH", b, off)
        off += 4
        for _ in range(count):
            key = b[off: off + key_size].rstrip(b"\0").decode()
            if is_leaf:
                cid, size = struct.unpack_from(e + "II", b, off + key_size)
                self._chroms[key] = (cid, size)
            else:
                (child,) = struct.unpack_from(e + "Q", b, off + key_size)
                self._walk_chrom_tree(child, key_size)
            off += key_size + 8

    def chroms(self, chrom=None):
        if chrom is not None:
            return self._chroms[chrom][1] if chrom in self._chroms else None
        return {c: s for c, (_, s) in self._chroms.items()}

    def header(self):
        b, e = self._buf, self._e
        n, mn, mx, sm, sq = struct.unpack_from(e + "Qdddd", b, self._summary_off)
        return {"version": self.version, "nLevels": self.n_zoom, "nBasesCovered": n, "minVal": mn,
                "maxVal": mx, "sumData": sm, "sumSquared": sq}

    def _load_index(self) -> None:
        """Flatten the R-tree once: every leaf (data section) as rows of numpy arrays, per chromosome
        sorted by start with a running maximum of the ends, so a query is two binary searches."""
        b, e = self._buf, self._e
        leaf_t = np.dtype([("sc", e + "u4"), ("sb", e + "u4"), ("ec", e + "u4"), ("eb", e + "u4"),
                           ("off", e + "u8"), ("size", e + "u8")])
        node_t = np.dtype([("sc", e + "u4"), ("sb", e + "u4"), ("ec", e + "u4"), ("eb", e + "u4"), ("child", e + "u8")])
        leaves, stack = [], [self._index_off + 48]
        while stack:
            off = stack.pop()
            is_leaf, _, count = struct.unpack_from(e + "BBH", b, off)
            if is_leaf:
                leaves.append(np.frombuffer(b, leaf_t, count, off + 4))
            else:
                stack.extend(np.frombuffer(b, node_t, count, off + 4)["child"].tolist()[::-1])
        rows = np.concatenate(leaves) if leaves else np.zeros(0, leaf_t)
        self._sections = {}
        for cid in np.unique(rows["sc"]).tolist() if rows.size else []:
            r = rows[(rows["sc"] == cid) & (rows["ec"] == cid)]
            r = r[np.argsort(r["sb"], kind="stable")]
            ends = r["eb"].astype(np.int64)
            self._sections[cid] = (r["sb"].astype(np.int64), ends, np.maximum.accumulate(ends) if ends.size else ends,
                                   r["off"].astype(np.int64), r["size"].astype(np.int64))
        # sections spanning several chromosomes do not occur in bigWig data; keep them reachable anyway
        self._spanning = rows[rows["sc"] != rows["ec"]]

    def _blocks(self, cid: int, start: int, end: int):
        if self._sections is None:
            self._load_index()
        out = []
        sec = self._sections.get(cid)
        if sec is not None:
            sb, eb, eb_max, off, size = sec
            lo = int(np.searchsorted(eb_max, start, side="right"))     # first section whose running end > start
            hi = int(np.searchsorted(sb, end, side="left"))            # sections starting before the query end
            if hi > lo:
                keep = eb[lo:hi] > start
                out = list(zip(off[lo:hi][keep].tolist(), size[lo:hi][keep].tolist()))
        for r in self._spanning:
            if (int(r["sc"]), int(r["sb"])) < (cid, end) and (int(r["ec"]), int(r["eb"])) > (cid, start):
                out.append((int(r["off"]), int(r["size"])))
        return out

    def _section(self, doff: int, dsize: int) -> bytes:
        raw = self._cache.get(doff)
        if raw is None:
            raw = self._buf[doff: doff + dsize]
            if self._uncomp:
                (raw,) = _inflate_sections(self._buf, [(doff, dsize)], self._uncomp)
        return raw

    def prefetch(self, queries) -> None:
        """Inflate, in one multi-threaded batch, every section the (chrom, start, end) queries touch.

        Later ``intervals`` / ``intervals_arrays`` / ``values`` calls on those ranges are served from
        the cache.  Invalid queries are ignored here; they raise when actually queried."""
        if not self._uncomp:
            return
        need: dict[int, int] = {}
        for chrom, start, end in queries:
            if chrom not in self._chroms:
                continue
            cid, size = self._chroms[chrom]
            start = 0 if start is None else int(start)
            end = size if end is None or end == 0 else int(end)
            if start < 0 or end > size or start >= end:
                continue
            for doff, dsize in self._blocks(cid, start, end):
                if doff not in self._cache:
                    need[doff] = dsize
        blocks = sorted(need.items())
        per_batch = max(1, _FLUSH_BYTES // max(int(self._uncomp), 1))
        for i in range(0, len(blocks), per_batch):
            part = blocks[i: i + per_batch]
            for (doff, _), raw in zip(part, _inflate_sections(self._buf, part, self._uncomp)):
                self._cache[doff] = raw

    def drop_cache(self) -> None:
        """Forget the sections inflated by ``prefetch`` (they are re-read on demand)."""
        self._cache = {}

    def _intervals_arrays(self, chrom, start, end):
        if chrom not in self._chroms:
            raise RuntimeError("Invalid interval bounds!")
        cid, size = self._chroms[chrom]
        start = 0 if start is None else int(start)
        end = size if end is None or end == 0 else int(end)
        if start < 0 or end > size or start >= end:
            raise RuntimeError("Invalid interval bounds!")
        b, e = self._buf, self._e
        S, E, V = [], [], []
        blocks = self._blocks(cid, start, end)
        local = {}
        if self._uncomp and len(blocks) > 8:
            # a long query: inflate its sections in multi-threaded batches instead of one zlib call each
            missing = [b for b in blocks if b[0] not in self._cache]
            per_batch = max(1, _FLUSH_BYTES // max(int(self._uncomp), 1))
            for i in range(0, len(missing), per_batch):
                part = missing[i: i + per_batch]
                local.update(zip((d for d, _ in part), _inflate_sections(self._buf, part, self._uncomp)))
        for doff, dsize in blocks:
            raw = local.pop(doff, None)
            if raw is None:
                raw = self._section(doff, dsize)
            bcid, bstart, bend, step, span, typ, _, n = struct.unpack_from(e + "IIIIIBBH", raw, 0)
            if bcid != cid:
                continue
            if typ == 3:
                if n == 0:
                    continue
                # fixedStep: clip by index arithmetic instead of building and masking all items
                i0 = max(0, -(-(start - span + 1 - bstart) // step)) if start > bstart else 0   # first item with end > start
                i1 = min(n, -(-(end - bstart) // step))                                          # items with start < end
                if i1 <= i0:
                    continue
                v = np.frombuffer(raw, e + "f4", i1 - i0, 24 + 4 * i0)
                s = np.arange(bstart + i0 * step, bstart + i1 * step, step, dtype=np.int64)
                S.append(s); E.append(s + span); V.append(v)
                continue
            elif typ == 2:
                rec = np.frombuffer(raw, np.dtype([("s", e + "u4"), ("v", e + "f4")]), n, 24)
                s = rec["s"].astype(np.int64); en = s + span; v = rec["v"]
            else:
                rec = np.frombuffer(raw, np.dtype([("s", e + "u4"), ("e", e + "u4"), ("v", e + "f4")]), n, 24)
                s = rec["s"].astype(np.int64); en = rec["e"].astype(np.int64); v = rec["v"]
            if n == 0:
                continue
            if bstart < start or bend > end:      # only boundary sections need clipping
                m = (en > start) & (s < end)
                if not m.any():
                    continue
                s, en, v = s[m], en[m], v[m]
            S.append(s); E.append(en); V.append(v)
        if not S:
            return None
        if len(S) == 1:
            return S[0], E[0], V[0]
        s, en, v = np.concatenate(S), np.concatenate(E), np.concatenate(V)
        if np.all(s[1:] >= s[:-1]):
            return s, en, v
        order = np.argsort(s, kind="stable")
        return s[order], en[order], v[order]

    def per_base_run(self, chrom, start, end):
        """``(first position, float32 values)`` when the entries overlapping ``[start, end)`` are per-base
        fixedStep items (span 1, step 1: what ``multi_wps`` writes) forming ONE run without gap or overlap - the
        same information as ``intervals_arrays`` without building a position per value.  ``None`` when the
        range holds anything else (nothing, other section types, gaps): the caller then takes the general
        query.  Invalid bounds raise like ``intervals``."""
        if chrom not in self._chroms:
            raise RuntimeError("Invalid interval bounds!")
        cid, size = self._chroms[chrom]
        start = 0 if start is None else int(start)
        end = size if end is None or end == 0 else int(end)
        if start < 0 or end > size or start >= end:
            raise RuntimeError("Invalid interval bounds!")
        e = self._e
        first, nxt, parts = None, None, []
        for doff, dsize in self._blocks(cid, start, end):
            raw = self._section(doff, dsize)
            bcid, bstart, _, step, span, typ, _, n = struct.unpack_from(e + "IIIIIBBH", raw, 0)
            if bcid != cid or typ != 3 or step != 1 or span != 1:
                return None
            i0, i1 = max(start - bstart, 0), min(end - bstart, n)
            if i1 <= i0:
                continue
            if nxt is not None and bstart + i0 != nxt:
                return None
            if first is None:
                first = bstart + i0
            nxt = bstart + i1
            parts.append(np.frombuffer(raw, e + "f4", i1 - i0, 24 + 4 * i0))
        if first is None:
            return None
        return first, (parts[0] if len(parts) == 1 else np.concatenate(parts))

    def intervals(self, chrom, start=0, end=0):
        """pyBigWig.intervals: tuple of (start, end, value) overlapping the range, or None."""
        r = self._intervals_arrays(chrom, start, end)
        if r is None:
            return None
        return tuple((int(a), int(b), float(c)) for a, b, c in zip(*r))

    def intervals_arrays(self, chrom, start=0, end=0):
        """Same query as ``intervals`` but as (starts int64, ends int64, values float32) arrays."""
        return self._intervals_arrays(chrom, start, end)

    def values(self, chrom, start=0, end=0, numpy=True):
        """pyBigWig.values: one float32 per base of [start, end), NaN where uncovered."""
        if chrom not in self._chroms:
            raise RuntimeError("Invalid interval bounds!")
        size = self._chroms[chrom][1]
        end = size if end == 0 else end
        r = self._intervals_arrays(chrom, start, end)      # validates the bounds like pyBigWig
        out = np.full(end - start, np.nan, dtype=np.float32)
        if r is not None:
            s, e_, v = r
            if np.all(e_ - s == 1):                          # per-base entries (WPS tracks)
                out[s - start] = v
            else:
                for a, b, x in zip(s.tolist(), e_.tolist(), v.tolist()):
                    out[max(a, start) - start: min(b, end) - start] = x
        return out if numpy else out.tolist()

    def close(self) -> None:
        self._buf = b""
        self._cache = {}

    def __enter__(self):
        return self

    def __exit__(self, *exc) -> None:
        self.close()


def open(path, mode: str = "r"):
    """pyBigWig.open look-alike."""
    if "w" in mode:
        return BigWigWriter(path)
    return BigWigReader(path)
