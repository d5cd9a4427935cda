"""In-memory stand-in for ``pyBigWig`` (TEST INFRASTRUCTURE ONLY).

Lets the unmodified reference write/read "bigWig files" that live in a
process-global dict keyed by path.  Values are rounded through float32 exactly
as libBigWig stores them.  Used only by ``oracle/make_golden.py``.

Reference call sites: src/finaletoolkit/frag/_multi_wps.py:300-325 (writer,
``addEntries(chrom, start, values=, span=1, step=1)``),
src/finaletoolkit/frag/_adjust_wps.py:80-105 (``intervals``) and :275-291
(``addEntries(chroms, starts, ends=, values=)``), src/finaletoolkit/utils/_agg_bw.py:89
(``values``).
"""
from __future__ import annotations

import numpy as np

_STORE: dict = {}


class _BW:
    def __init__(self, path, mode):
        self.path = str(path)
        self.mode = mode
        if "w" in mode:
            _STORE[self.path] = {"header": None, "data": {}}
        elif self.path not in _STORE:
            raise RuntimeError(f"fake pyBigWig: no such file {path}")
        self._f = _STORE[self.path]

    # -- writer ---------------------------------------------------------
    def addHeader(self, header, maxZooms=10):
        self._f["header"] = list(header)

    def addEntries(self, chroms, starts, ends=None, values=None, span=None,
                   step=None, validate=True):
        vals = np.asarray(values, dtype=np.float64).astype(np.float32)
        if isinstance(chroms, str):
            if span is not None and step is not None:
                pos = int(starts) + np.arange(len(vals), dtype=np.int64) * int(step)
                stops = pos + int(span)
            else:
                pos = np.asarray(starts, dtype=np.int64)
                stops = pos + int(span) if ends is None else np.asarray(ends, dtype=np.int64)
            chrom_list = [chroms] * len(vals)
        else:
            chrom_list = list(chroms)
            pos = np.asarray(starts, dtype=np.int64)
            stops = np.asarray(ends, dtype=np.int64)
        if len(chrom_list) == 0:
            return
        d = self._f["data"]
        c0 = chrom_list[0]
        if any(c != c0 for c in chrom_list):
            raise RuntimeError("fake pyBigWig: mixed contigs in one call")
        ent = d.setdefault(c0, [])
        if ent and int(pos[0]) < int(ent[-1][1][-1]):
            raise RuntimeError("The entries you tried to add are out of order")
        ent.append((pos.copy(), stops.copy(), vals.copy()))

    # -- reader ---------------------------------------------------------
    def chroms(self):
        return dict(self._f["header"] or [])

    def intervals(self, chrom, start=0, end=None):
        # pyBigWig.c pyBwGetIntervals: unknown contig, end <= start or end > chromLen
        # -> RuntimeError("Invalid interval bounds!")
        sizes = dict(self._f["header"] or [])
        if chrom not in sizes:
            raise RuntimeError("Invalid interval bounds!")
        if end is None:
            end = sizes[chrom]
        if end <= start or end > sizes[chrom] or start < 0:
            raise RuntimeError("Invalid interval bounds!")
        ent = self._f["data"].get(chrom)
        if not ent:
            return None
        pos = np.concatenate([e[0] for e in ent])
        stops = np.concatenate([e[1] for e in ent])
        vals = np.concatenate([e[2] for e in ent])
        if end is None:
            end = int(stops.max())
        m = (stops > start) & (pos < end)
        if not m.any():
            return None
        return tuple(
            (int(a), int(b), float(v)) for a, b, v in zip(pos[m], stops[m], vals[m])
        )

    def values(self, chrom, start=0, end=None, numpy=False):
        # pyBigWig.c pyBwGetValues: same bounds rule; uncovered bases are nan; python floats
        sizes = dict(self._f["header"] or [])
        if chrom not in sizes:
            raise RuntimeError("Invalid interval bounds!")
        if end is None:
            end = sizes[chrom]
        if end <= start or end > sizes[chrom] or start < 0:
            raise RuntimeError("Invalid interval bounds!")
        out = np.full(end - start, np.nan, dtype=np.float32)
        for pos, stops, vals in self._f["data"].get(chrom, []):
            m = (stops > start) & (pos < end)
            if np.all(stops[m] - pos[m] == 1):
                out[pos[m] - start] = vals[m]
                continue
            for a, b, v in zip(pos[m].tolist(), stops[m].tolist(), vals[m].tolist()):
                out[max(a, start) - start: min(b, end) - start] = v
        return out if numpy else [float(x) for x in out]

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def open(path, mode="r"):
    return _BW(path, mode)
