"""In-memory stand-in for ``pysam`` (TEST INFRASTRUCTURE ONLY).

pysam is not installable in this image (no network).  This module lets the
UNMODIFIED reference package under /root/reference/src be imported so that
``oracle/make_golden.py`` can record its outputs as golden vectors.  It is
never imported by the product package.

Only what the reference's hot path touches is modelled:

* ``TabixFile``  - ``.contigs``, ``.fetch(reference, start, end, parser,
  multiple_iterators)`` with tabix overlap semantics
  (``rec.stop > start and rec.start < end``), ``.close()``.
  (reference call site: src/finaletoolkit/io/alignment.py:270-302)
* ``asTuple``    - parser marker.
* ``AlignmentFile`` / ``AlignedSegment`` / ``FastaFile`` - class names for the
  reference's ``isinstance`` checks; not functional.
"""
from __future__ import annotations

import bisect
import gzip


class AlignedSegment:  # pragma: no cover - name only
    pass


class AlignmentHeader:  # pragma: no cover - name only
    pass


class AlignmentFile:  # pragma: no cover - name only
    def __init__(self, *a, **k):
        raise NotImplementedError("fake pysam has no BAM/CRAM reader")


class FastaFile:  # pragma: no cover - name only
    def __init__(self, *a, **k):
        raise NotImplementedError("fake pysam has no FASTA reader")


def faidx(*a, **k):  # pragma: no cover
    raise NotImplementedError


class asTuple:
    pass


class TabixFile:
    """Serve fragment rows from memory.

    ``TabixFile(path)`` parses a (b)gzip text file; ``TabixFile.from_columns``
    serves columnar arrays.  Rows are kept per contig in file order (which is
    start-sorted for a tabix-indexable file).
    """

    def __init__(self, path=None, *_, **__):
        self._rows = {}      # contig -> list[tuple[str,...]]
        self._starts = {}    # contig -> list[int]
        self._maxlen = {}    # contig -> int
        self.contigs = []
        self.filename = path
        if path is not None:
            with gzip.open(str(path), "rt") as fh:
                for line in fh:
                    if not line.strip() or line.startswith("#"):
                        continue
                    f = tuple(line.rstrip("\n").split("\t"))
                    self._rows.setdefault(f[0], []).append(f)
            self._finish()

    @classmethod
    def from_columns(cls, columns, bed6=False):
        """columns: {contig: (start[], stop[], mapq[], strand_is_plus[])}"""
        self = cls(None)
        for contig, (st, sp, mq, fw) in columns.items():
            rows = []
            for a, b, q, s in zip(st.tolist(), sp.tolist(), mq.tolist(), fw.tolist()):
                if bed6:
                    rows.append((contig, str(a), str(b), ".", str(q), "+" if s else "-"))
                else:
                    rows.append((contig, str(a), str(b), str(q), "+" if s else "-"))
            self._rows[contig] = rows
        self._finish()
        return self

    def _finish(self):
        self.contigs = list(self._rows.keys())
        for c, rows in self._rows.items():
            self._starts[c] = [int(r[1]) for r in rows]
            self._maxlen[c] = max((int(r[2]) - int(r[1]) for r in rows), default=0)

    def fetch(self, reference=None, start=None, end=None, region=None,
              parser=None, multiple_iterators=False):
        if reference is None:
            for c in self.contigs:
                yield from self._rows[c]
            return
        if reference not in self._rows:
            raise ValueError(f"could not create iterator for region '{reference}'")
        rows = self._rows[reference]
        starts = self._starts[reference]
        lo_q = 0 if start is None else int(start)
        hi_q = None if end is None else int(end)
        lo = bisect.bisect_left(starts, lo_q - self._maxlen[reference])
        hi = len(rows) if hi_q is None else bisect.bisect_left(starts, hi_q)
        for i in range(lo, hi):
            r = rows[i]
            if int(r[2]) > lo_q:
                yield r

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
