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
* ``AlignmentFile`` / ``AlignedSegment`` - a pure-Python BAM record reader exposing the fields
  the reference's read filter and fragment reconstruction use (io/alignment.py:60-71,242-268).
* ``FastaFile`` - class name for the reference's ``isinstance`` checks; not functional.
"""
from __future__ import annotations

import bisect
import gzip


class AlignedSegment:
    """The fields of one BAM record that io/alignment.py:60-71,242-268 looks at, with pysam's meanings."""

    def __init__(self, ref_name, pos, mapq, flag, tlen, cigar):
        self.reference_name, self.reference_start, self.mapping_quality = ref_name, pos, mapq
        self.flag, self.template_length, self._cigar = flag, tlen, cigar
        self.is_paired = bool(flag & 0x1); self.is_proper_pair = bool(flag & 0x2)
        self.is_unmapped = bool(flag & 0x4); self.mate_is_unmapped = bool(flag & 0x8)
        self.is_reverse = bool(flag & 0x10); self.is_forward = not self.is_reverse
        self.is_read1 = bool(flag & 0x40); self.is_read2 = bool(flag & 0x80)
        self.is_secondary = bool(flag & 0x100); self.is_qcfail = bool(flag & 0x200)
        self.is_duplicate = bool(flag & 0x400); self.is_supplementary = bool(flag & 0x800)

    @property
    def reference_end(self):   # pysam: aligned end (exclusive); None without an alignment
        if self.is_unmapped or not self._cigar:
            return None
        return self.reference_start + sum(n for op, n in self._cigar if op in (0, 2, 3, 7, 8))


class AlignmentHeader:  # pragma: no cover - name only
    pass


class AlignmentFile:
    """Minimal BAM reader (whole file, pure Python): header names / lengths and ``fetch`` in file order
    with htslib's region rule (read overlaps [start, stop)).  CRAM / SAM text are not modelled."""

    def __init__(self, path, mode="r", *_, **__):
        import struct
        self.filename = str(path)
        if not self.filename.lower().endswith(".bam"):
            raise NotImplementedError("fake pysam reads BAM only")
        raw = gzip.open(self.filename, "rb").read()
        if raw[:4] != b"BAM\x01":
            raise ValueError("not a BAM file")
        (l_text,) = struct.unpack_from("<i", raw, 4); off = 8 + l_text
        (n_ref,) = struct.unpack_from("<i", raw, off); off += 4
        self.references, self.lengths = [], []
        for _ in range(n_ref):
            (ln,) = struct.unpack_from("<i", raw, off)
            self.references.append(raw[off + 4: off + 4 + ln - 1].decode())
            self.lengths.append(struct.unpack_from("<i", raw, off + 4 + ln)[0]); off += 8 + ln
        self._reads = []
        while off < len(raw):
            (bs,) = struct.unpack_from("<i", raw, off)
            ref_id, pos, l_rn, mapq, _bin, n_cig, flag, _l_seq, _nref, _npos, tlen = struct.unpack_from("<iiBBHHHiiii", raw, off + 4)
            cig = [(c & 15, c >> 4) for c in struct.unpack_from(f"<{n_cig}I", raw, off + 36 + l_rn)]
            self._reads.append(AlignedSegment(self.references[ref_id] if ref_id >= 0 else None, pos, mapq, flag, tlen, cig))
            off += 4 + bs

    def fetch(self, contig=None, start=None, stop=None, **_):
        for r in self._reads:
            if r.reference_name is None:
                continue   # unplaced reads are not part of an indexed fetch
            if contig is not None:
                if r.reference_name != contig:
                    continue
                end = r.reference_end if r.reference_end is not None else r.reference_start + 1
                if (stop is not None and r.reference_start >= stop) or (start is not None and end <= start):
                    continue
            yield r

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


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
