"""``AlignmentWrapper`` / ``Fragment`` - the reference's unified fragment reader (io/alignment.py:25-320) over the
columnar ``FragmentTable``.

The reference opens a pysam handle and streams one Python object per fragment; the features of this package
never do that (their predicates run in the CUDA kernels on whole columns).  This class exists for callers that
used the reader directly: same constructor arguments, ``chroms`` / ``is_sam`` / ``fetch`` / context manager, the
same records.  ``fetch(contig, start, stop)`` makes the selection an index query makes - rows overlapping the
region for a tabix-indexed fragment file (io/alignment.py:270-302), READS overlapping it for BAM
(:242-268, ``FragmentTable.fetched(fetch_only=True)``) - and applies the mapq cut.  Records come in table order
(sorted by fragment start; a BAM streams them by read position instead).
"""
from __future__ import annotations

from typing import Dict, Generator, NamedTuple, Optional

import numpy as np

from .fragments import FragmentTable, as_table

__all__ = ["Fragment", "AlignmentWrapper"]


class Fragment(NamedTuple):
    """One fragment: 0-based half-open ``[start, stop)``, mapping quality, strand of read 1."""
    contig: str
    start: int
    stop: int
    mapq: int
    is_forward: bool

    @property
    def length(self) -> int:
        return self.stop - self.start


class AlignmentWrapper:
    def __init__(self, path, reference_file=None, threads: int = 1, quality_threshold: int = 30,
                 read1_only: bool = True) -> None:
        if not read1_only:
            raise NotImplementedError("read1_only=False (every fragment once per mate) is not available: the decoder "
                                      "keeps one row per fragment")
        self.path = path if isinstance(path, str) else getattr(path, "filename", None)
        self.reference_file = str(reference_file) if reference_file else None
        self.threads = threads
        self.quality_threshold = quality_threshold
        self.read1_only = read1_only
        self._table: Optional[FragmentTable] = as_table(path, reference_file)

    @property
    def chroms(self) -> Dict[str, Optional[int]]:
        """Contig -> length from a BAM header, contig -> None for a fragment file (io/alignment.py:206-209)."""
        t = self._table
        if t.is_sam and t.contig_lengths:
            return dict(t.contig_lengths)
        return {c: None for c in t.contigs}

    @property
    def is_sam(self) -> bool:
        return bool(self._table.is_sam)

    def fetch(self, contig: Optional[str] = None, start: Optional[int] = None,
              stop: Optional[int] = None) -> Generator[Fragment, None, None]:
        table = self._table
        for c in ([contig] if contig is not None else table.contigs):
            rows = table
            if contig is not None and table.has_read1(c):
                rows = table.fetched(c, start, stop, fetch_only=True)          # BAM: the reads decide
            st, sp, mq, sd = rows.host(c)
            keep = mq >= self.quality_threshold
            if contig is not None and not table.has_read1(c):                 # tabix: rows overlapping the region
                keep &= sp.astype(np.int64) > (0 if start is None else start)
                if stop is not None:
                    keep &= st.astype(np.int64) < stop
            for i in np.flatnonzero(keep).tolist():
                yield Fragment(c, int(st[i]), int(sp[i]), int(mq[i]), bool(sd[i]))

    def close(self) -> None:
        self._table = None

    def __enter__(self) -> "AlignmentWrapper":
        return self

    def __exit__(self, exc_type, exc_val, exc_tb) -> None:
        self.close()
