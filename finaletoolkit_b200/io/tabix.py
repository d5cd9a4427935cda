"""Tabix (.tbi) index reader: where each contig's records live in a BGZF fragment file.

The reference reaches fragments through ``pysam.TabixFile.fetch(contig, start, stop)``
(io/alignment.py:270-279), i.e. through this index.  The hot path wants whole contigs as columns,
so the only thing read from the index is, per sequence name, the virtual-offset range of its
records - enough to inflate and parse one contig of a large file without touching the rest
(``ftk_fragfile_open_slice``).  Format: SAM/tabix specification section 5.2 (BGZF-compressed,
magic ``TBI\\1``, binning index + linear index per sequence; bin 37450 is htslib's pseudo-bin with
the sequence's first / last virtual offsets).
"""
from __future__ import annotations

import gzip
import struct
from typing import NamedTuple

__all__ = ["TabixIndex", "read_tbi"]

_PSEUDO_BIN = 37450


class TabixIndex(NamedTuple):
    names: list            # sequence names in file order
    ranges: dict           # name -> (coffset_beg, uoffset_beg, coffset_end, uoffset_end)
    col_seq: int
    col_beg: int
    col_end: int
    meta: str


def read_tbi(path: str) -> TabixIndex | None:
    """Parse ``path`` (a .tbi); None when it is not a tabix index (e.g. a placeholder file)."""
    try:
        with gzip.open(path, "rb") as fh:
            raw = fh.read()
    except (OSError, EOFError):
        return None
    if len(raw) < 36 or raw[:4] != b"TBI\x01":
        return None
    try:
        n_ref, _fmt, col_seq, col_beg, col_end, meta, _skip, l_nm = struct.unpack_from("<8i", raw, 4)
        names = [n.decode() for n in raw[36: 36 + l_nm].split(b"\0")[:n_ref]]
        off = 36 + l_nm
        ranges = {}
        for name in names:
            (n_bin,) = struct.unpack_from("<i", raw, off); off += 4
            lo = hi = None
            pseudo = None
            for _ in range(n_bin):
                bin_id, n_chunk = struct.unpack_from("<Ii", raw, off); off += 8
                chunks = struct.unpack_from(f"<{2 * n_chunk}Q", raw, off); off += 16 * n_chunk
                if bin_id == _PSEUDO_BIN:
                    if n_chunk:
                        pseudo = (chunks[0], chunks[1])
                    continue
                for beg, end in zip(chunks[0::2], chunks[1::2]):
                    lo = beg if lo is None else min(lo, beg)
                    hi = end if hi is None else max(hi, end)
            (n_intv,) = struct.unpack_from("<i", raw, off); off += 4 + 8 * n_intv
            if lo is None and pseudo is not None:
                lo, hi = pseudo
            if lo is not None:
                ranges[name] = (lo >> 16, lo & 0xFFFF, hi >> 16, hi & 0xFFFF)
        return TabixIndex(names, ranges, col_seq, col_beg, col_end, chr(meta) if 0 < meta < 128 else "#")
    except (struct.error, UnicodeDecodeError):
        return None
