"""Pure-Python stand-in for ``py2bit`` (TEST INFRASTRUCTURE ONLY).

Reads real UCSC .2bit files (signature 0x1A412743, codes T0 C1 A2 G3,
MSB-first, N-blocks, soft-mask blocks ignored -> upper case like py2bit's
default) so the unmodified reference's ``ReferenceWrapper``
(src/finaletoolkit/io/reference.py:89-96,185-189) works in this image.
Used only by ``oracle/make_golden.py``.
"""
from __future__ import annotations

import builtins
import struct

import numpy as np

_CODE = np.frombuffer(b"TCAG", dtype=np.uint8)


class _TwoBit:
    def __init__(self, path):
        with builtins.open(path, "rb") as fh:
            buf = fh.read()
        sig, ver, n_seq, _ = struct.unpack_from("<IIII", buf, 0)
        if sig != 0x1A412743:
            raise RuntimeError("fake py2bit: bad signature / big-endian file")
        off = 16
        index = []
        for _ in range(n_seq):
            ln = buf[off]
            name = buf[off + 1: off + 1 + ln].decode()
            (o,) = struct.unpack_from("<I", buf, off + 1 + ln)
            index.append((name, o))
            off += 1 + ln + 4
        self._seqs = {}
        self._chroms = {}
        for name, o in index:
            (dna_size,) = struct.unpack_from("<I", buf, o)
            o += 4
            (nb,) = struct.unpack_from("<I", buf, o)
            o += 4
            n_starts = np.frombuffer(buf, "<u4", nb, o)
            o += 4 * nb
            n_sizes = np.frombuffer(buf, "<u4", nb, o)
            o += 4 * nb
            (mb,) = struct.unpack_from("<I", buf, o)
            o += 4 + 8 * mb + 4
            packed = np.frombuffer(buf, np.uint8, (dna_size + 3) // 4, o)
            codes = np.empty(packed.size * 4, np.uint8)
            codes[0::4] = packed >> 6
            codes[1::4] = (packed >> 4) & 3
            codes[2::4] = (packed >> 2) & 3
            codes[3::4] = packed & 3
            seq = _CODE[codes[:dna_size]].copy()
            for s, z in zip(n_starts.tolist(), n_sizes.tolist()):
                seq[s: s + z] = ord("N")
            self._seqs[name] = seq
            self._chroms[name] = int(dna_size)

    def chroms(self, chrom=None):
        if chrom is not None:
            return self._chroms[chrom]
        return dict(self._chroms)

    def sequence(self, chrom, start=0, end=0):
        if chrom not in self._seqs:
            raise RuntimeError("Invalid chromosome")
        n = self._chroms[chrom]
        if end == 0:
            end = n
        if start < 0 or end > n or start > end:
            raise RuntimeError("bounds are invalid")
        return self._seqs[chrom][start:end].tobytes().decode("ascii")

    def close(self):
        pass


_CACHE: dict = {}


def open(path, storeMasked=False):
    path = str(path)
    if path not in _CACHE:
        _CACHE[path] = _TwoBit(path)
    return _CACHE[path]
