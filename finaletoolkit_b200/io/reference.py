"""Reference genome access for the end-motif kernels: .2bit / FASTA -> 2-bit packed HBM contigs.

Mirrors ``ReferenceWrapper`` (reference io/reference.py:35-222): ``.chroms``,
``.sequence(contig, start, stop, fail_on_excess_range=True)`` (upper-cased, bounds
errors as ``OutOfBoundsError``, unknown contig ``ContigNotFoundError``) and the
context-manager protocol, but decodes with numpy instead of py2bit/pysam and adds
``device_contig()`` which packs a contig for ``ftk_end_motif_hist_u64``
(2 bits per base A0 C1 G2 T3 + an N bit-mask; layout in ``synth.pack_twobit``).
"""
from __future__ import annotations

import gzip
import os
import struct
from typing import Dict

import numpy as np

from ..exceptions import ContigNotFoundError, OutOfBoundsError

__all__ = ["ReferenceWrapper", "open_reference"]

_TWOBIT_SUFFIXES = (".2bit", ".tb2")
# UCSC .2bit codes T0 C1 A2 G3 -> ours A0 C1 G2 T3
_UCSC_TO_ACGT = np.array([3, 1, 0, 2], dtype=np.uint8)
_ASCII = np.frombuffer(b"ACGT", dtype=np.uint8)
# one .2bit byte (4 bases, first base in the top two bits, UCSC codes) -> one byte of the device
# layout (first base in the low two bits, codes A0 C1 G2 T3)
_TWOBIT_BYTE_LUT = np.array([sum(int(_UCSC_TO_ACGT[(v >> (6 - 2 * j)) & 3]) << (2 * j) for j in range(4))
                             for v in range(256)], dtype=np.uint8)
_OPEN: dict = {}


def open_reference(reference_path) -> "ReferenceWrapper":
    """A shared ``ReferenceWrapper`` per file (path, mtime, size): decoded / packed / uploaded contigs
    survive across API calls instead of being rebuilt by every call (the reference re-opens the
    file per worker, frag/_motif_common.py:580-610)."""
    if isinstance(reference_path, ReferenceWrapper):
        return reference_path
    path = os.path.abspath(str(reference_path))
    if not os.path.exists(path):
        raise FileNotFoundError(f"Reference file not found: {reference_path}")
    st = os.stat(path)
    key = (path, st.st_mtime_ns, st.st_size)
    ref = _OPEN.get(key)
    if ref is None:
        if len(_OPEN) >= 4:
            _OPEN.pop(next(iter(_OPEN)))
        ref = _OPEN[key] = ReferenceWrapper(path, use_lock=False)
    return ref


class ReferenceWrapper:
    def __init__(self, reference_path, use_lock: bool = True) -> None:
        self.reference_path = str(reference_path)
        self.use_lock = use_lock
        if not os.path.exists(self.reference_path):
            raise FileNotFoundError(f"Reference file not found: {self.reference_path}")
        self._codes: Dict[str, np.ndarray] = {}
        self._nmask: Dict[str, np.ndarray] = {}
        self._device: dict = {}
        if self.reference_path.endswith(_TWOBIT_SUFFIXES):
            self._is_2bit = True
            self._open_2bit()
        else:
            self._is_2bit = False
            self._open_fasta()

    # -- decoding --------------------------------------------------------
    def _open_2bit(self) -> None:
        with open(self.reference_path, "rb") as fh:
            buf = fh.read()
        sig = struct.unpack_from("<I", buf, 0)[0]
        end = "<"
        if sig != 0x1A412743:
            if struct.unpack_from(">I", buf, 0)[0] != 0x1A412743:
                raise ValueError(f"{self.reference_path} is not a .2bit file")
            end = ">"
        _, _, n_seq, _ = struct.unpack_from(end + "IIII", buf, 0)
        off = 16
        index = []
        for _ in range(n_seq):
            ln = buf[off]
            name = buf[off + 1: off + 1 + ln].decode()
            (o,) = struct.unpack_from(end + "I", buf, off + 1 + ln)
            index.append((name, o))
            off += 1 + ln + 4
        self._buf, self._end, self._index = buf, end, dict(index)
        self._chroms = {}
        for name, o in index:
            self._chroms[name] = int(struct.unpack_from(end + "I", buf, o)[0])

    def _decode_2bit(self, contig: str) -> None:
        buf, end, o = self._buf, self._end, self._index[contig]
        (dna_size,) = struct.unpack_from(end + "I", buf, o); o += 4
        (nb,) = struct.unpack_from(end + "I", buf, o); o += 4
        n_starts = np.frombuffer(buf, end + "u4", nb, o); o += 4 * nb
        n_sizes = np.frombuffer(buf, end + "u4", nb, o); o += 4 * nb
        (mb,) = struct.unpack_from(end + "I", buf, o); o += 4 + 8 * mb + 4
        packed = np.frombuffer(buf, np.uint8, (dna_size + 3) // 4, o)
        codes = np.empty(packed.size * 4, np.uint8)
        codes[0::4] = packed >> 6
        codes[1::4] = (packed >> 4) & 3
        codes[2::4] = (packed >> 2) & 3
        codes[3::4] = packed & 3
        codes = _UCSC_TO_ACGT[codes[:dna_size]]
        nmask = np.zeros(dna_size, dtype=bool)
        for s, z in zip(n_starts.tolist(), n_sizes.tolist()):
            nmask[s: s + z] = True
        self._codes[contig], self._nmask[contig] = codes, nmask

    def packed_words(self, contig: str):
        """``(seq_words, nmask_words)`` in the device layout of ``synth.pack_twobit``.  For .2bit input
        this is one byte-LUT pass over the file's own 2-bit payload (no per-base expansion)."""
        if contig not in self._chroms:
            raise ContigNotFoundError(f"Contig {contig} not found in reference.")
        if not self._is_2bit:
            from ..synth import pack_twobit
            self._ensure(contig)
            return pack_twobit(self._codes[contig], self._nmask[contig])
        buf, end, o = self._buf, self._end, self._index[contig]
        (n,) = struct.unpack_from(end + "I", buf, o); o += 4
        (nb,) = struct.unpack_from(end + "I", buf, o); o += 4
        n_starts = np.frombuffer(buf, end + "u4", nb, o); o += 4 * nb
        n_sizes = np.frombuffer(buf, end + "u4", nb, o); o += 4 * nb
        (mb,) = struct.unpack_from(end + "I", buf, o); o += 4 + 8 * mb + 4
        raw = np.frombuffer(buf, np.uint8, (n + 3) // 4, o)
        nw = (n + 15) // 16 + 2
        seq_bytes = np.zeros(nw * 4, dtype=np.uint8)
        seq_bytes[: raw.size] = _TWOBIT_BYTE_LUT[raw]
        if n % 4:                                     # bases past the end of the contig read as 0, like pack_twobit
            seq_bytes[raw.size - 1] &= (1 << (2 * (n % 4))) - 1
        nm = (n + 31) // 32 + 2
        flags = np.zeros(nm * 32, dtype=np.uint8)
        for a, z in zip(n_starts.tolist(), n_sizes.tolist()):
            flags[a: min(a + z, n)] = 1
        nmask_words = np.packbits(flags.reshape(nm, 32), axis=1, bitorder="little").view("<u4").reshape(nm)
        return seq_bytes.view("<u4"), nmask_words.astype(np.uint32)

    def _open_fasta(self) -> None:
        opener = gzip.open if self.reference_path.endswith(".gz") else open
        self._chroms = {}
        name, chunks = None, []
        lut = np.full(256, 255, np.uint8)
        for ch, v in zip(b"ACGTacgt", (0, 1, 2, 3, 0, 1, 2, 3)):
            lut[ch] = v

        def flush():
            if name is None:
                return
            raw = np.frombuffer(b"".join(chunks), np.uint8)
            c = lut[raw]
            self._nmask[name] = c == 255
            self._codes[name] = np.where(c == 255, 0, c).astype(np.uint8)
            self._chroms[name] = int(raw.size)

        with opener(self.reference_path, "rb") as fh:
            for line in fh:
                if line.startswith(b">"):
                    flush()
                    name, chunks = line[1:].split()[0].decode(), []
                else:
                    chunks.append(line.strip())
        flush()

    def _ensure(self, contig: str) -> None:
        if contig not in self._chroms:
            raise ContigNotFoundError(f"Contig {contig} not found in reference.")
        if contig not in self._codes:
            self._decode_2bit(contig)

    # -- public interface (io/reference.py:114-176) -----------------------
    @property
    def chroms(self) -> Dict[str, int]:
        return self._chroms

    def sequence(self, contig, start=None, stop=None, fail_on_excess_range: bool = True) -> str:
        if contig not in self._chroms:
            raise ContigNotFoundError(f"Contig {contig} not found in reference.")
        chrom_len = self._chroms[contig]
        start = 0 if start is None else start
        stop = chrom_len if stop is None else stop
        if start < 0 or stop > chrom_len or start > stop:
            if fail_on_excess_range:
                raise OutOfBoundsError(
                    f"Requested range {contig}:{start}-{stop} is out of bounds (0-{chrom_len}).")
            start, stop = max(0, start), min(chrom_len, stop)
            if start > stop:
                return ""
        self._ensure(contig)
        seq = _ASCII[self._codes[contig][start:stop]].copy()
        seq[self._nmask[contig][start:stop]] = ord("N")
        return seq.tobytes().decode("ascii")

    def contig_arrays(self, contig):
        """``(codes uint8 A0C1G2T3, n_mask bool)`` of a contig (host)."""
        self._ensure(contig)
        return self._codes[contig], self._nmask[contig]

    def device_contig(self, contig, device=None):
        """``PackedContig`` in HBM (packed + uploaded once per contig)."""
        from ..device import PackedContig, require_cuda
        dev = require_cuda(device)
        key = (contig, str(dev))
        if key not in self._device:
            seq_words, nmask_words = self.packed_words(contig)
            self._device[key] = PackedContig(seq_words, nmask_words, self._chroms[contig], device=dev)
        return self._device[key]

    def close(self) -> None:
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc) -> None:
        self.close()

    def __getitem__(self, contig):
        if contig not in self._chroms:
            raise ContigNotFoundError(f"Contig {contig} not found in reference.")
        return _ContigSlicer(self, contig)


class _ContigSlicer:
    def __init__(self, wrapper, contig):
        self.wrapper, self.contig = wrapper, contig

    def __getitem__(self, key):
        if isinstance(key, slice):
            return self.wrapper.sequence(self.contig, key.start, key.stop)
        if isinstance(key, int):
            return self.wrapper.sequence(self.contig, key, key + 1)
        raise TypeError("Slicer indices must be integers or slices.")

    def __len__(self):
        return self.wrapper.chroms[self.contig]
