"""Device-resident fragment columns and the CUDA op wrappers.

PyTorch is plumbing here (HBM allocations, streams); every computation is a
call into ``libftk_b200.so`` through the C ABI of ``include/ftk_b200.h``.
"""
from __future__ import annotations

from ctypes import POINTER, c_int32, c_int64

import numpy as np

from ._lib import FTK_NONE, FtkLibraryError, check, lib

__all__ = ["ContigFragments", "WpsPlan", "require_cuda", "none_to_ftk"]

_torch = None


def torch():
    global _torch
    if _torch is None:
        import torch as t
        _torch = t
    return _torch


def require_cuda(device=None):
    """Return a CUDA ``torch.device`` or raise: there is no CPU compute path."""
    t = torch()
    if not t.cuda.is_available():
        raise FtkLibraryError(
            "finaletoolkit_b200 needs a CUDA device (sm_100a); no CPU fallback exists")
    lib()  # fail loudly if the extension is missing
    if device is None:
        return t.device("cuda", t.cuda.current_device())
    return t.device(device)


def none_to_ftk(v) -> int:
    return FTK_NONE if v is None else int(v)


def _stream_ptr(device) -> int:
    return int(torch().cuda.current_stream(device).cuda_stream)


def _np_ptr(a, ctype):
    return a.ctypes.data_as(POINTER(ctype))


def _to_device(a: np.ndarray, device, dtype):
    """H2D through pinned staging (async on the current stream)."""
    t = torch()
    a = np.ascontiguousarray(a, dtype=dtype)
    host = t.from_numpy(a)
    if a.nbytes >= (1 << 16):
        host = host.pin_memory()
    return host.to(device, non_blocking=True)


class ContigFragments:
    """Start-sorted fragments of one contig as columns in HBM.

    ``start``/``stop`` int32, ``mapq``/``strand`` uint8 (strand 1 = '+'), the
    columnar form of the reference's fragment stream
    (utils/_frag_generator.py:124-130 yields contig,start,stop,mapq,is_forward).
    """

    def __init__(self, start, stop, mapq=None, strand=None, device=None, contig=None,
                 max_len=None):
        t = torch()
        self.device = require_cuda(device)
        self.contig = contig
        if isinstance(start, np.ndarray) or not t.is_tensor(start):
            start = np.asarray(start)
            stop = np.asarray(stop)
            if start.size and not np.all(start[1:] >= start[:-1]):
                order = np.argsort(start, kind="stable")
                start, stop = start[order], stop[order]
                mapq = None if mapq is None else np.asarray(mapq)[order]
                strand = None if strand is None else np.asarray(strand)[order]
            if max_len is None:
                max_len = int((stop.astype(np.int64) - start).max()) if start.size else 0
            self.start = _to_device(start, self.device, np.int32)
            self.stop = _to_device(stop, self.device, np.int32)
            self.mapq = None if mapq is None else _to_device(mapq, self.device, np.uint8)
            self.strand = None if strand is None else _to_device(strand, self.device, np.uint8)
        else:
            self.start = start.to(self.device, t.int32).contiguous()
            self.stop = stop.to(self.device, t.int32).contiguous()
            self.mapq = None if mapq is None else mapq.to(self.device, t.uint8).contiguous()
            self.strand = None if strand is None else strand.to(self.device, t.uint8).contiguous()
            if max_len is None:
                max_len = int((self.stop - self.start).max().item()) if self.start.numel() else 0
        self.n = int(self.start.numel())
        self.max_len = max(int(max_len), 0)

    def ptrs(self):
        return (self.start.data_ptr(), self.stop.data_ptr(),
                0 if self.mapq is None else self.mapq.data_ptr())


class WpsPlan:
    """Tile table for a set of intervals of one contig (host-planned, HBM-resident).

    Mirrors what ``multi_wps`` hands to ``wps`` per interval
    (frag/_multi_wps.py:176-191): (start, stop, chrom_size) plus the padded
    fetch window of frag/_wps.py:156-157, expanded into <= FTK_WPS_TILE tiles.
    """

    def __init__(self, ivl_start, ivl_stop, chrom_size: int, max_length: int, device=None):
        t = torch()
        self.device = require_cuda(device)
        s = np.ascontiguousarray(ivl_start, dtype=np.int64)
        e = np.ascontiguousarray(ivl_stop, dtype=np.int64)
        ln = np.maximum(e - s, 0)
        self.offsets = np.zeros(len(s) + 1, dtype=np.int64)
        np.cumsum(ln, out=self.offsets[1:])
        self.n_positions = int(self.offsets[-1])
        self.max_length = int(max_length)
        L = lib()
        i64p, i32p = POINTER(c_int64), POINTER(c_int32)
        null32, null64 = i32p(), i64p()
        n_tiles = L.ftk_wps_plan_tiles(_np_ptr(s, c_int64), _np_ptr(e, c_int64), _np_ptr(self.offsets, c_int64),
                                       len(s), int(chrom_size), int(max_length), null32, null32, null32, null32, null64)
        check(n_tiles, "ftk_wps_plan_tiles")
        self.n_tiles = int(n_tiles)
        p0 = np.empty(self.n_tiles, np.int32)
        tl = np.empty(self.n_tiles, np.int32)
        mlo = np.empty(self.n_tiles, np.int32)
        mhi = np.empty(self.n_tiles, np.int32)
        off = np.empty(self.n_tiles, np.int64)
        if self.n_tiles:
            check(L.ftk_wps_plan_tiles(_np_ptr(s, c_int64), _np_ptr(e, c_int64), _np_ptr(self.offsets, c_int64),
                                       len(s), int(chrom_size), int(max_length), _np_ptr(p0, c_int32), _np_ptr(tl, c_int32),
                                       _np_ptr(mlo, c_int32), _np_ptr(mhi, c_int32), _np_ptr(off, c_int64)),
                  "ftk_wps_plan_tiles")
        self.tile_p0 = _to_device(p0, self.device, np.int32)
        self.tile_len = _to_device(tl, self.device, np.int32)
        self.tile_mid_lo = _to_device(mlo, self.device, np.int32)
        self.tile_mid_hi = _to_device(mhi, self.device, np.int32)
        self.tile_out_off = _to_device(off, self.device, np.int64)
        self.scratch = t.empty(2 * max(self.n_tiles, 1), dtype=t.int64, device=self.device)

    def run(self, frags: ContigFragments, window_size=120, min_length=120, max_length=180,
            quality_threshold=30, out=None):
        """Launch the WPS kernels on the current stream; returns int32[n_positions] (device)."""
        t = torch()
        if int(max_length) != self.max_length:
            raise ValueError("plan was built for a different max_length")
        if out is None:
            out = t.empty(self.n_positions, dtype=t.int32, device=self.device)
        elif out.dtype != t.int32 or out.numel() < self.n_positions or not out.is_contiguous():
            raise ValueError("out must be a contiguous int32 tensor with n_positions elements")
        if self.n_tiles == 0:
            return out
        fs, fe, mq = frags.ptrs()
        check(lib().ftk_wps_tiles_i32(
            fs, fe, mq, frags.n,
            self.tile_p0.data_ptr(), self.tile_len.data_ptr(), self.tile_mid_lo.data_ptr(),
            self.tile_mid_hi.data_ptr(), self.tile_out_off.data_ptr(), self.n_tiles,
            int(window_size), none_to_ftk(min_length), int(max_length), int(quality_threshold),
            self.scratch.data_ptr(), out.data_ptr(), _stream_ptr(self.device)), "ftk_wps_tiles_i32")
        return out
