"""End-to-end pipeline from HOST fragment columns: chunked H2D -> kernels -> D2H, overlapped.

The device kernels finish a chr1-scale shard in well under a millisecond, so the end-to-end
time of the reference-facing call is PCIe: 9 B per fragment in, 2-4 B per position out.  This
module hides as much of that as the link allows:

* the contig is cut into ``n_chunks`` position ranges; chunk c's fragment slice (plus halo) is
  copied host->device on one stream while chunk c-1 computes on a second stream and chunk c-2's
  WPS drains device->host on a third (PCIe is full duplex), with double-buffered device
  staging and CUDA events for every hand-off;
* WPS travels back as int16 (``ftk_wps_tiles_i16``) or, with ``wps_dtype="int8"``, as int8
  (``ftk_wps_tiles_i8``) - both exact: a score that does not fit raises the overflow flag and the
  caller reruns one size up - which halves / quarters the D2H bytes.  PCIe is full duplex but not
  free of interference: measured on this box 720 MB H2D takes 13.0 ms alone, 15.2 ms next to a
  500 MB D2H and 13.5 ms next to a 250 MB one (``tools/pcie_bw.py``);
* per-interval coverage counts and the pooled length histogram are computed per chunk on the
  same resident slice and come back once at the end.

Host buffers must be pinned (``torch.Tensor.pin_memory``) for the copies to be asynchronous.
"""
from __future__ import annotations

import numpy as np

from ._lib import check, lib
from .device import (ContigFragments, IntervalSet, WpsPlan, _stream_ptr, interval_hist, none_to_ftk,
                     require_cuda, torch)

__all__ = ["StreamedContig"]


class StreamedContig:
    """Streamed L-WPS (+ coverage + length histogram) of one contig from pinned host columns.

    ``ivl_start/ivl_stop``: intervals sorted by start (e.g. ``multi_wps`` windows).  Outputs:
    ``h_wps`` (int16 or int8 pinned, all intervals back to back; ``offsets`` as in ``WpsPlan``),
    ``h_cov`` (int64 per interval), ``h_hist`` (int64[n_bins]), ``h_total`` (int64[1]).
    """

    def __init__(self, h_start, h_stop, h_mapq, ivl_start, ivl_stop, chrom_size, window_size=120,
                 min_length=120, max_length=180, quality_threshold=30, max_frag_len=None, n_chunks=8,
                 device=None, coverage=True, length_hist=True, wps_dtype="int16"):
        t = torch()
        self.device = require_cuda(device)
        if wps_dtype not in ("int16", "int8"):
            raise ValueError("wps_dtype must be 'int16' or 'int8'")
        self.wps_dtype = wps_dtype
        wire = t.int16 if wps_dtype == "int16" else t.int8
        self.h_start, self.h_stop, self.h_mapq = h_start, h_stop, h_mapq
        for h in (h_start, h_stop, h_mapq):
            if not h.is_pinned():
                raise ValueError("host columns must be pinned (tensor.pin_memory())")
        self.params = (int(window_size), min_length, int(max_length), int(quality_threshold))
        st_np = h_start.numpy()
        n = st_np.shape[0]
        if max_frag_len is None:
            max_frag_len = int((h_stop.numpy().astype(np.int64) - st_np).max()) if n else 0
        self.max_frag_len = int(max_frag_len)
        s = np.asarray(ivl_start, dtype=np.int64); e = np.asarray(ivl_stop, dtype=np.int64)
        n_ivl = len(s)
        n_chunks = max(1, min(int(n_chunks), n_ivl))
        bounds = np.linspace(0, n_ivl, n_chunks + 1).astype(np.int64)
        ln = np.maximum(e - s, 0)
        self.offsets = np.zeros(n_ivl + 1, np.int64); np.cumsum(ln, out=self.offsets[1:])
        self.n_positions = int(self.offsets[-1])
        halo = max(self.max_frag_len, int(max_length) + int(window_size)) + 2
        self.chunks = []
        for c in range(n_chunks):
            i0, i1 = int(bounds[c]), int(bounds[c + 1])
            p_lo, p_hi = int(s[i0:i1].min()), int(e[i0:i1].max())
            f0 = int(np.searchsorted(st_np, p_lo - halo, side="left")) & ~15   # 64-byte aligned slice starts
            f1 = int(np.searchsorted(st_np, p_hi + halo, side="left"))
            plan = WpsPlan(s[i0:i1], e[i0:i1], int(chrom_size), int(max_length), self.device)
            self.chunks.append(dict(i0=i0, i1=i1, f0=f0, f1=f1, plan=plan, out_off=int(self.offsets[i0]),
                                    n_pos=plan.n_positions,
                                    ivl=IntervalSet(s[i0:i1].tolist(), e[i0:i1].tolist(), self.device) if coverage else None,
                                    region=IntervalSet([0 if c == 0 else p_lo], [None if c == n_chunks - 1 else p_hi],
                                                       self.device) if length_hist else None))
        max_f = max(c["f1"] - c["f0"] for c in self.chunks) + 16
        max_p = max(c["n_pos"] for c in self.chunks)
        self.d_start = [t.empty(max_f, dtype=t.int32, device=self.device) for _ in range(2)]
        self.d_stop = [t.empty(max_f, dtype=t.int32, device=self.device) for _ in range(2)]
        self.d_mapq = [t.empty(max_f, dtype=t.uint8, device=self.device) for _ in range(2)]
        self.d_out = [t.empty(max(max_p, 1), dtype=wire, device=self.device) for _ in range(2)]
        self.d_flag = t.zeros(1, dtype=t.int32, device=self.device)
        self.n_bins = self.max_frag_len + 1
        self.d_cov = t.zeros(max(n_ivl, 1), dtype=t.int64, device=self.device)
        self.d_total = t.zeros(1, dtype=t.int64, device=self.device)
        self.d_hist = t.zeros((1, self.n_bins), dtype=t.int64, device=self.device)
        self.h_wps = t.empty(max(self.n_positions, 1), dtype=wire).pin_memory()
        self.h_cov = t.empty(max(n_ivl, 1), dtype=t.int64).pin_memory()
        self.h_hist = t.empty((1, self.n_bins), dtype=t.int64).pin_memory()
        self.h_total = t.empty(1, dtype=t.int64).pin_memory()
        self.h_flag = t.zeros(1, dtype=t.int32).pin_memory()
        self.s_in, self.s_comp, self.s_out = (t.cuda.Stream(self.device) for _ in range(3))
        self.ev_in = [t.cuda.Event() for _ in range(2)]
        self.ev_comp = [t.cuda.Event() for _ in range(2)]
        self.ev_out = [t.cuda.Event() for _ in range(2)]
        self.h2d_bytes = sum((c["f1"] - c["f0"]) * 9 for c in self.chunks)
        self.d2h_bytes = self.n_positions * (2 if wps_dtype == "int16" else 1) + n_ivl * 8 + self.n_bins * 8 + 8 + 4
        self.kernel_launches = 0

    def run(self):
        """One end-to-end pass; returns after every result is in the pinned host buffers."""
        t = torch()
        W, lo, hi, q = self.params
        L = lib()
        wps_tiles = L.ftk_wps_tiles_i16 if self.wps_dtype == "int16" else L.ftk_wps_tiles_i8
        cur = t.cuda.current_stream(self.device)
        for s_ in (self.s_in, self.s_comp, self.s_out):
            s_.wait_stream(cur)
        with t.cuda.stream(self.s_comp):
            self.d_cov.zero_(); self.d_total.zero_(); self.d_hist.zero_(); self.d_flag.zero_()
        launches = 0
        for k, c in enumerate(self.chunks):
            b = k & 1
            nf = c["f1"] - c["f0"]
            with t.cuda.stream(self.s_in):
                if k >= 2:
                    self.s_in.wait_event(self.ev_comp[b])        # staging buffer b consumed
                self.d_start[b][:nf].copy_(self.h_start[c["f0"]:c["f1"]], non_blocking=True)
                self.d_stop[b][:nf].copy_(self.h_stop[c["f0"]:c["f1"]], non_blocking=True)
                self.d_mapq[b][:nf].copy_(self.h_mapq[c["f0"]:c["f1"]], non_blocking=True)
                self.ev_in[b].record(self.s_in)
            with t.cuda.stream(self.s_comp):
                self.s_comp.wait_event(self.ev_in[b])
                if k >= 2:
                    self.s_comp.wait_event(self.ev_out[b])       # output buffer b drained
                frags = ContigFragments(self.d_start[b][:nf], self.d_stop[b][:nf], self.d_mapq[b][:nf], None,
                                        device=self.device, max_len=self.max_frag_len)
                plan = c["plan"]
                if plan.n_tiles:
                    check(wps_tiles(
                        frags.start.data_ptr(), frags.stop.data_ptr(), frags.mapq.data_ptr(), frags.n,
                        plan.tile_p0.data_ptr(), plan.tile_len.data_ptr(), plan.tile_mid_lo.data_ptr(),
                        plan.tile_mid_hi.data_ptr(), plan.tile_out_off.data_ptr(), plan.n_tiles,
                        W, none_to_ftk(lo), hi, q, 0, plan.scratch.data_ptr(), self.d_out[b].data_ptr(),
                        self.d_flag.data_ptr(), _stream_ptr(self.device)), "ftk_wps_tiles_" + self.wps_dtype[3:])
                    launches += 2
                if c["ivl"] is not None:
                    interval_hist(frags, intersect_policy="midpoint", quality_threshold=q, ivl_set=c["ivl"],
                                  out=(self.d_cov[c["i0"]:c["i1"]], None, None))
                    launches += 2
                if c["region"] is not None:
                    interval_hist(frags, intersect_policy="midpoint", quality_threshold=q, n_bins=self.n_bins,
                                  pooled=True, ivl_set=c["region"], out=(self.d_total, self.d_hist, None))
                    launches += 2
                self.ev_comp[b].record(self.s_comp)
            with t.cuda.stream(self.s_out):
                self.s_out.wait_event(self.ev_comp[b])
                if c["n_pos"]:
                    self.h_wps[c["out_off"]: c["out_off"] + c["n_pos"]].copy_(self.d_out[b][:c["n_pos"]], non_blocking=True)
                self.ev_out[b].record(self.s_out)
        with t.cuda.stream(self.s_out):
            self.s_out.wait_stream(self.s_comp)
            self.h_cov.copy_(self.d_cov, non_blocking=True)
            self.h_hist.copy_(self.d_hist, non_blocking=True)
            self.h_total.copy_(self.d_total, non_blocking=True)
            self.h_flag.copy_(self.d_flag, non_blocking=True)
        cur.wait_stream(self.s_out)
        cur.wait_stream(self.s_in)
        cur.synchronize()
        self.kernel_launches = launches
        if int(self.h_flag[0]):
            raise OverflowError(f"WPS does not fit {self.wps_dtype} on this input; rerun with "
                                + ("wps_dtype='int16'" if self.wps_dtype == "int8" else "the int32 path (WpsPlan.run)"))
        return self.h_wps, self.h_cov, self.h_hist, self.h_total
