#!/usr/bin/env python
"""Time the UNMODIFIED reference (TEST INFRASTRUCTURE; build container only, needs /root/reference).

    python oracle/time_reference.py [n_intervals] [workers]

Runs the reference's own ``finaletoolkit.frag.multi_wps`` (imported from /root/reference/src, with the
in-memory stand-ins of ``oracle/fakes`` for pysam / pyBigWig / py2bit, which are not installable here)
with ``workers = os.cpu_count()`` over a seeded sample of 5-kb intervals of a synthetic 30x contig, and
the oracle's OpenMP port on the same intervals, so that the "port vs real reference" ratio quoted in
BASELINE.md is a measurement.  The fake TabixFile serves the fragments from memory (registered by
path, inherited by the Pool's forked workers), so BGZF/tabix decode cost is excluded - this favours
the reference.  Writes profiles/r2_reference_timing.json.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = "/root/reference"
sys.path[:0] = [os.path.join(HERE, "fakes"), os.path.join(REF, "src"), REPO]

import pysam  # noqa: E402  (the fake)
import finaletoolkit.frag as F  # noqa: E402  (the reference)

from finaletoolkit_b200.synth import synth_fragments  # noqa: E402
from oracle import oracle as O  # noqa: E402

warnings.simplefilter("ignore")
_REGISTRY: dict = {}
_orig_init = pysam.TabixFile.__init__


def _init(self, path=None, *a, **k):
    """Serve a registered path from the prebuilt in-memory rows instead of parsing text."""
    src = _REGISTRY.get(str(path)) if path is not None else None
    if src is None:
        return _orig_init(self, path, *a, **k)
    self._rows, self._starts, self._maxlen = src._rows, src._starts, src._maxlen
    self.contigs, self.filename = src.contigs, str(path)


pysam.TabixFile.__init__ = _init


def main():
    n_ivl = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    workers = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
    clen = n_ivl * 5000
    n = int(round(clen * 80_000_000 / 249_250_621))          # the chr1-scale density of the benchmark
    st, sp, mq, sd = synth_fragments(clen, n, 0)
    tmp = tempfile.mkdtemp(prefix="ftk_reftime_")
    path = os.path.join(tmp, "synth.frag.gz")
    for p in (path, path + ".tbi"):
        open(p, "wb").close()                                 # the reference only checks that they exist
    _REGISTRY[path] = pysam.TabixFile.from_columns({"1": (st, sp, mq, sd)})
    bed = os.path.join(tmp, "sites.bed")
    with open(bed, "w") as fh:
        for a in range(0, clen, 5000):
            fh.write(f"1\t{a + 2499}\t{a + 2501}\t.\t0\t+\n")
    cs = os.path.join(tmp, "cs")
    open(cs, "w").write(f"1\t{clen}\n")
    out = os.path.join(tmp, "wps.bed.gz")
    # warm numba (the reference's _single_nt_wps is @jit) on a tiny call, outside the timing
    F.wps(path, "1", 0, 200, clen)
    t0 = time.perf_counter()
    F.multi_wps(path, bed, chrom_sizes=cs, output_file=out, window_size=120, interval_size=5000, min_length=120,
                max_length=180, quality_threshold=30, workers=workers)
    t_ref = time.perf_counter() - t0
    pos = clen
    # the oracle port on the same intervals
    fr = O.Frags(st, sp, mq, sd)
    s = np.arange(0, clen, 5000, dtype=np.int64); e = np.minimum(s + 5000, clen)
    t0 = time.perf_counter()
    exp, off = O.wps_intervals(fr, s, e, clen, 120, 120, 180, 30, threads=workers)
    t_port = time.perf_counter() - t0
    # same answer?  (reference output text -> scores)
    import gzip
    got = np.array([int(line.rsplit("\t", 1)[1]) for line in gzip.open(out, "rt")], dtype=np.int64)
    same = bool(got.shape == exp.shape and np.array_equal(got, exp))
    res = {"what": "unmodified reference finaletoolkit.frag.multi_wps (pure Python + numba) vs the oracle's OpenMP port, "
                   "same intervals, same fragments, in-memory fragment source (no BGZF/tabix decode)",
           "host": "build container", "cores": os.cpu_count(), "workers": workers,
           "intervals": n_ivl, "positions": pos, "fragments": n,
           "reference_seconds": t_ref, "reference_positions_per_sec": pos / t_ref,
           "reference_positions_per_sec_per_core": pos / t_ref / workers,
           "port_seconds": t_port, "port_positions_per_sec": pos / t_port,
           "port_over_reference": t_ref / t_port, "outputs_identical": same}
    os.makedirs(os.path.join(REPO, "profiles"), exist_ok=True)
    with open(os.path.join(REPO, "profiles", "r2_reference_timing.json"), "w") as fh:
        json.dump(res, fh, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
