#!/usr/bin/env python
"""Generate tests/golden/* by running the UNMODIFIED reference (TEST INFRASTRUCTURE).

Runs only in the build container, where /root/reference exists:

    python oracle/make_golden.py

The reference package (pure Python) is imported from /root/reference/src with
in-memory stand-ins for pysam / pyBigWig / py2bit / loess (``oracle/fakes``)
because those C extensions are not installable here.  Every array or string
written below is an output of the reference's own functions; the inputs that
produced it are stored next to it so the GPU box (which has no
/root/reference) can replay them through the CUDA path and the oracle.

Outputs: tests/golden/fixture17.npz, synth_small.npz, motif.npz, adjust.npz,
cleavage.npz, delfi.npz, agg.npz, manifest.json.
"""
from __future__ import annotations

import gzip
import hashlib
import io
import json
import os
import struct
import sys
import tempfile
import warnings
from contextlib import redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = "/root/reference"
sys.path[:0] = [os.path.join(HERE, "fakes"), os.path.join(REF, "src"), REPO]

import pysam  # noqa: E402  (the fake)
import pyBigWig  # noqa: E402  (the fake)
import finaletoolkit.frag as F  # noqa: E402  (the reference)
import finaletoolkit.utils as U  # noqa: E402
from finaletoolkit.frag._adjust_wps import _local_filter  # noqa: E402
from scipy.signal import savgol_filter  # noqa: E402

from finaletoolkit_b200.synth import synth_fragments, synth_twobit  # noqa: E402

warnings.simplefilter("ignore")
OUT = os.path.join(REPO, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
DATA = os.path.join(REF, "tests", "data")
TMP = tempfile.mkdtemp(prefix="ftk_golden_")

manifest: dict = {"generator": "oracle/make_golden.py", "reference": "epifluidlab/FinaleToolkit v1.1.0 tree"}


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def read_gz_text(path):
    with gzip.open(path, "rt") as fh:
        return fh.read()


def columns_of(path):
    st, sp, mq, fw, ct = [], [], [], [], []
    with gzip.open(path, "rt") as fh:
        for line in fh:
            f = line.rstrip("\n").split("\t")
            ct.append(f[0]); st.append(int(f[1])); sp.append(int(f[2]))
            mq.append(int(f[3])); fw.append(1 if "+" in f[4] else 0)
    return ct, np.array(st, np.int32), np.array(sp, np.int32), np.array(mq, np.uint8), np.array(fw, np.uint8)


# ----------------------------------------------------------------------------
# A. the reference's own 17-fragment fixture
# ----------------------------------------------------------------------------
def fixture17():
    frag = os.path.join(DATA, "12.3444.b37.frag.gz")
    bed6 = os.path.join(DATA, "12.3444.b37.frag.bed.gz")
    ivl = os.path.join(DATA, "intervals.bed")
    ivl_ov = os.path.join(DATA, "intervals_overlapped.bed")
    cs = os.path.join(DATA, "b37.chrom.sizes")
    ct, st, sp, mq, fw = columns_of(frag)
    arrays = dict(start=st, stop=sp, mapq=mq, strand=fw)
    # the fixture files themselves (bgzip + tabix output of htslib): exercises the BGZF / .tbi readers
    for key, path in (("frag_gz_file", frag), ("frag_gz_tbi_file", frag + ".tbi"),
                      ("bed_gz_file", bed6), ("bed_gz_tbi_file", bed6 + ".tbi")):
        arrays[key] = np.frombuffer(open(path, "rb").read(), np.uint8)
    # the same 17 fragments as a BAM (tests/data/12.3444.b37.bam): what the reference's own
    # AlignmentWrapper._fetch_sam yields from it (read filter + fragment reconstruction)
    bam = os.path.join(DATA, "12.3444.b37.bam")
    arrays["bam_file"] = np.frombuffer(open(bam, "rb").read(), np.uint8)
    from finaletoolkit.io.alignment import AlignmentWrapper
    with AlignmentWrapper(bam, quality_threshold=0) as wrapper:
        frs = list(wrapper.fetch())
        bam_chroms = dict(wrapper._chroms)
    arrays["bam_start"] = np.array([f.start for f in frs], np.int64); arrays["bam_stop"] = np.array([f.stop for f in frs], np.int64)
    arrays["bam_mapq"] = np.array([f.mapq for f in frs], np.int64); arrays["bam_strand"] = np.array([f.is_forward for f in frs], np.uint8)
    arrays["bam_wps"] = F.wps(bam, "12", 34442500, 34447500, 133851895)["wps"].astype(np.int64)
    m = {"contig": "12", "chrom_size": 133851895,
         "bam_contigs": sorted({f.contig for f in frs}), "bam_n_refs": len(bam_chroms), "bam_ref_12": bam_chroms["12"],
         "bam_single_coverage": F.single_coverage(bam, "12", 34442500, 34447500, quality_threshold=30),
         "bam_frag_length_bins": [np.asarray(x).tolist() for x in F.frag_length_bins(bam, "12", 34442500, 34447500, bin_size=10)],
         "frag_gz_text": read_gz_text(frag), "frag_bed_gz_text": read_gz_text(bed6),
         "intervals_bed": open(ivl).read(), "intervals_overlapped_bed": open(ivl_ov).read(),
         "chrom_sizes": open(cs).read()}

    # --- wps known answers (reference tests/test_wps.py:18-26 + extra params)
    wps_cases = []
    for i, (s, e, kw) in enumerate([
        (34444145, 34444155, dict(quality_threshold=0)),
        (34442500, 34447500, dict()),
        (34442500, 34447500, dict(window_size=121)),
        (34442500, 34447500, dict(window_size=60, min_length=35, max_length=80)),
        (34442500, 34447500, dict(window_size=200, min_length=120, max_length=180)),
        (34442500, 34447500, dict(window_size=120, min_length=50, max_length=400, quality_threshold=0)),
        (34443000, 34446700, dict(window_size=7, min_length=1, max_length=1000, quality_threshold=20)),
        (34443200, 34443200, dict()),
    ]):
        r = F.wps(frag, "12", s, e, 133851895, **kw)
        arrays[f"wps_{i}"] = r["wps"].astype(np.int64)
        wps_cases.append(dict(start=s, stop=e, kwargs=kw, n=int(r.shape[0]),
                              sum=int(r["wps"].sum()),
                              sha256_i32=sha(r["wps"].astype("<i4"))))
    m["wps_cases"] = wps_cases

    # --- multi_wps -> .bed.gz (config 1) and with overlapped BED
    mw = []
    for name, bed, kw in [("cfg1", ivl, dict()),
                          ("overlapped", ivl_ov, dict(interval_size=400, window_size=60, min_length=35, max_length=300, quality_threshold=0))]:
        out = os.path.join(TMP, f"mw_{name}.bed.gz")
        try:
            F.multi_wps(frag, bed, chrom_sizes=cs, output_file=out, workers=1, **kw)
        except Exception as e:  # noqa: BLE001
            mw.append(dict(name=name, error=type(e).__name__, kwargs=kw))
            continue
        txt = read_gz_text(out)
        rows = [ln.split("\t") for ln in txt.splitlines()]
        arrays[f"mwps_{name}_pos"] = np.array([int(r[1]) for r in rows], np.int64)
        arrays[f"mwps_{name}_score"] = np.array([int(r[3]) for r in rows], np.int64)
        mw.append(dict(name=name, kwargs=kw, n_lines=len(rows), first=txt.splitlines()[0] if rows else "",
                       last=txt.splitlines()[-1] if rows else "", sha256_text=hashlib.sha256(txt.encode()).hexdigest()))
    m["multi_wps"] = mw
    # multi_wps -> fake bigWig, then adjust_wps over the same BED
    bw = os.path.join(TMP, "cfg1.bw")
    F.multi_wps(frag, ivl, chrom_sizes=cs, output_file=bw, workers=1)
    ent = pyBigWig._STORE[bw]["data"]["12"]
    arrays["bw_cfg1_pos"] = np.concatenate([e[0] for e in ent])
    arrays["bw_cfg1_val"] = np.concatenate([e[2] for e in ent]).astype(np.float32)
    adj_cases = []
    for j, kw in enumerate([dict(), dict(mean=True), dict(subtract_edges=True, edge_size=200),
                            dict(savgol=False), dict(median_window_size=500, savgol_window_size=11, savgol_poly_deg=3)]):
        # a 1-row BED so the interval has enough raw WPS for the median window
        bed1 = os.path.join(TMP, f"adj_{j}.bed")
        with open(bed1, "w") as fh:
            fh.write("12\t34445500\t34445582\n")
        obw = os.path.join(TMP, f"adj_{j}.bw")
        F.adjust_wps(bw, bed1, obw, cs, workers=1, **kw)
        d = pyBigWig._STORE[obw]["data"].get("12", [])
        pos = np.concatenate([e[0] for e in d]) if d else np.zeros(0, np.int64)
        val = np.concatenate([e[2] for e in d]) if d else np.zeros(0, np.float32)
        arrays[f"adj_{j}_pos"] = pos
        arrays[f"adj_{j}_val_f32"] = val
        adj_cases.append(dict(kwargs=kw, bed="12\t34445500\t34445582\n", n=int(pos.shape[0])))
    m["adjust_wps_cases"] = adj_cases

    # --- coverage (reference tests/test_coverage.py)
    cov = []
    for args in [dict(), dict(contig="12"), dict(contig="12", start=34443118, stop=34443538),
                 dict(contig="12", start=34443118, stop=34443538, intersect_policy="any"),
                 dict(contig="12", start=34443118, stop=34443538, min_length=150, max_length=170, quality_threshold=0),
                 dict(contig="12", start=34445000, stop=None, quality_threshold=0)]:
        r = F.single_coverage(frag, **args)
        cov.append(dict(kwargs=args, result=list(r)))
    m["single_coverage"] = cov
    covs = []
    for kw, suffix in [(dict(), ".bed"), (dict(normalize=True), ".bed"), (dict(normalize=True, scale_factor=1e6, intersect_policy="any"), ".bedgraph"),
                       (dict(min_length=100, max_length=170, quality_threshold=0, normalize=True), ".bed.gz")]:
        out = os.path.join(TMP, "cov" + suffix)
        r = F.coverage(frag, ivl, out, workers=1, **kw)
        txt = read_gz_text(out) if suffix.endswith(".gz") else open(out).read()
        covs.append(dict(kwargs=kw, suffix=suffix, results=[list(x) for x in r], text=txt))
    m["coverage"] = covs

    # --- frag_length*, frag_array, frag_generator
    fl = []
    for args in [dict(contig="12", start=34443119, stop=34443538), dict(), dict(contig="12", intersect_policy="any", start=34443300, stop=34445000, quality_threshold=0)]:
        r = F.frag_length(frag, **args)
        fl.append(dict(kwargs=args, lengths=r.tolist()))
    m["frag_length"] = fl
    flb = []
    for args in [dict(contig="12", start=34443119, stop=34443538), dict(), dict(bin_size=5, summary_stats=True, short_fraction=150),
                 dict(min_length=100, max_length=170, bin_size=10, summary_stats=True), dict(contig="12", start=1, stop=2)]:
        out = os.path.join(TMP, "flb.tsv")
        if os.path.exists(out):
            os.remove(out)
        bins, counts = F.frag_length_bins(frag, output_file=out, **args)
        flb.append(dict(kwargs=args, bins=np.asarray(bins).tolist(), counts=np.asarray(counts).tolist(),
                        text=open(out).read() if os.path.exists(out) else None))
    m["frag_length_bins"] = flb
    fli = []
    for kw in [dict(), dict(short_reads=160, intersect_policy="any", quality_threshold=0), dict(min_length=200, max_length=300)]:
        out = os.path.join(TMP, "fli.bed")
        r = F.frag_length_intervals(frag, ivl, output_file=out, workers=1, **kw)
        fli.append(dict(kwargs=kw, results=[list(x) for x in r], text=open(out).read()))
    m["frag_length_intervals"] = fli
    fa = U.frag_array(frag, "12", min_length=120, max_length=180)
    m["frag_array_120_180"] = [[int(a), int(b), bool(c)] for a, b, c in fa.tolist()]
    m["frag_generator_detail"] = [list(x) for x in U.frag_generator(frag, "12", start=34443119, stop=34443538)]
    m["frag_generator_bed6"] = [list(x) for x in U.frag_generator(bed6, "12", start=34443119, stop=34443538)]
    m["frag_generator_count_all"] = sum(1 for _ in U.frag_generator(frag, None, quality_threshold=0))

    # --- regional MDS from the reference's golden TSV (tests/test_end_motifs.py:170-247)
    tsv = os.path.join(DATA, "end_motifs", "end_motifs_intervals_dif.tsv")
    emi = F.EndMotifsIntervals.from_file(tsv, 30, sep="\t")
    m["end_motifs_intervals_dif_tsv"] = open(tsv).read()
    m["end_motifs_dif_tsv"] = open(os.path.join(DATA, "end_motifs", "end_motifs_dif.tsv")).read()
    m["regional_mds"] = [[list(iv), v] for iv, v in emi.motif_diversity_score()]
    m["regional_mds_mm"] = [[list(iv), v] for iv, v in emi.motif_diversity_score(miller_madow=True)]
    emf = F.EndMotifFreqs.from_file(os.path.join(DATA, "end_motifs", "end_motifs_dif.tsv"), 30)
    m["mds_from_dif_tsv"] = emf.motif_diversity_score()
    np.savez_compressed(os.path.join(OUT, "fixture17.npz"), **arrays)
    manifest["fixture17"] = m


# ----------------------------------------------------------------------------
# B. seeded synthetic small genome (WPS / coverage / lengths)
# ----------------------------------------------------------------------------
def synth_small():
    contigs = [("chrA", 200_000, 36_000), ("chrB", 80_000, 9_000), ("chrC", 4_000, 300)]
    cols, arrays = {}, {}
    for idx, (name, ln, n) in enumerate(contigs):
        st, sp, mq, fw = synth_fragments(ln, n, idx, seed_base=7000)
        # sprinkle quirk cases: duplicates, tiny and huge fragments, contig ends
        if name == "chrA":
            st[100:140] = st[100]; sp[100:140] = st[100] + np.arange(120, 160)
            st[200:260] = st[200]; sp[200:260] = st[200] + 167
            sp[300:310] = st[300:310] + np.array([1, 2, 3, 59, 60, 61, 119, 120, 121, 122])
            order = np.argsort(st, kind="stable")
            st, sp, mq, fw = st[order], sp[order], mq[order], fw[order]
        cols[name] = (st, sp, mq, fw)
        for k, v in zip(("start", "stop", "mapq", "strand"), (st, sp, mq, fw)):
            arrays[f"{name}_{k}"] = v
    tbx = pysam.TabixFile.from_columns(cols)
    cs_path = os.path.join(TMP, "synth.chrom.sizes")
    with open(cs_path, "w") as fh:
        for name, ln, _ in contigs:
            fh.write(f"{name}\t{ln}\n")
    m = {"contigs": [[c, l] for c, l, _ in contigs]}

    rng = np.random.Generator(np.random.PCG64(4242))
    # WPS per-interval cases
    wps_cases = []
    param_sets = [dict(), dict(window_size=121), dict(window_size=60, min_length=35, max_length=80),
                  dict(window_size=200), dict(window_size=120, min_length=30, max_length=600, quality_threshold=0),
                  dict(window_size=2, min_length=1, max_length=50, quality_threshold=0),
                  dict(window_size=1, min_length=30, max_length=200)]
    ivls = [("chrA", 0, 5000), ("chrA", 195_000, 200_000), ("chrA", 60_000, 65_000), ("chrA", 1000, 1003),
            ("chrB", 77_000, 80_000), ("chrC", 0, 4000), ("chrB", 0, 137)]
    for _ in range(5):
        c = "chrA"; s = int(rng.integers(0, 190_000)); e = s + int(rng.integers(1, 9000))
        ivls.append((c, s, min(e, 200_000)))
    k = 0
    sizes = dict((c, l) for c, l, _ in contigs)
    for pi, kw in enumerate(param_sets):
        for (c, s, e) in (ivls if pi < 2 else ivls[:7:2] + ivls[7:9]):
            r = F.wps(tbx, c, s, e, sizes[c], **kw)
            arrays[f"wps_{k}"] = r["wps"].astype(np.int32)
            wps_cases.append(dict(contig=c, start=s, stop=e, kwargs=kw, key=f"wps_{k}"))
            k += 1
    m["wps_cases"] = wps_cases

    # multi_wps over a site BED with overlaps / out-of-order contigs / unknown contig
    sites = []
    for c, ln in (("chrB", 80_000), ("chrA", 200_000)):
        pts = np.sort(rng.integers(0, ln, 14))
        for p in pts.tolist():
            sites.append((c, max(p - 50, 0), p + 50))
    sites.insert(3, ("chrUn", 5, 10))
    sites.append(("chrC", 3900, 4000))
    bed_txt = "".join(f"{c}\t{s}\t{e}\n" for c, s, e in sites)
    bed = os.path.join(TMP, "sites.bed")
    open(bed, "w").write(bed_txt)
    mws = []
    for j, kw in enumerate([dict(), dict(interval_size=2000, window_size=121), dict(interval_size=600, window_size=60, min_length=35, max_length=80)]):
        out = os.path.join(TMP, f"smw_{j}.bed.gz")
        F.multi_wps(tbx, bed, chrom_sizes=cs_path, output_file=out, workers=1, **kw)
        txt = read_gz_text(out)
        rows = [ln.split("\t") for ln in txt.splitlines()]
        arrays[f"mwps_{j}_pos"] = np.array([int(r[1]) for r in rows], np.int64)
        arrays[f"mwps_{j}_score"] = np.array([int(r[3]) for r in rows], np.int32)
        # contig runs (contig name, count) to rebuild the text
        runs, prev = [], None
        for r in rows:
            if r[0] != prev:
                runs.append([r[0], 0]); prev = r[0]
            runs[-1][1] += 1
        mws.append(dict(kwargs=kw, runs=runs, n_lines=len(rows), sha256_text=hashlib.sha256(txt.encode()).hexdigest()))
    m["multi_wps"] = mws
    m["sites_bed"] = bed_txt

    # coverage: random (overlapping) intervals, both policies
    civ = []
    for i in range(40):
        c = ("chrA", "chrB")[i % 2]; ln = sizes[c]
        s = int(rng.integers(0, ln - 10)); e = s + int(rng.integers(1, 30_000))
        civ.append((c, s, min(e, ln + 500), f"iv{i}" if i % 3 else "."))
    civ += [("chrA", 0, 200_000, "whole"), ("chrC", 0, 4000, "c"), ("chrA", 500, 500, "empty")]
    ivbed_txt = "".join(f"{c}\t{s}\t{e}\t{n}\n" for c, s, e, n in civ)
    ivbed = os.path.join(TMP, "cov_iv.bed")
    open(ivbed, "w").write("# comment\ntrack name=x\n\n" + ivbed_txt)
    m["cov_intervals_bed"] = open(ivbed).read()
    covs = []
    for kw in [dict(), dict(intersect_policy="any"), dict(min_length=120, max_length=180, normalize=True),
               dict(min_length=100, max_length=None, intersect_policy="any", normalize=True, scale_factor=1e6, quality_threshold=0)]:
        out = os.path.join(TMP, "scov.bed")
        r = F.coverage(tbx, ivbed, out, workers=1, **kw)
        covs.append(dict(kwargs=kw, results=[list(x) for x in r], text=open(out).read()))
    m["coverage"] = covs
    m["single_coverage_genome"] = [list(F.single_coverage(tbx)), list(F.single_coverage(tbx, quality_threshold=0, min_length=200))]

    # fragment lengths
    flb = []
    for args in [dict(), dict(contig="chrA"), dict(contig="chrA", start=50_000, stop=90_000, bin_size=5, summary_stats=True, short_fraction=150),
                 dict(contig="chrB", intersect_policy="any", start=100, stop=40_000, min_length=100, max_length=220, bin_size=7, summary_stats=True),
                 dict(quality_threshold=0, min_length=None, max_length=None, bin_size=50, summary_stats=True, short_fraction=100)]:
        out = os.path.join(TMP, "sflb.tsv")
        bins, counts = F.frag_length_bins(tbx, output_file=out, **args)
        flb.append(dict(kwargs=args, bins=np.asarray(bins).tolist(), counts=np.asarray(counts).tolist(), text=open(out).read()))
    m["frag_length_bins"] = flb
    fl = []
    for j, args in enumerate([dict(contig="chrC"), dict(contig="chrA", start=10_000, stop=12_000, intersect_policy="any", quality_threshold=10), dict()]):
        r = F.frag_length(tbx, **args)
        arrays[f"frag_length_{j}"] = r
        fl.append(dict(kwargs=args, key=f"frag_length_{j}"))
    m["frag_length"] = fl
    fli = []
    for kw in [dict(), dict(intersect_policy="any", short_reads=120, min_length=50, max_length=400, quality_threshold=0)]:
        out = os.path.join(TMP, "sfli.bed")
        r = F.frag_length_intervals(tbx, ivbed, output_file=out, workers=1, **kw)
        fli.append(dict(kwargs=kw, results=[list(x) for x in r], text=open(out).read()))
    m["frag_length_intervals"] = fli
    np.savez_compressed(os.path.join(OUT, "synth_small.npz"), **arrays)
    manifest["synth_small"] = m
    return tbx, cols, cs_path


# ----------------------------------------------------------------------------
# C. end motifs on a synthetic 2bit
# ----------------------------------------------------------------------------
def write_2bit(path, seqs):
    """seqs: list of (name, codes A0C1G2T3 uint8, n_mask bool). UCSC codes T0 C1 A2 G3."""
    remap = np.array([2, 1, 3, 0], np.uint8)
    recs = []
    for name, codes, nm in seqs:
        n = codes.shape[0]
        u = remap[codes].copy()
        u[nm] = 0
        pad = np.zeros((-n) % 4, np.uint8)
        q = np.concatenate([u, pad]).reshape(-1, 4)
        packed = ((q[:, 0] << 6) | (q[:, 1] << 4) | (q[:, 2] << 2) | q[:, 3]).astype(np.uint8)
        d = np.diff(np.concatenate([[0], nm.astype(np.int8), [0]]))
        starts = np.flatnonzero(d == 1); ends = np.flatnonzero(d == -1)
        body = struct.pack("<I", n) + struct.pack("<I", len(starts))
        body += starts.astype("<u4").tobytes() + (ends - starts).astype("<u4").tobytes()
        body += struct.pack("<I", 0) + struct.pack("<I", 0) + packed.tobytes()
        recs.append((name, body))
    hdr = struct.pack("<IIII", 0x1A412743, 0, len(recs), 0)
    idx_size = sum(1 + len(n.encode()) + 4 for n, _ in recs)
    off = 16 + idx_size
    idx = b""
    for name, body in recs:
        idx += bytes([len(name.encode())]) + name.encode() + struct.pack("<I", off)
        off += len(body)
    with open(path, "wb") as fh:
        fh.write(hdr + idx + b"".join(b for _, b in recs))


def motifs():
    contigs = [("chrM1", 1_200_345, 14_000), ("chrM2", 90_000, 6_000)]
    arrays, cols, seqs = {}, {}, []
    for idx, (name, ln, n) in enumerate(contigs):
        st, sp, mq, fw = synth_fragments(ln, n, idx, seed_base=9000)
        if name == "chrM1":
            # fragments straddling the 1 Mb window edge (counted in both windows)
            st[:6] = np.array([999_900, 999_950, 999_990, 999_999, 1_000_000, 999_830]); sp[:6] = st[:6] + np.array([167, 100, 20, 2, 150, 170])
            # fragments near contig start (forward k-mer in-bounds, len < k at reverse)
            st[6:9] = np.array([0, 1, 2]); sp[6:9] = np.array([4, 5, 200])
            # forward k-mer out of bounds at the contig end -> fragment skipped entirely
            # (reference frag/_end_motifs.py:135-136), even when stop > contig length
            st[9:12] = np.array([ln - 2, ln - 1, ln - 2]); sp[9:12] = np.array([ln, ln + 5, ln])
            # breakpoint motifs: forward window in bounds, reverse window [stop-3, stop+3) past the contig end
            st[12:15] = np.array([ln - 120, ln - 90, 3]); sp[12:15] = np.array([ln - 1, ln - 2, 150])
            order = np.argsort(st, kind="stable")
            st, sp, mq, fw = st[order], sp[order], mq[order], fw[order]
        codes, nm = synth_twobit(ln, idx, seed_base=9100, telomere=2_000, n_blocks=3, block_len=5_000)
        cols[name] = (st, sp, mq, fw)
        seqs.append((name, codes, nm))
        for k, v in zip(("start", "stop", "mapq", "strand"), (st, sp, mq, fw)):
            arrays[f"{name}_{k}"] = v
        arrays[f"{name}_codes_packed"] = np.packbits(np.unpackbits(codes[:, None], axis=1)[:, 6:].reshape(-1))
        arrays[f"{name}_nmask_packed"] = np.packbits(nm)
    tb_path = os.path.join(TMP, "synth.2bit")
    write_2bit(tb_path, seqs)
    tbx = pysam.TabixFile.from_columns(cols)
    m = {"contigs": [[c, l] for c, l, _ in contigs]}
    reg = []
    j = 0
    for (c, s, e, kw) in [("chrM1", 0, 1_000_000, dict()), ("chrM1", 1_000_000, 1_200_345, dict()),
                          ("chrM1", 400_000, 400_500, dict(quality_threshold=0)),
                          ("chrM2", 0, 90_000, dict(both_strands=False)), ("chrM2", 0, 90_000, dict(both_strands=False, negative_strand=True)),
                          ("chrM2", 10_000, 30_000, dict(k=3)), ("chrM2", 10_000, 30_000, dict(k=6, quality_threshold=30)),
                          ("chrM2", 500, 2500, dict(k=1))]:
        d = F.region_end_motifs(tbx, c, s, e, tb_path, **kw)
        arrays[f"region_{j}"] = np.array(list(d.values()), np.int64)
        reg.append(dict(contig=c, start=s, stop=e, kwargs=kw, key=f"region_{j}"))
        j += 1
    m["region_end_motifs"] = reg
    # reverse k-mer out of bounds (stop < k) raises RuntimeError (frag/_end_motifs.py:144-151)
    bad = pysam.TabixFile.from_columns({"chrM2": (np.array([0, 50], np.int32), np.array([3, 220], np.int32), np.array([60, 60], np.uint8), np.array([1, 0], np.uint8))})
    try:
        F.region_end_motifs(bad, "chrM2", 0, 1000, tb_path)
        m["reverse_oob_error"] = None
    except Exception as e:  # noqa: BLE001
        m["reverse_oob_error"] = type(e).__name__
    gw = []
    for j, kw in enumerate([dict(), dict(k=3, quality_threshold=0), dict(both_strands=False), dict(both_strands=False, negative_strand=True, k=5)]):
        out = os.path.join(TMP, f"em_{j}.tsv")
        r = F.end_motifs(tbx, tb_path, output_file=out, workers=1, **kw)
        arrays[f"genome_freq_{j}"] = np.array(r.frequencies(), np.float64)
        gw.append(dict(kwargs=kw, key=f"genome_freq_{j}", mds=r.motif_diversity_score(), tsv=open(out).read()))
    m["end_motifs"] = gw
    ivs = [("chrM1", 0, 50_000, "a"), ("chrM1", 990_000, 1_010_000, "edge"), ("chrM2", 100, 90_000, "."), ("chrM1", 3, 4, "tiny"), ("chrM2", 1000, 1000, "empty")]
    iv = []
    for j, kw in enumerate([dict(), dict(k=2, quality_threshold=0, both_strands=False)]):
        out = os.path.join(TMP, f"iem_{j}.tsv")
        r = F.interval_end_motifs(tbx, tb_path, [tuple(x) for x in ivs], output_file=out, workers=1, **kw)
        arrays[f"interval_counts_{j}"] = np.array([list(d.values()) for _, d in r.intervals], np.int64)
        buf_counts = os.path.join(TMP, f"iem_{j}_counts.tsv")
        r.to_tsv(buf_counts, calc_freq=False)
        mdsb = os.path.join(TMP, f"iem_{j}_mds.bed")
        r.mds_bed(mdsb)
        iv.append(dict(kwargs=kw, key=f"interval_counts_{j}", tsv=open(out).read(), tsv_counts=open(buf_counts).read(),
                       mds=[v for _, v in r.motif_diversity_score()], mds_mm=[v for _, v in r.motif_diversity_score(miller_madow=True)],
                       mds_bed=open(mdsb).read()))
    m["interval_end_motifs"] = iv
    m["intervals"] = [list(x) for x in ivs]
    # breakpoint motifs (frag/_breakpoint_motifs.py) on the same fragments / reference
    reg = []
    for j, (c, s, e, kw) in enumerate([
            ("chrM1", 0, 1_000_000, dict()), ("chrM1", 1_000_000, 1_200_345, dict()),
            ("chrM1", 1_190_000, 1_200_345, dict(k=4, quality_threshold=0)),
            ("chrM1", 0, 3_000, dict(k=4, quality_threshold=0)),
            ("chrM2", 0, 90_000, dict(both_strands=False)),
            ("chrM2", 0, 90_000, dict(both_strands=False, negative_strand=True)),
            ("chrM2", 10_000, 30_000, dict(k=3)), ("chrM2", 10_000, 30_000, dict(k=2, quality_threshold=30)),
            ("chrM2", 500, 2500, dict(k=8))]):
        d = F.region_breakpoint_motifs(tbx, c, s, e, tb_path, **kw)
        arrays[f"bp_region_{j}"] = np.array(list(d.values()), np.int64)
        reg.append(dict(contig=c, start=s, stop=e, kwargs=kw, key=f"bp_region_{j}"))
    m["region_breakpoint_motifs"] = reg
    gw = []
    for j, kw in enumerate([dict(), dict(k=4, quality_threshold=0), dict(both_strands=False),
                            dict(both_strands=False, negative_strand=True, k=2)]):
        out = os.path.join(TMP, f"bm_{j}.tsv")
        r = F.breakpoint_motifs(tbx, tb_path, output_file=out, workers=1, **kw)
        arrays[f"bp_genome_freq_{j}"] = np.array(r.frequencies(), np.float64)
        gw.append(dict(kwargs=kw, key=f"bp_genome_freq_{j}", mds=r.motif_diversity_score(), tsv=open(out).read()))
    m["breakpoint_motifs"] = gw
    iv = []
    for j, kw in enumerate([dict(k=4), dict(k=2, quality_threshold=0, both_strands=False)]):
        out = os.path.join(TMP, f"ibm_{j}.tsv")
        r = F.interval_breakpoint_motifs(tbx, tb_path, [tuple(x) for x in ivs], output_file=out, workers=1, **kw)
        arrays[f"bp_interval_counts_{j}"] = np.array([list(d.values()) for _, d in r.intervals], np.int64)
        iv.append(dict(kwargs=kw, key=f"bp_interval_counts_{j}", tsv=open(out).read(),
                       mds=[v for _, v in r.motif_diversity_score()]))
    m["interval_breakpoint_motifs"] = iv
    np.savez_compressed(os.path.join(OUT, "motif.npz"), **arrays)
    manifest["motif"] = m


# ----------------------------------------------------------------------------
# D. adjust_wps numeric core on synthetic WPS (numpy + the installed scipy,
#    exactly the calls at reference frag/_adjust_wps.py:131-138)
# ----------------------------------------------------------------------------
def adjust(tbx, cols, cs_path):
    arrays = {}
    m = {}
    # raw WPS bigWig over tiled 5 kb intervals of chrB via the reference's multi_wps
    bed = os.path.join(TMP, "tile.bed")
    with open(bed, "w") as fh:
        for mid in range(2500, 80_000, 5000):
            fh.write(f"chrB\t{mid}\t{mid}\n")
    bw = os.path.join(TMP, "tile.bw")
    F.multi_wps(tbx, bed, chrom_sizes=cs_path, output_file=bw, workers=1)
    ent = pyBigWig._STORE[bw]["data"]["chrB"]
    raw_pos = np.concatenate([e[0] for e in ent]); raw_val = np.concatenate([e[2] for e in ent])
    arrays["raw_pos"] = raw_pos; arrays["raw_val_f32"] = raw_val.astype(np.float32)
    m["tile_bed"] = open(bed).read()
    cases = []
    # BED for adjust: same sites (tab-separated) -> merge rule kicks in for adjacent 5 kb tiles
    for j, (bedtxt, kw) in enumerate([
        (open(bed).read(), dict()),
        ("chrB\t10000\t10000\nchrB\t30000\t30000\nchrB\t31000\t31000\n", dict()),
        ("chrB\t10000\t10000\nchrB\t30000\t30000\n", dict(mean=True)),
        ("chrB\t10000\t10000\nchrB\t30000\t30000\n", dict(subtract_edges=True)),
        ("chrB\t10000\t10000\n", dict(savgol=False, median_window_size=300)),
        ("chrB\t10000\t10000\n", dict(savgol_window_size=31, savgol_poly_deg=4, median_window_size=100, interval_size=2000)),
        ("chrB\t1000\t1000\nchrB\t79000\t79000\n", dict()),
    ]):
        b = os.path.join(TMP, f"a_{j}.bed")
        open(b, "w").write(bedtxt)
        obw = os.path.join(TMP, f"a_{j}.bw")
        try:
            F.adjust_wps(bw, b, obw, cs_path, workers=1, **kw)
            err = None
        except Exception as e:  # noqa: BLE001
            err = type(e).__name__
        d = pyBigWig._STORE.get(obw, {"data": {}})["data"].get("chrB", [])
        pos = np.concatenate([e[0] for e in d]) if d else np.zeros(0, np.int64)
        val = np.concatenate([e[2] for e in d]) if d else np.zeros(0, np.float32)
        arrays[f"adj_{j}_pos"] = pos; arrays[f"adj_{j}_val_f32"] = val.astype(np.float32)
        cases.append(dict(bed=bedtxt, kwargs=kw, n=int(pos.shape[0]), error=err))
    m["adjust_cases"] = cases
    # fp64 numeric core on raw integer vectors (no float32 output rounding)
    core = []
    rng = np.random.Generator(np.random.PCG64(99))
    seg = raw_val[:5000].astype(np.float64)
    vectors = {"wps5000": seg, "rand_small": rng.integers(-30, 30, 1500).astype(np.float64),
               "const": np.full(1200, 3.0), "ramp": np.arange(2500, dtype=np.float64) % 17 - 8}
    for name, x in vectors.items():
        arrays[f"core_in_{name}"] = x
        for w, use_mean, sg in [(1000, False, (21, 2)), (1000, True, (21, 2)), (100, False, (11, 3)), (64, False, None), (2, False, (5, 2))]:
            if w > len(x):
                continue
            pos, adj = _local_filter(np.arange(len(x), dtype=np.int64), x, w, use_mean)
            y = savgol_filter(adj, sg[0], sg[1]) if sg else adj
            key = f"core_{name}_{w}_{int(use_mean)}_{sg[0] if sg else 0}_{sg[1] if sg else 0}"
            arrays[key + "_pre"] = adj; arrays[key + "_out"] = y
            core.append(dict(input=f"core_in_{name}", w=w, mean=use_mean, sg=list(sg) if sg else None, key=key, first_pos=int(pos[0]) if len(pos) else -1))
    m["core_cases"] = core
    np.savez_compressed(os.path.join(OUT, "adjust.npz"), **arrays)
    manifest["adjust"] = m


# ----------------------------------------------------------------------------
# E. cleavage profile
# ----------------------------------------------------------------------------
def cleavage(tbx, cols, cs_path):
    arrays, m = {}, {}
    frag = os.path.join(DATA, "12.3444.b37.frag.gz")
    cs = os.path.join(DATA, "b37.chrom.sizes")
    cases = []
    for j, (src, contig, size, s, e, kw) in enumerate([
        ("fixture", "12", 133851895, 34443000, 34446700, dict()),
        ("fixture", "12", 133851895, 34443118, 34443119, dict(left=40, right=60, quality_threshold=0)),
        ("fixture", "12", 133851895, 34444900, 34445200, dict(min_length=150, max_length=180)),
        ("synth", "chrA", 200_000, 0, 7000, dict()),
        ("synth", "chrA", 200_000, 193_000, 200_000, dict(right=500, quality_threshold=0)),
        ("synth", "chrA", 200_000, 50_000, 62_000, dict(min_length=120, max_length=180, left=10, right=10)),
        ("synth", "chrC", 4_000, 100, 3_900, dict(left=500, right=500, quality_threshold=60)),
        ("synth", "chrB", 80_000, 40_000, 40_000, dict()),
    ]):
        r = F.cleavage_profile(frag if src == "fixture" else tbx, size, contig, s, e, **kw)
        arrays[f"clv_{j}_pos"] = r["pos"].astype(np.int64)
        arrays[f"clv_{j}_prop"] = r["proportion"].astype(np.float64)
        cases.append(dict(src=src, contig=contig, chrom_size=size, start=s, stop=e, kwargs=kw, n=int(r.shape[0])))
    m["cases"] = cases
    # multi_cleavage_profile: BED with overlaps / unknown contig / padding -> bedgraph.gz + bigWig
    bed_txt = "chrA\t1000\t1400\nchrA\t1300\t2500\nchrA\t9000\t9001\nchrUn\t5\t9\nchrB\t79000\t80000\nchrB\t100\t700\nchrC\t0\t4000\n"
    bed = os.path.join(TMP, "clv.bed")
    open(bed, "w").write(bed_txt)
    m["bed"] = bed_txt
    multi = []
    for j, kw in enumerate([dict(), dict(left=30, right=45, min_length=100, max_length=220, quality_threshold=20)]):
        out = os.path.join(TMP, f"clv_{j}.bed.gz")
        F.multi_cleavage_profile(tbx, bed, cs_path, output_file=out, workers=1, **kw)
        txt = read_gz_text(out)
        obw = os.path.join(TMP, f"clv_{j}.bw")
        F.multi_cleavage_profile(tbx, bed, cs_path, output_file=obw, workers=1, **kw)
        d = pyBigWig._STORE[obw]["data"]
        arrays[f"multi_{j}_bw_val_f32"] = np.concatenate([np.concatenate([e[2] for e in d[c]]) for c in d]).astype(np.float32)
        multi.append(dict(kwargs=kw, n_lines=len(txt.splitlines()), sha256_text=hashlib.sha256(txt.encode()).hexdigest(),
                          head=txt.splitlines()[:3], bw_contigs=list(d)))
    m["multi"] = multi
    np.savez_compressed(os.path.join(OUT, "cleavage.npz"), **arrays)
    manifest["cleavage"] = m


# ----------------------------------------------------------------------------
# F. DELFI bin counts + GC content (frag/_delfi.py): _delfi_single_window per bin and the delfi()
#    table without LOESS (the `loess` package is not installable here)
# ----------------------------------------------------------------------------
def delfi_golden():
    import finaletoolkit.frag._delfi as DL
    from finaletoolkit.genome.gaps import GenomeGaps
    arrays, m = {}, {}
    contigs = [("chr7", 1_300_000, 60_000), ("chr21", 120_000, 8_000), ("chrX9", 64_000, 3_000)]
    cols, seqs = {}, []
    for idx, (name, ln, n) in enumerate(contigs):
        st, sp, mq, fw = synth_fragments(ln, n, idx, seed_base=12_000)
        codes, nm = synth_twobit(ln, idx, seed_base=12_100, telomere=1_500, n_blocks=2, block_len=3_000)
        if n:
            cols[name] = (st, sp, mq, fw)
        seqs.append((name, codes, nm))
        for k, v in zip(("start", "stop", "mapq", "strand"), (st, sp, mq, fw)):
            arrays[f"{name}_{k}"] = v
        arrays[f"{name}_codes_packed"] = np.packbits(np.unpackbits(codes[:, None], axis=1)[:, 6:].reshape(-1))
        arrays[f"{name}_nmask_packed"] = np.packbits(nm)
    tb_path = os.path.join(TMP, "delfi.2bit")
    write_2bit(tb_path, seqs)
    tbx = pysam.TabixFile.from_columns(cols)
    cs_path = os.path.join(TMP, "delfi.chrom.sizes")
    # chrom.sizes order != bins order; chrNoBins has no bins; chrX9 has no gap annotation
    cs_txt = "chr21\t120000\nchr7\t1300000\nchrX9\t64000\nchrEmpty\t30000\nchrNoBins\t5000\n"
    open(cs_path, "w").write(cs_txt)
    gap_txt = ("chr7\t0\t10000\ttelomere\nchr7\t1290000\t1300000\ttelomere\nchr7\t600000\t640000\tcentromere\n"
               "chr21\t0\t2000\ttelomere\nchr21\t50000\t56000\tcentromere\nchr21\t2000\t40000\tshort_arm\n"
               "chrEmpty\t10000\t12000\tcentromere\n")   # contig without bins
    gap_path = os.path.join(TMP, "delfi.gaps.bed")
    open(gap_path, "w").write(gap_txt)
    # blacklist: contained / straddling a bin edge / overlapping pair / tiny / on a contig without gaps
    bl_txt = ("chr7\t101000\t103500\nchr7\t104900\t105300\nchr7\t102000\t102900\nchr7\t700100\t700400\tname\n"
              "chr7\t20000\t25000\nchr7\t900000\t900150\n\nbadline\nchr21\t60000\t64000\nchrX9\t1000\t3000\n"
              "chrX9\t2000\t2600\nchr7\t45000\t50000\n")
    bl_path = os.path.join(TMP, "delfi.blacklist.bed")
    open(bl_path, "w").write(bl_txt)
    bins = []
    for a in range(0, 1_300_000, 5_000):
        bins.append(("chr7", a, a + 5_000))
    for a in range(0, 120_000, 2_000):
        bins.append(("chr21", a, a + 2_000))
    bins.append(("chr21", 118_500, 121_000))          # past the contig end: GC skipped (invalid interval)
    for a in range(0, 64_000, 4_000):
        bins.append(("chrX9", a, a + 4_000))
    bins.append(("chrX9", 500, 3500))                 # overlapping bin, blacklist region contained
    bins_txt = "#comment line\n" + "".join(f"{c}\t{a}\t{b}\n" for c, a, b in bins)
    bins_path = os.path.join(TMP, "delfi.bins.bed")
    open(bins_path, "w").write(bins_txt)
    m.update(contigs=[[c, l] for c, l, _ in contigs], chrom_sizes=cs_txt, gaps=gap_txt, blacklist=bl_txt, bins=bins_txt)

    # per-bin tuples straight from _delfi_single_window (all bins, no gap-overlap prefilter)
    gaps = GenomeGaps(gap_path)
    singles = []
    for q in (30, 0):
        DL._delfi_pool_initializer(tbx, tb_path, q, DL._load_blacklist_indexed(bl_path),
                                   {c: gaps.get_contig_gaps(c) for c, _, _ in contigs})
        rows = [DL._delfi_single_window(c, a, b) for c, a, b in bins]
        key = f"single_q{q}"
        arrays[key + "_short"] = np.array([r[4] for r in rows], np.float64)
        arrays[key + "_long"] = np.array([r[5] for r in rows], np.float64)
        arrays[key + "_gc"] = np.array([r[6] for r in rows], np.float64)
        arrays[key + "_num"] = np.array([r[7] for r in rows], np.int64)
        singles.append(dict(quality_threshold=q, key=key, arms=[r[3] for r in rows], use_gaps=True, use_blacklist=True))
    DL._delfi_pool_initializer(tbx, tb_path, 30, {}, None)
    rows = [DL._delfi_single_window(c, a, b) for c, a, b in bins]
    key = "single_plain"
    arrays[key + "_short"] = np.array([r[4] for r in rows], np.float64)
    arrays[key + "_long"] = np.array([r[5] for r in rows], np.float64)
    arrays[key + "_gc"] = np.array([r[6] for r in rows], np.float64)
    arrays[key + "_num"] = np.array([r[7] for r in rows], np.int64)
    singles.append(dict(quality_threshold=30, key=key, arms=[r[3] for r in rows], use_gaps=False, use_blacklist=False))
    m["single_window"] = singles
    m["bins_list"] = [list(b) for b in bins]

    # the delfi() table (no LOESS)
    tables = []
    for j, kw in enumerate([dict(gap_file=gap_path, blacklist_file=bl_path, merge_bins=False, remove_nocov=False),
                            dict(gap_file=gap_path, blacklist_file=bl_path, merge_bins=True),
                            dict(merge_bins=False, quality_threshold=0),
                            dict(gap_file=gap_path, merge_bins=True, remove_nocov=False, quality_threshold=0)]):
        out = os.path.join(TMP, f"delfi_{j}.tsv")
        df = DL.delfi(tbx, cs_path, bins_path, tb_path, output_file=out, no_gc_correct=True, workers=1, **kw)
        outc = os.path.join(TMP, f"delfi_{j}.csv")
        DL._write_delfi(df, outc)
        tables.append(dict(kwargs={k: (os.path.basename(v) if isinstance(v, str) else v) for k, v in kw.items()},
                           tsv=open(out).read(), csv=open(outc).read(), columns=list(df.columns),
                           dtypes=[str(t) for t in df.dtypes], n_rows=int(df.shape[0])))
    m["delfi"] = tables
    np.savez_compressed(os.path.join(OUT, "delfi.npz"), **arrays)
    manifest["delfi"] = m


# ----------------------------------------------------------------------------
# G. agg_bw (utils/_agg_bw.py): strand-aware aggregate of a bigWig over BED6 intervals -> WIG
# ----------------------------------------------------------------------------
def agg():
    from finaletoolkit.utils._agg_bw import agg_bw
    arrays, m = {}, {}
    rng = np.random.default_rng(77)
    sizes = [("chrA", 60_000), ("chrB", 9_000)]
    # non-integer float32 signal (adjusted-WPS-like): the fp64 running sum is order dependent
    sig = {"chrA": (rng.normal(0, 7, 50_000)).astype(np.float32), "chrB": rng.integers(-30, 9, 8_000).astype(np.float32)}
    starts = {"chrA": 2_000, "chrB": 500}
    path = os.path.join(TMP, "agg.bw")
    bw = pyBigWig.open(path, "w")
    bw.addHeader(sizes)
    # chrA written in two runs with a 300-base hole (values() -> nan there)
    bw.addEntries("chrA", 2_000, values=sig["chrA"][:20_000].astype(np.float64), span=1, step=1)
    bw.addEntries("chrA", 22_300, values=sig["chrA"][20_300:].astype(np.float64), span=1, step=1)
    bw.addEntries("chrB", 500, values=sig["chrB"].astype(np.float64), span=1, step=1)
    bw.close()
    arrays["chrA_signal"], arrays["chrB_signal"] = sig["chrA"], sig["chrB"]
    m["sizes"] = [list(x) for x in sizes]
    m["layout"] = [["chrA", 2_000, 0, 20_000], ["chrA", 22_300, 20_300, 50_000], ["chrB", 500, 0, 8_000]]  # contig, pos, lo, hi
    rows = []
    for i in range(40):
        c = "chrA" if i % 4 else "chrB"
        lim = 50_000 if c == "chrA" else 7_000
        a = int(rng.integers(starts[c], starts[c] + lim - 2_000))
        rows.append((c, a, a + 2_000, ".", "0", "+-."[int(rng.integers(0, 3)) if i % 7 else 2]))
    rows[3] = ("chrA", 21_000, 23_000, ".", "0", "-")        # spans the hole
    rows[5] = ("chrA", 59_000, 61_000, ".", "0", "+")        # past the contig end -> RuntimeError, skipped
    rows[6] = ("chrA", 30_000, 31_999, ".", "0", "+")        # wrong size -> skipped
    rows[8] = ("chrZ", 100, 2_100, ".", "0", "+")            # unknown contig -> skipped
    rows[9] = ("chrA", 100, 2_100, ".", "0", "-")            # mostly uncovered
    bed_txt = "".join("\t".join(str(x) for x in r) + "\n" for r in rows)
    bed = os.path.join(TMP, "agg.bed")
    open(bed, "w").write(bed_txt)
    m["bed"] = bed_txt
    cases = []
    for j, kw in enumerate([dict(), dict(median_window_size=120), dict(median_window_size=121, mean=True),
                            dict(median_window_size=0), dict(median_window_size=1000, mean=True)]):
        out = os.path.join(TMP, f"agg_{j}.wig")
        with redirect_stdout(io.StringIO()):
            r = agg_bw(path, bed, out, **kw)
        arrays[f"agg_{j}"] = np.asarray(r)
        cases.append(dict(kwargs=kw, key=f"agg_{j}", dtype=str(np.asarray(r).dtype), wig=open(out).read()))
    m["cases"] = cases
    # the reference's own fixture (tests/data/test.bw, bw_test.bed) and known answers (tests/test_agg_bw.py:15-26)
    arrays["ref_test_bw"] = np.frombuffer(open(os.path.join(DATA, "test.bw"), "rb").read(), np.uint8)
    m["ref_bed"] = open(os.path.join(DATA, "bw_test.bed")).read()
    m["ref_known"] = [dict(median_window_size=0, expect=[0., 0., 0., 0., 0.]), dict(median_window_size=2, expect=[1., 2., 3.])]
    np.savez_compressed(os.path.join(OUT, "agg.npz"), **arrays)
    manifest["agg"] = m


if __name__ == "__main__":
    fixture17()
    tbx, cols, cs_path = synth_small()
    motifs()
    adjust(tbx, cols, cs_path)
    cleavage(tbx, cols, cs_path)
    delfi_golden()
    agg()
    with open(os.path.join(OUT, "manifest.json"), "w") as fh:
        json.dump(manifest, fh, indent=1, default=lambda o: o.item() if hasattr(o, "item") else str(o))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
