#!/usr/bin/env python
"""Golden vectors for the READ-level region selection of BAM input (TEST INFRASTRUCTURE).

    python oracle/make_golden_bam.py        # build container only (/root/reference)

For BAM / CRAM the reference does not select fragments by their own span: ``AlignmentWrapper._fetch_sam``
(io/alignment.py:242-268) iterates ``pysam.AlignmentFile.fetch(contig, start, stop)``, i.e. the READS
overlapping the region, and only then rebuilds the fragment from ``template_length``.  A fragment that
reaches into the region with its mate's end only is therefore invisible to every per-region feature.
This script writes a synthetic BAM with short reads (30-70 reference bases for 80-260 bp templates, so
nearly every region edge cuts some fragment between read 1 and mate), runs the UNMODIFIED reference on it
under the in-memory stand-ins of ``oracle/fakes`` (whose ``AlignmentFile.fetch`` applies htslib's rule:
``pos < stop and bam_endpos > start``) and stores every output next to the BAM bytes:

    tests/golden/bam_read1.npz, tests/golden/bam_read1.json

``tests/test_oracle_golden.py`` pins the oracle's restatement on them (CPU), ``tests/test_gpu_bam.py``
replays them through the CUDA path.
"""
from __future__ import annotations

import gzip
import hashlib
import io
import json
import os
import sys
import tempfile
import warnings
from contextlib import redirect_stderr, redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = "/root/reference"
sys.path[:0] = [os.path.join(HERE, "fakes"), os.path.join(REF, "src"), REPO, os.path.join(REPO, "tests")]

import pysam  # noqa: E402,F401  (the fake)
import finaletoolkit.frag as F  # noqa: E402  (the reference)
import finaletoolkit.frag._delfi as DL  # noqa: E402
import finaletoolkit.utils as U  # noqa: E402

from finaletoolkit_b200.synth import synth_twobit  # noqa: E402
from helpers import write_bam  # noqa: E402
from make_golden import write_2bit  # noqa: E402  (same directory; its __main__ block does not run on import)

warnings.simplefilter("ignore")
OUT = os.path.join(REPO, "tests", "golden")
TMP = tempfile.mkdtemp(prefix="ftk_golden_bam_")

REFS = [("chrA", 60_000), ("chrB", 25_000)]
GOOD = 0x1 | 0x2 | 0x40            # paired, proper pair, read 1


def synth_records(seed=20261017):
    """Coordinate-sorted BAM records: templates of 80-260 bp, read 1 of 30-70 reference bases on either end
    (forward: tlen > 0 at the template start; reverse: tlen < 0 at the template end), CIGARs with soft
    clips / insertions / deletions, a sprinkle of reads the flag filter drops, and - on chrB only - a few
    dovetailed pairs whose read 1 is longer than the template."""
    rng = np.random.default_rng(seed)
    recs = []
    for ref, (name, length), n in ((0, REFS[0], 5200), (1, REFS[1], 1500)):
        for _ in range(n):
            tl = int(rng.integers(80, 261))
            rlen = int(rng.integers(30, 71))
            if ref == 1 and rng.random() < 0.05:
                rlen = tl + int(rng.integers(1, 25))                      # dovetail: read 1 runs past its template
            f_start = int(rng.integers(20, length - 300))
            forward_anchor = rng.random() < 0.5
            ops = []
            if rng.random() < 0.25:
                ops.append((4, int(rng.integers(1, 9))))                   # soft clip: no reference bases
            u = rng.random()
            if u < 0.2 and rlen > 20:                                     # deletion inside the read
                d = int(rng.integers(1, 8)); a = int(rng.integers(5, rlen - d - 5))
                ops += [(0, a), (2, d), (0, rlen - a - d)]
            elif u < 0.35 and rlen > 20:                                  # insertion: query bases only
                a = int(rng.integers(5, rlen - 5))
                ops += [(0, a), (1, int(rng.integers(1, 6))), (0, rlen - a)]
            else:
                ops.append((0, rlen))
            if rng.random() < 0.15:
                ops.append((5, 4))                                        # hard clip
            if forward_anchor:
                pos, tlen = f_start, tl
            else:
                pos, tlen = f_start + tl - rlen, -tl                      # reference_end = template end
            if pos < 0:
                continue
            flag = GOOD | (0 if (forward_anchor ^ (rng.random() < 0.04)) else 0x10)
            v = rng.random()
            if v < 0.04:
                flag ^= int(rng.choice([0x2, 0x4, 0x8, 0x100, 0x200, 0x400, 0x800]))
            elif v < 0.08:
                flag = (flag & ~0x40) | 0x80                              # read 2
            mapq = int(rng.choice([0, 5, 19, 20, 29, 30, 31, 42, 60], p=[.04, .03, .03, .05, .05, .1, .1, .2, .4]))
            recs.append(dict(ref=ref, pos=pos, mapq=mapq, flag=flag, tlen=tlen, cigar=ops))
    # hand-placed templates at the edges of the WPS cases below: long template, 30-base read 1 wholly outside the
    # padded fetch window [start - max_length, stop + max_length) while the template still reaches the first /
    # last scoring windows - the only constellation in which the read-level fetch changes a WPS value
    for ref, edge_lo, edge_hi, max_len in ((0, 10_000, 12_000, 180), (1, 12_000, 14_000, 240)):
        for d in (0, 3, 7):
            recs.append(dict(ref=ref, pos=edge_lo - max_len - 45 + d, mapq=60, flag=GOOD, tlen=max_len - d,
                             cigar=[(0, 30)]))                                     # forward read 1 left of the window
            end = edge_hi + max_len + 45 - d                                       # reverse read 1 right of it
            recs.append(dict(ref=ref, pos=end - 30, mapq=60, flag=GOOD | 0x10, tlen=-(max_len - d), cigar=[(0, 30)]))
    recs.sort(key=lambda r: (r["ref"], r["pos"]))
    return recs


def dovetail_regions(records, n=4):
    """Small regions on chrB that only the OVERHANG of a dovetailed read 1 touches (read past its template)."""
    out = []
    for r in records:
        ref_len = sum(ln for op, ln in r["cigar"] if op in (0, 2, 3, 7, 8))
        if r["ref"] != 1 or r["flag"] != GOOD and r["flag"] != (GOOD | 0x10) or ref_len <= abs(r["tlen"]):
            continue
        if r["tlen"] > 0:
            out.append(("chrB", r["pos"] + r["tlen"], r["pos"] + r["tlen"] + 3))        # just right of the template
        else:
            end = r["pos"] + ref_len
            out.append(("chrB", max(end + r["tlen"] - 3, 0), end + r["tlen"]))           # just left of it
        if len(out) == n:
            break
    return out


def main():
    arrays, m = {}, {"generator": "oracle/make_golden_bam.py", "refs": [list(r) for r in REFS]}
    records = synth_records()
    bam = write_bam(os.path.join(TMP, "read1.bam"), REFS, records, block=6000)
    arrays["bam_file"] = np.frombuffer(open(bam, "rb").read(), np.uint8)
    seqs = []
    for idx, (name, ln) in enumerate(REFS):
        codes, nm = synth_twobit(ln, idx, seed_base=31_000, telomere=0, n_blocks=1, block_len=300)
        seqs.append((name, codes, nm))
        arrays[f"{name}_codes_packed"] = np.packbits(np.unpackbits(codes[:, None], axis=1)[:, 6:].reshape(-1))
        arrays[f"{name}_nmask_packed"] = np.packbits(nm)
    tb = os.path.join(TMP, "read1.2bit")
    write_2bit(tb, seqs)
    cs = os.path.join(TMP, "cs")
    m["chrom_sizes"] = "".join(f"{c}\t{n}\n" for c, n in REFS)
    open(cs, "w").write(m["chrom_sizes"])

    # ---- the stream itself (utils/_frag_generator.py:58-141 over io/alignment.py:242-268)
    gen = []
    for j, (c, s, e, kw) in enumerate([
            ("chrA", 10_000, 12_000, dict()), ("chrA", 10_000, 12_000, dict(intersect_policy="any")),
            ("chrA", 0, 700, dict(quality_threshold=0)), ("chrA", 59_000, 60_000, dict(intersect_policy="any", quality_threshold=20)),
            ("chrA", 30_000, 30_050, dict(min_length=120, max_length=180)), ("chrA", None, None, dict()),
            ("chrA", 40_000, None, dict(intersect_policy="any")), ("chrA", None, 3_000, dict()),
            ("chrB", 5_000, 9_000, dict(intersect_policy="any")), ("chrB", 5_000, 9_000, dict()),
            (None, None, None, dict(quality_threshold=42))]):
        rows = list(U.frag_generator(bam, c, start=s, stop=e, **kw))
        arrays[f"gen_{j}"] = np.array([[r[1], r[2], r[3], int(r[4])] for r in rows], np.int64).reshape(-1, 4)
        gen.append(dict(contig=c, start=s, stop=e, kwargs=kw, key=f"gen_{j}", contigs=sorted({r[0] for r in rows}) if c is None else None))
    m["frag_generator"] = gen
    fa = U.frag_array(bam, "chrA", start=20_000, stop=21_000, min_length=100, max_length=200)
    arrays["frag_array_0"] = np.stack([fa["start"], fa["stop"], fa["strand"].astype(np.int64)], axis=1)
    m["frag_array"] = [dict(contig="chrA", start=20_000, stop=21_000, kwargs=dict(min_length=100, max_length=200), key="frag_array_0")]

    # ---- WPS (frag/_wps.py:142-188: padded fetch window, midpoint policy on the window)
    cases = []
    for j, (c, s, e, kw) in enumerate([
            ("chrA", 10_000, 12_000, dict()), ("chrA", 0, 1_500, dict()), ("chrA", 58_800, 60_000, dict()),
            ("chrA", 25_000, 25_400, dict(window_size=60, min_length=80, max_length=260, quality_threshold=0)),
            ("chrB", 12_000, 14_000, dict(window_size=121, min_length=100, max_length=240))]):
        r = F.wps(bam, c, s, e, dict(REFS)[c], **kw)
        arrays[f"wps_{j}"] = r["wps"].astype(np.int64)
        cases.append(dict(contig=c, start=s, stop=e, kwargs=kw, key=f"wps_{j}"))
    m["wps"] = cases
    sites = os.path.join(TMP, "sites.bed")
    site_txt = "".join(f"chrA\t{a}\t{a + 1}\t.\t0\t+\n" for a in range(1_000, 60_000, 2_000)) + \
        "".join(f"chrB\t{a}\t{a + 1}\t.\t0\t-\n" for a in range(1_500, 25_000, 3_000))
    open(sites, "w").write(site_txt)
    out = os.path.join(TMP, "mw.bed.gz")
    with redirect_stderr(io.StringIO()):
        F.multi_wps(bam, sites, chrom_sizes=cs, output_file=out, interval_size=2_000, min_length=100, max_length=220,
                    workers=1)
    text = gzip.open(out, "rt").read()
    arrays["multi_wps_scores"] = np.array([int(ln.split("\t")[3]) for ln in text.splitlines()], np.int64)
    m["multi_wps"] = dict(sites=site_txt, kwargs=dict(interval_size=2_000, min_length=100, max_length=220),
                          sha256=hashlib.sha256(text.encode()).hexdigest(), n_lines=len(text.splitlines()),
                          head=text[:200])

    # ---- coverage (frag/_coverage.py:53-137, 145-305)
    cov = []
    for c, s, e, kw in [("chrA", 10_000, 12_000, dict()), ("chrA", 10_000, 12_000, dict(intersect_policy="any")),
                        ("chrA", 0, None, dict()), ("chrA", 33_333, 33_400, dict(intersect_policy="any", min_length=100, max_length=200)),
                        ("chrB", 8_000, 8_500, dict(quality_threshold=0)), (None, 0, None, dict())]:
        cov.append(dict(contig=c, start=s, stop=e, kwargs=kw, result=list(F.single_coverage(bam, c, s, e, **kw))))
    m["single_coverage"] = cov
    tiles = os.path.join(TMP, "tiles.bed")
    tile_txt = "".join(f"chrA\t{a}\t{a + 1_000}\tt{a}\n" for a in range(0, 60_000, 1_000)) + \
        "".join(f"chrB\t{a}\t{a + 700}\n" for a in range(0, 24_500, 500)) + "chrA\t5\t59990\twide\n"
    open(tiles, "w").write(tile_txt)
    m["tiles"] = tile_txt
    covs = []
    for j, kw in enumerate([dict(), dict(intersect_policy="any", normalize=True, scale_factor=1e6),
                            dict(min_length=120, max_length=180, quality_threshold=20)]):
        out = os.path.join(TMP, f"cov_{j}.bed")
        F.coverage(bam, tiles, out, workers=1, **kw)
        covs.append(dict(kwargs=kw, text=open(out).read()))
    m["coverage"] = covs

    # ---- fragment lengths (frag/_frag_length.py)
    fl = []
    for j, (c, s, e, kw) in enumerate([("chrA", 10_000, 12_000, dict()), ("chrA", 10_000, 12_000, dict(intersect_policy="any")),
                                        ("chrB", None, None, dict(quality_threshold=0)), (None, None, None, dict())]):
        arrays[f"frag_length_{j}"] = np.asarray(F.frag_length(bam, c, s, e, **kw), np.int64)
        fl.append(dict(contig=c, start=s, stop=e, kwargs=kw, key=f"frag_length_{j}"))
    m["frag_length"] = fl
    flb = []
    for c, s, e, kw in [("chrA", 10_000, 14_000, dict(bin_size=5)), ("chrA", 20_000, 20_600, dict(bin_size=10, intersect_policy="any")),
                        ("chrB", 0, 25_000, dict(bin_size=20, min_length=100, max_length=200))]:
        bins, counts = F.frag_length_bins(bam, c, s, e, **kw)
        flb.append(dict(contig=c, start=s, stop=e, kwargs=kw, bins=np.asarray(bins).tolist(), counts=np.asarray(counts).tolist()))
    m["frag_length_bins"] = flb
    fli = []
    for kw in [dict(), dict(intersect_policy="any", short_reads=120, quality_threshold=20)]:
        rows = F.frag_length_intervals(bam, tiles, workers=1, **kw)
        fli.append(dict(kwargs=kw, rows=[[x if isinstance(x, (str, int)) else float(x) for x in r] for r in rows]))
    m["frag_length_intervals"] = fli

    # ---- end / breakpoint motifs (frag/_end_motifs.py:115-120: every fetched fragment counts, dovetails included)
    ivs = [("chrA", a, a + 1_500, ".") for a in range(500, 58_000, 1_500)] + [("chrA", 30_000, 30_040, "tiny"), ("chrA", 100, 59_900, "wide")]
    mot = []
    for j, (fn, kw) in enumerate([("region_end_motifs", dict(k=4)), ("region_end_motifs", dict(k=3, both_strands=False)),
                                  ("region_end_motifs", dict(k=2, both_strands=False, negative_strand=True, quality_threshold=0)),
                                  ("region_breakpoint_motifs", dict(k=4)), ("region_breakpoint_motifs", dict(k=6, both_strands=False))]):
        for r, (c, s, e) in enumerate([("chrA", 10_000, 12_000), ("chrA", 45_100, 45_160), ("chrA", 1_000, 59_000),
                                       ("chrB", 5_000, 5_600), ("chrB", 12_000, 12_100)] + dovetail_regions(records)):
            d = getattr(F, fn)(bam, c, s, e, tb, **kw)
            arrays[f"motif_{j}_{r}"] = np.array(list(d.values()), np.int64)
            mot.append(dict(fn=fn, contig=c, start=s, stop=e, kwargs=kw, key=f"motif_{j}_{r}"))
    m["region_motifs"] = mot
    imot = []
    for j, (fn, kw) in enumerate([("interval_end_motifs", dict(k=3)), ("interval_breakpoint_motifs", dict(k=4, quality_threshold=20))]):
        res = getattr(F, fn)(bam, tb, ivs, workers=1, **kw)
        arrays[f"imotif_{j}"] = np.array([list(d.values()) for _, d in res.intervals], np.int64)
        imot.append(dict(fn=fn, kwargs=kw, key=f"imotif_{j}"))
    # chrB carries dovetailed pairs: the motif counters take a fetched fragment without any fragment-level test
    ivs_b = [("chrB", a, a + 400, ".") for a in range(300, 24_000, 400)] + [tuple(x) + ("dovetail",) for x in dovetail_regions(records)]
    for j, (fn, kw) in enumerate([("interval_end_motifs", dict(k=3)), ("interval_breakpoint_motifs", dict(k=4, quality_threshold=0))]):
        res = getattr(F, fn)(bam, tb, ivs_b, workers=1, **kw)
        arrays[f"imotif_b_{j}"] = np.array([list(d.values()) for _, d in res.intervals], np.int64)
        imot.append(dict(fn=fn, kwargs=kw, key=f"imotif_b_{j}", intervals="motif_intervals_b"))
    m["interval_motifs"] = imot
    m["motif_intervals"] = [list(iv) for iv in ivs]
    m["motif_intervals_b"] = [list(iv) for iv in ivs_b]

    # ---- cleavage profile (frag/_cleavage_profile.py:188-217: fetch of the padded region, policy "any")
    clv = []
    for j, (c, s, e, kw) in enumerate([("chrA", 10_000, 11_000, dict()), ("chrA", 40, 600, dict(left=100, right=50)),
                                        ("chrB", 24_500, 24_990, dict(left=0, right=100, min_length=100, max_length=220, quality_threshold=0))]):
        r = F.cleavage_profile(bam, dict(REFS)[c], c, s, e, **kw)
        arrays[f"cleavage_{j}"] = r["proportion"].astype(np.float64)
        arrays[f"cleavage_{j}_pos"] = r["pos"].astype(np.int64)
        clv.append(dict(contig=c, start=s, stop=e, kwargs=kw, key=f"cleavage_{j}"))
    m["cleavage_profile"] = clv
    out = os.path.join(TMP, "clv.bed.gz")
    clv_bed = os.path.join(TMP, "clv.bed")
    clv_txt = "".join(f"chrA\t{a}\t{a + 400}\n" for a in range(2_000, 58_000, 3_000)) + "chrB\t100\t900\nchrB\t20000\t20500\n"
    open(clv_bed, "w").write(clv_txt)
    with redirect_stderr(io.StringIO()), redirect_stdout(io.StringIO()):
        F.multi_cleavage_profile(bam, clv_bed, cs, left=20, right=20, output_file=out, workers=1)
    text = gzip.open(out, "rt").read()
    m["multi_cleavage_profile"] = dict(bed=clv_txt, kwargs=dict(left=20, right=20),
                                       sha256=hashlib.sha256(text.encode()).hexdigest(), n_lines=len(text.splitlines()),
                                       head=text[:200])

    # ---- DELFI bins (frag/_delfi.py:404-511: fetch of the bin, 100-220 bp, midpoint in the bin)
    bins = [("chrA", a, a + 2_000) for a in range(0, 60_000, 2_000)] + [("chrB", a, a + 2_500) for a in range(0, 25_000, 2_500)]
    dl = []
    for q in (30, 0):
        DL._delfi_pool_initializer(bam, tb, q, {}, None)
        rows = [DL._delfi_single_window(c, a, b) for c, a, b in bins]
        key = f"delfi_q{q}"
        arrays[key] = np.array([[r[4], r[5], r[7]] for r in rows], np.int64)
        arrays[key + "_gc"] = np.array([r[6] for r in rows], np.float64)
        dl.append(dict(quality_threshold=q, key=key))
    m["delfi"] = dict(bins=[list(b) for b in bins], cases=dl)

    np.savez_compressed(os.path.join(OUT, "bam_read1.npz"), **arrays)
    with open(os.path.join(OUT, "bam_read1.json"), "w") as fh:
        json.dump(m, fh, indent=1, default=lambda o: o.item() if hasattr(o, "item") else str(o))
    for f in ("bam_read1.npz", "bam_read1.json"):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
