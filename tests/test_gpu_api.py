"""The public API (finaletoolkit_b200.frag / flat namespace) replayed against outputs of the
unmodified reference (tests/golden/manifest.json): same calls, same kwargs, same files out."""
import hashlib
import warnings

import numpy as np
import pytest

from helpers import golden_codes, read_gz, write_2bit, write_frag_gz, write_text_gz

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fx(tmp_path_factory, manifest):
    """The reference's 17-fragment fixture as files."""
    d = tmp_path_factory.mktemp("fx17")
    m = manifest["fixture17"]
    p = {"frag": write_text_gz(d / "12.3444.b37.frag.gz", m["frag_gz_text"]),
         "bed6": write_text_gz(d / "12.3444.b37.frag.bed.gz", m["frag_bed_gz_text"]),
         "ivl": str(d / "intervals.bed"), "ivl_ov": str(d / "intervals_overlapped.bed"), "cs": str(d / "b37.chrom.sizes"),
         "dir": d}
    open(p["ivl"], "w").write(m["intervals_bed"])
    open(p["ivl_ov"], "w").write(m["intervals_overlapped_bed"])
    open(p["cs"], "w").write(m["chrom_sizes"])
    return p


@pytest.fixture(scope="module")
def syn(tmp_path_factory, manifest, golden):
    d = tmp_path_factory.mktemp("syn")
    g = golden("synth_small"); m = manifest["synth_small"]
    cols = {c: tuple(g[f"{c}_{k}"] for k in ("start", "stop", "mapq", "strand")) for c, _ in m["contigs"]}
    p = {"frag": write_frag_gz(d / "syn.frag.gz", cols), "cs": str(d / "syn.chrom.sizes"), "sites": str(d / "sites.bed"),
         "ivbed": str(d / "cov_iv.bed"), "dir": d}
    open(p["cs"], "w").write("".join(f"{c}\t{n}\n" for c, n in m["contigs"]))
    open(p["sites"], "w").write(m["sites_bed"])
    open(p["ivbed"], "w").write(m["cov_intervals_bed"])
    return p


def test_wps_fixture(fx, manifest, golden):
    import finaletoolkit_b200 as F
    g = golden("fixture17"); m = manifest["fixture17"]
    for i, c in enumerate(m["wps_cases"]):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            r = F.wps(fx["frag"], "12", c["start"], c["stop"], m["chrom_size"], **c["kwargs"])
        assert r.dtype.names == ("contig", "start", "wps") and r["wps"].dtype == np.int64
        assert np.array_equal(r["wps"], g[f"wps_{i}"])
        assert np.all(r["contig"] == "12") and np.array_equal(r["start"], np.arange(c["start"], max(c["stop"], c["start"])))
    with pytest.warns(UserWarning):
        assert F.wps(fx["frag"], "12", 5, 5, 100).shape == (0,)
    with pytest.raises(ValueError):
        F.wps(fx["frag"], "12", 1, 2, 100, fraction_low=100)
    with pytest.raises(TypeError):
        F.wps(fx["frag"], "12", 34444145, 34444155, m["chrom_size"], output_file=3)
    out = str(fx["dir"] / "w.wig")
    F.wps(fx["frag"], "12", 34444145, 34444155, m["chrom_size"], output_file=out, quality_threshold=0)
    assert open(out).read() == "fixedStep\tchrom=12\tstart=34444145\tstep=1\tspan=10\n" + "-1\n" * 5 + "1\n" * 5


def test_multi_wps(fx, syn, manifest, golden):
    import finaletoolkit_b200 as F
    from finaletoolkit_b200.io import bigwig
    m = manifest["fixture17"]
    c = m["multi_wps"][0]
    out = str(fx["dir"] / "cfg1.bed.gz")
    assert F.multi_wps(fx["frag"], fx["ivl"], chrom_sizes=fx["cs"], output_file=out, workers=1) == out
    txt = read_gz(out)
    assert hashlib.sha256(txt.encode()).hexdigest() == c["sha256_text"] and len(txt.splitlines()) == 7213
    with pytest.raises(ValueError):   # BED row with start > stop (frag/_multi_wps.py:258-263)
        F.multi_wps(fx["frag"], fx["ivl_ov"], chrom_sizes=fx["cs"], output_file=str(fx["dir"] / "x.bed.gz"), **m["multi_wps"][1]["kwargs"])
    with pytest.raises(ValueError):   # suffix rule
        F.multi_wps(fx["frag"], fx["ivl"], chrom_sizes=fx["cs"], output_file="-")
    with pytest.raises(ValueError):   # chrom_sizes mandatory for fragment files
        F.multi_wps(fx["frag"], fx["ivl"], output_file=out)
    # bigWig output holds the same values (float32) at the same positions
    bw = str(fx["dir"] / "cfg1.bw")
    F.multi_wps(fx["frag"], fx["ivl"], chrom_sizes=fx["cs"], output_file=bw)
    g = golden("fixture17")
    s, e, v = bigwig.open(bw).intervals_arrays("12", 34440000, 34450000)
    assert np.array_equal(s, g["bw_cfg1_pos"]) and np.array_equal(v, g["bw_cfg1_val"])
    # synthetic sites: overlaps, unknown contig (warning), header-order sort
    ms = manifest["synth_small"]
    for j, c in enumerate(ms["multi_wps"]):
        out = str(syn["dir"] / f"mw{j}.bed.gz")
        with pytest.warns(UserWarning):
            F.multi_wps(syn["frag"], syn["sites"], chrom_sizes=syn["cs"], output_file=out, **c["kwargs"])
        assert hashlib.sha256(read_gz(out).encode()).hexdigest() == c["sha256_text"], c


def test_coverage(fx, syn, manifest):
    import finaletoolkit_b200 as F
    m = manifest["fixture17"]
    for c in m["single_coverage"]:
        assert list(F.single_coverage(fx["frag"], **c["kwargs"])) == c["result"]
    for c in m["coverage"]:
        out = str(fx["dir"] / ("cov" + c["suffix"]))
        r = F.coverage(fx["frag"], fx["ivl"], out, **c["kwargs"])
        assert [list(x) for x in r] == c["results"]
        assert (read_gz(out) if out.endswith(".gz") else open(out).read()) == c["text"]
    assert [list(x) for x in F.coverage(fx["frag"], fx["ivl"], None)] == m["coverage"][0]["results"]
    with pytest.raises(ValueError):
        F.coverage(fx["frag"], fx["ivl"], "out.txt")
    from finaletoolkit_b200.exceptions import InvalidInputError
    with pytest.raises(InvalidInputError):
        F.single_coverage(fx["frag"], "12", 1, 2, intersect_policy="nope")
    with pytest.raises(InvalidInputError):
        F.single_coverage(fx["frag"], None, 5, 10)
    ms = manifest["synth_small"]
    for c in ms["coverage"]:
        out = str(syn["dir"] / "scov.bed")
        r = F.coverage(syn["frag"], syn["ivbed"], out, **c["kwargs"])
        assert [list(x) for x in r] == c["results"] and open(out).read() == c["text"]
    assert [list(F.single_coverage(syn["frag"])), list(F.single_coverage(syn["frag"], quality_threshold=0, min_length=200))] == ms["single_coverage_genome"]


def test_frag_length(fx, syn, manifest, golden):
    import finaletoolkit_b200 as F
    m = manifest["fixture17"]
    for c in m["frag_length"]:
        assert F.frag_length(fx["frag"], **c["kwargs"]).tolist() == c["lengths"]
    for src, mm in ((fx, m), (syn, manifest["synth_small"])):
        for c in mm["frag_length_bins"]:
            out = str(src["dir"] / "flb.tsv")
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                bins, counts = F.frag_length_bins(src["frag"], output_file=out, **c["kwargs"])
            assert np.asarray(bins).tolist() == c["bins"] and np.asarray(counts).tolist() == c["counts"]
            if c["text"] is not None and len(c["bins"]):
                assert open(out).read() == c["text"]
    with pytest.warns(RuntimeWarning):
        b, c_ = F.frag_length_bins(fx["frag"], contig="12", start=1, stop=2)
        assert len(b) == 0 and len(c_) == 0
    for src, mm, bed in ((fx, m, fx["ivl"]), (syn, manifest["synth_small"], syn["ivbed"])):
        for c in mm["frag_length_intervals"]:
            out = str(src["dir"] / "fli.bed")
            r = F.frag_length_intervals(src["frag"], bed, output_file=out, **c["kwargs"])
            assert [list(x) for x in r] == c["results"] and open(out).read() == c["text"]
    g = golden("synth_small")
    for c in manifest["synth_small"]["frag_length"]:
        assert np.array_equal(F.frag_length(syn["frag"], **c["kwargs"]), g[c["key"]])


def test_host_stream_helpers(fx, manifest):
    import finaletoolkit_b200 as F
    m = manifest["fixture17"]
    fa = F.frag_array(fx["frag"], "12", min_length=120, max_length=180)
    assert [[int(a), int(b), bool(c)] for a, b, c in fa.tolist()] == m["frag_array_120_180"]
    assert [list(x) for x in F.frag_generator(fx["frag"], "12", start=34443119, stop=34443538)] == m["frag_generator_detail"]
    with pytest.warns(UserWarning):
        from finaletoolkit_b200.io import fragments
        fragments._CACHE.clear()
        assert [list(x) for x in F.frag_generator(fx["bed6"], "12", start=34443119, stop=34443538)] == m["frag_generator_bed6"]
    assert sum(1 for _ in F.frag_generator(fx["frag"], None, quality_threshold=0)) == m["frag_generator_count_all"]


def test_end_motifs(tmp_path, manifest, golden):
    import finaletoolkit_b200 as F
    g = golden("motif"); m = manifest["motif"]
    sizes = dict(m["contigs"])
    cols = {c: tuple(g[f"{c}_{k}"] for k in ("start", "stop", "mapq", "strand")) for c in sizes}
    frag = write_frag_gz(tmp_path / "m.frag.gz", cols)
    tb = write_2bit(tmp_path / "m.2bit", [(c, *golden_codes(g, c, n)) for c, n in sizes.items()])
    for c in m["region_end_motifs"]:
        d = F.region_end_motifs(frag, c["contig"], c["start"], c["stop"], tb, **c["kwargs"])
        assert np.array_equal(np.array(list(d.values())), g[c["key"]]) and list(d)[:2] == F.gen_kmers(c["kwargs"].get("k", 4))[:2]
    for j, c in enumerate(m["end_motifs"]):
        out = str(tmp_path / f"em{j}.tsv")
        r = F.end_motifs(frag, tb, output_file=out, **c["kwargs"])
        assert np.array_equal(np.array(r.frequencies()), g[c["key"]])
        assert r.motif_diversity_score() == c["mds"] and open(out).read() == c["tsv"]
    for j, c in enumerate(m["interval_end_motifs"]):
        out = str(tmp_path / f"iem{j}.tsv")
        r = F.interval_end_motifs(frag, tb, [tuple(x) for x in m["intervals"]], output_file=out, **c["kwargs"])
        assert open(out).read() == c["tsv"]
        r.to_tsv(out, calc_freq=False); assert open(out).read() == c["tsv_counts"]
        r.mds_bed(out); assert open(out).read() == c["mds_bed"]
        got = [v for _, v in r.motif_diversity_score(miller_madow=True)]
        assert all((a == b) or (np.isnan(a) and np.isnan(b)) for a, b in zip(got, c["mds_mm"]))
    with pytest.raises(ValueError):
        F.region_end_motifs(frag, "chrM2", 0, 10, tb, negative_strand=True)
    bad = write_frag_gz(tmp_path / "bad.frag.gz", {"chrM2": (np.array([0, 50]), np.array([3, 220]), np.array([60, 60]), np.array([1, 0]))})
    with pytest.raises(RuntimeError):
        F.region_end_motifs(bad, "chrM2", 0, 1000, tb)
    # the reference's own golden TSVs: regional MDS and round trip (tests/test_end_motifs.py:170-247)
    f = manifest["fixture17"]
    p = str(tmp_path / "ivl_dif.tsv"); open(p, "w").write(f["end_motifs_intervals_dif_tsv"])
    emi = F.EndMotifsIntervals.from_file(p, 30, sep="\t")
    assert [[list(iv), v] for iv, v in emi.motif_diversity_score()] == f["regional_mds"]
    assert [[list(iv), v] for iv, v in emi.motif_diversity_score(miller_madow=True)] == f["regional_mds_mm"]
    p2 = str(tmp_path / "dif.tsv"); open(p2, "w").write(f["end_motifs_dif_tsv"])
    emf = F.EndMotifFreqs.from_file(p2, 30)
    assert emf.motif_diversity_score() == f["mds_from_dif_tsv"]
    emf.to_tsv(str(tmp_path / "rt.tsv")); assert open(tmp_path / "rt.tsv").read() == f["end_motifs_dif_tsv"]


def test_breakpoint_motifs(tmp_path, manifest, golden):
    """frag/_breakpoint_motifs.py public functions vs outputs of the reference itself."""
    import finaletoolkit_b200 as F
    g = golden("motif"); m = manifest["motif"]
    sizes = dict(m["contigs"])
    cols = {c: tuple(g[f"{c}_{k}"] for k in ("start", "stop", "mapq", "strand")) for c in sizes}
    frag = write_frag_gz(tmp_path / "m.frag.gz", cols)
    tb = write_2bit(tmp_path / "m.2bit", [(c, *golden_codes(g, c, n)) for c, n in sizes.items()])
    for c in m["region_breakpoint_motifs"]:
        d = F.region_breakpoint_motifs(frag, c["contig"], c["start"], c["stop"], tb, **c["kwargs"])
        assert np.array_equal(np.array(list(d.values())), g[c["key"]]) and list(d) == F.gen_kmers(c["kwargs"].get("k", 6))
    for j, c in enumerate(m["breakpoint_motifs"]):
        out = str(tmp_path / f"bm{j}.tsv")
        r = F.breakpoint_motifs(frag, tb, output_file=out, **c["kwargs"])
        assert isinstance(r, F.BreakpointMotifFreqs) and np.array_equal(np.array(r.frequencies()), g[c["key"]])
        assert r.motif_diversity_score() == c["mds"] and open(out).read() == c["tsv"]
    for j, c in enumerate(m["interval_breakpoint_motifs"]):
        out = str(tmp_path / f"ibm{j}.tsv")
        r = F.interval_breakpoint_motifs(frag, tb, [tuple(x) for x in m["intervals"]], output_file=out, **c["kwargs"])
        assert open(out).read() == c["tsv"]
        got = [v for _, v in r.motif_diversity_score()]
        assert all((a == b) or (np.isnan(a) and np.isnan(b)) for a, b in zip(got, c["mds"]))
    with pytest.raises(ValueError):
        F.region_breakpoint_motifs(frag, "chrM2", 0, 10, tb, negative_strand=True)


def test_adjust_wps(fx, syn, manifest, golden, tmp_path):
    import finaletoolkit_b200 as F
    from finaletoolkit_b200.io import bigwig
    # fixture: multi_wps -> .bw -> adjust_wps, vs the reference's float32 bigWig values
    g = golden("fixture17"); m = manifest["fixture17"]
    bw = str(tmp_path / "cfg1.bw")
    F.multi_wps(fx["frag"], fx["ivl"], chrom_sizes=fx["cs"], output_file=bw)
    for j, c in enumerate(m["adjust_wps_cases"]):
        bed = str(tmp_path / f"a{j}.bed"); open(bed, "w").write(c["bed"])
        out = str(tmp_path / f"a{j}.bw")
        F.adjust_wps(bw, bed, out, fx["cs"], **c["kwargs"])
        s, e, v = bigwig.open(out).intervals_arrays("12", 34440000, 34450000)
        assert np.array_equal(s, g[f"adj_{j}_pos"]) and np.array_equal(e, s + 1)
        np.testing.assert_allclose(v, g[f"adj_{j}_val_f32"], rtol=1e-5, atol=1e-6)
    # synthetic: tiled intervals (merge rule), mean, subtract_edges, no savgol, custom windows
    g = golden("adjust"); ma = manifest["adjust"]
    bed = str(tmp_path / "tile.bed"); open(bed, "w").write(ma["tile_bed"])
    bw = str(tmp_path / "tile.bw")
    F.multi_wps(syn["frag"], bed, chrom_sizes=syn["cs"], output_file=bw)
    s, e, v = bigwig.open(bw).intervals_arrays("chrB", 0, 80_000)
    assert np.array_equal(s, g["raw_pos"]) and np.array_equal(v, g["raw_val_f32"])
    for j, c in enumerate(ma["adjust_cases"]):
        b = str(tmp_path / f"s{j}.bed"); open(b, "w").write(c["bed"])
        out = str(tmp_path / f"s{j}.bw")
        F.adjust_wps(bw, b, out, syn["cs"], **c["kwargs"])
        r = bigwig.open(out).intervals_arrays("chrB", 0, 80_000)
        assert np.array_equal(r[0], g[f"adj_{j}_pos"])
        np.testing.assert_allclose(r[2], g[f"adj_{j}_val_f32"], rtol=1e-5, atol=1e-6)
    with pytest.raises(ValueError):
        F.adjust_wps(bw, str(tmp_path / "s0.bed"), str(tmp_path / "e.bw"), syn["cs"], median_window_size=100000)
    with pytest.raises(ValueError):
        F.adjust_wps(bw, "intervals.txt", str(tmp_path / "e.bw"), syn["cs"])


def test_bam_input(tmp_path, manifest, golden):
    """The reference's own BAM fixture (tests/data/12.3444.b37.bam) through the native BAM decoder:
    wps / single_coverage / frag_length_bins equal what the reference computes from the BAM."""
    import finaletoolkit_b200 as F
    g = golden("fixture17"); m = manifest["fixture17"]
    bam = str(tmp_path / "12.3444.b37.bam")
    open(bam, "wb").write(g["bam_file"].tobytes()); open(bam + ".bai", "wb").close()
    r = F.wps(bam, "12", 34442500, 34447500, 133851895)
    assert np.array_equal(r["wps"].astype(np.int64), g["bam_wps"]) and r["wps"].any()
    assert list(F.single_coverage(bam, "12", 34442500, 34447500, quality_threshold=30)) == m["bam_single_coverage"]
    bins, counts = F.frag_length_bins(bam, "12", 34442500, 34447500, bin_size=10)
    assert [np.asarray(bins).tolist(), np.asarray(counts).tolist()] == m["bam_frag_length_bins"]
    got = [list(x) for x in F.frag_generator(bam, "12", quality_threshold=0)]
    assert [x[1] for x in got] == np.sort(g["bam_start"]).tolist() and len(got) == 17


def test_multi_wps_streams_large_contigs(tmp_path, monkeypatch):
    """multi_wps sends a large contig through the streamed pipeline (host columns page-locked in place, chunked
    H2D || kernels || D2H, int16 scores in pinned memory) and a small one through the resident upload; the output
    text is byte-identical to the all-resident run, and spot intervals equal the oracle."""
    import finaletoolkit_b200 as F
    from finaletoolkit_b200.frag import _multi_wps, _wps
    from finaletoolkit_b200.io.fragments import FragmentTable
    from finaletoolkit_b200.synth import synth_fragments
    from oracle import oracle as O
    big_len, small_len = 8_000_000, 300_000
    cols = {"1": synth_fragments(big_len, 2_600_000, 3, seed_base=880), "2": synth_fragments(small_len, 90_000, 4, seed_base=880)}
    table = FragmentTable(cols)
    cs = tmp_path / "cs.txt"; cs.write_text(f"1\t{big_len}\n2\t{small_len}\n")
    rng = np.random.default_rng(12)
    rows = [("1", int(p)) for p in np.sort(rng.integers(3000, big_len - 3000, 400))] + \
           [("2", int(p)) for p in np.sort(rng.integers(3000, small_len - 3000, 20))]
    sites = tmp_path / "sites.bed"
    sites.write_text("".join(f"{c}\t{p}\t{p + 1}\tx\t0\t+\n" for c, p in rows))
    out_a, out_b = str(tmp_path / "a.bed.gz"), str(tmp_path / "b.bed.gz")
    F.multi_wps(table, str(sites), chrom_sizes=str(cs), output_file=out_a)
    assert _multi_wps.LAST_TIMINGS["streamed_contigs"] == 1
    monkeypatch.setattr(_wps, "_STREAM_MIN_FRAGMENTS", 1 << 40)
    F.multi_wps(table, str(sites), chrom_sizes=str(cs), output_file=out_b)
    assert _multi_wps.LAST_TIMINGS["streamed_contigs"] == 0
    ta, tb = read_gz(out_a), read_gz(out_b)
    assert ta == tb and len(ta.splitlines()) > 1_000_000
    # spot check against the oracle: the first window of contig 1
    p = rows[0][1]
    exp = O.wps_interval(O.Frags(*cols["1"]), p - 2500, p + 2500, big_len, 120, 120, 180, 30)
    got = [int(l.split("\t")[3]) for l in ta.splitlines()[:5000]]
    assert ta.splitlines()[0].startswith(f"1\t{p - 2500}\t") and np.array_equal(np.array(got[: len(exp)]), exp)
