"""Pin the CPU oracle (oracle/) against outputs of the UNMODIFIED reference.

tests/golden/* was produced by oracle/make_golden.py running the reference
package itself; here every oracle function must reproduce those outputs
exactly (integers) or to fp64 round-off (adjust_wps).
"""
import numpy as np
import pytest

from oracle import oracle as O


def frags_of(g, prefix=""):
    return O.Frags(g[prefix + "start"], g[prefix + "stop"], g[prefix + "mapq"], g[prefix + "strand"])


def test_fixture_wps(manifest, golden):
    g = golden("fixture17")
    fr = frags_of(g)
    m = manifest["fixture17"]
    for i, c in enumerate(m["wps_cases"]):
        kw = {"window_size": 120, "min_length": 120, "max_length": 180, "quality_threshold": 30, **c["kwargs"]}
        got = O.wps_interval(fr, c["start"], c["stop"], m["chrom_size"], **kw)
        assert np.array_equal(got, g[f"wps_{i}"]), c
    # the reference's own known answer, tests/test_wps.py:18-26
    assert O.wps_interval(fr, 34444145, 34444155, 133851895, quality_threshold=0).tolist() == [-1] * 5 + [1] * 5


def test_fixture_multi_wps_config1(manifest, golden):
    g = golden("fixture17")
    m = manifest["fixture17"]
    sizes = dict((l.split("\t")[0], int(l.split("\t")[1])) for l in m["chrom_sizes"].splitlines() if l)
    sites = O.read_sites(m["intervals_bed"].splitlines(), 5000, sizes)
    assert sites == [("12", 34440828, 34443041), ("12", 34443041, 34448041)]
    fr = frags_of(g)
    out, off = O.wps_intervals(fr, [s[1] for s in sites], [s[2] for s in sites], sizes["12"], threads=2)
    pos = np.concatenate([np.arange(s[1], s[2]) for s in sites])
    assert np.array_equal(pos, g["mwps_cfg1_pos"])
    assert np.array_equal(out, g["mwps_cfg1_score"])
    with pytest.raises(ValueError):
        O.read_sites(m["intervals_overlapped_bed"].splitlines(), 400, sizes)


def test_fixture_coverage_and_lengths(manifest, golden):
    g = golden("fixture17")
    fr = frags_of(g)
    m = manifest["fixture17"]
    for c in m["single_coverage"]:
        kw = dict(c["kwargs"]); kw.pop("contig", None)
        assert O.single_coverage(fr, **kw) == c["result"][4], c
    for c in m["frag_length"]:
        kw = dict(c["kwargs"]); kw.pop("contig", None)
        assert O.frag_lengths(fr, **kw).tolist() == c["lengths"]
    for c in m["frag_length_bins"]:
        kw = dict(c["kwargs"]); kw.pop("contig", None)
        bs = kw.pop("bin_size", 1); kw.pop("summary_stats", None); kw.pop("short_fraction", None)
        kw.setdefault("min_length", 0)
        d = O.length_dist(fr, **kw)
        if not d:
            assert c["bins"] == []
            continue
        bins, counts = O.length_bins(d, bs)
        assert bins.tolist() == c["bins"] and counts == c["counts"]
    ivs = [l.split("\t") for l in m["intervals_bed"].splitlines()]
    for c in m["frag_length_intervals"]:
        kw = dict(c["kwargs"]); sr = kw.pop("short_reads", 150); kw.setdefault("min_length", 0)
        for iv, exp in zip(ivs, c["results"]):
            d = O.length_dist(fr, int(iv[1]), int(iv[2]), **kw)
            assert list(O.length_stats(d, sr)) == exp[4:], (iv, exp)


def test_synth_wps(manifest, golden):
    g = golden("synth_small")
    m = manifest["synth_small"]
    sizes = dict(m["contigs"])
    frs = {c: frags_of(g, c + "_") for c in sizes}
    for c in m["wps_cases"]:
        kw = {"window_size": 120, "min_length": 120, "max_length": 180, "quality_threshold": 30, **c["kwargs"]}
        got = O.wps_interval(frs[c["contig"]], c["start"], c["stop"], sizes[c["contig"]], **kw)
        assert np.array_equal(got, g[c["key"]]), c


def test_synth_multi_wps(manifest, golden):
    g = golden("synth_small")
    m = manifest["synth_small"]
    sizes = dict(m["contigs"])
    frs = {c: frags_of(g, c + "_") for c in sizes}
    for j, c in enumerate(m["multi_wps"]):
        kw = {"window_size": 120, "interval_size": 5000, "min_length": 120, "max_length": 180, **c["kwargs"]}
        sites = O.read_sites(m["sites_bed"].splitlines(), kw.pop("interval_size"), sizes)
        pos, score = [], []
        for contig, s, e in sites:
            pos.append(np.arange(s, e))
            score.append(O.wps_interval(frs[contig], s, e, sizes[contig], **kw))
        assert np.array_equal(np.concatenate(pos), g[f"mwps_{j}_pos"])
        assert np.array_equal(np.concatenate(score), g[f"mwps_{j}_score"])
        runs = []
        for contig, s, e in sites:
            if runs and runs[-1][0] == contig:
                runs[-1][1] += e - s
            else:
                runs.append([contig, e - s])
        assert runs == c["runs"]


def _parse_ivs(text):
    out = []
    for line in text.splitlines(keepends=True):
        if line.startswith(("#", "track", "browser")) or not line.strip():
            continue
        p = line.strip().split("\t")
        if len(p) < 3:
            continue
        out.append((p[0], int(p[1]), int(p[2]), p[3] if len(p) > 3 else "."))
    return out


def test_synth_coverage_lengths(manifest, golden):
    g = golden("synth_small")
    m = manifest["synth_small"]
    sizes = dict(m["contigs"])
    frs = {c: frags_of(g, c + "_") for c in sizes}
    ivs = _parse_ivs(m["cov_intervals_bed"])
    for c in m["coverage"]:
        kw = dict(c["kwargs"]); norm = kw.pop("normalize", False); sf = kw.pop("scale_factor", 1.0)
        if norm:
            total = sum(O.single_coverage(fr, 0, None, **kw) for fr in frs.values())
            sf /= total
        for iv, exp in zip(ivs, c["results"]):
            cov = O.single_coverage(frs[iv[0]], iv[1], iv[2], **kw)
            assert cov * sf == exp[4], (iv, exp)
    assert sum(O.single_coverage(fr) for fr in frs.values()) == m["single_coverage_genome"][0][4]
    for c in m["frag_length_bins"]:
        kw = dict(c["kwargs"]); contig = kw.pop("contig", None)
        bs = kw.pop("bin_size", 1); kw.pop("summary_stats", None); sfrac = kw.pop("short_fraction", None)
        kw.setdefault("min_length", 0)
        d = O.merge_dists(O.length_dist(frs[cc], **kw) for cc in ([contig] if contig else sizes))
        bins, counts = O.length_bins(d, bs)
        assert bins.tolist() == c["bins"] and counts == c["counts"]
        if "#mean" in c["text"]:
            st = O.length_stats(d, sfrac if sfrac is not None else 0)
            lines = dict(l[1:].split(": ") for l in c["text"].splitlines() if l.startswith("#"))
            assert repr(st[0]) == lines["mean"] and repr(st[1]) == lines["median"] and repr(st[2]) == lines["stdev"]
    for c in m["frag_length"]:
        kw = dict(c["kwargs"]); contig = kw.pop("contig", None)
        got = np.concatenate([O.frag_lengths(frs[cc], **kw) for cc in ([contig] if contig else sizes)])
        assert np.array_equal(got, g[c["key"]])
    for c in m["frag_length_intervals"]:
        kw = dict(c["kwargs"]); sr = kw.pop("short_reads", 150); kw.setdefault("min_length", 0)
        for iv, exp in zip(ivs, c["results"]):
            d = O.length_dist(frs[iv[0]], iv[1], iv[2], **kw)
            assert list(O.length_stats(d, sr)) == exp[4:], (iv, exp)


def _seq_ascii(g, name, n):
    codes = np.unpackbits(g[f"{name}_codes_packed"]).reshape(-1, 2)[:n]
    codes = codes[:, 0] * 2 + codes[:, 1]
    seq = np.frombuffer(b"ACGT", np.uint8)[codes].copy()
    seq[np.unpackbits(g[f"{name}_nmask_packed"])[:n].astype(bool)] = ord("N")
    return seq.tobytes()


def test_motifs(manifest, golden):
    g = golden("motif")
    m = manifest["motif"]
    sizes = dict(m["contigs"])
    frs = {c: frags_of(g, c + "_") for c in sizes}
    seqs = {c: _seq_ascii(g, c, n) for c, n in sizes.items()}
    for c in m["region_end_motifs"]:
        kw = dict(c["kwargs"])
        got = O.region_end_motifs(frs[c["contig"]], seqs[c["contig"]], c["start"], c["stop"], **kw)
        assert np.array_equal(got, g[c["key"]]), c
    for c in m["end_motifs"]:
        kw = dict(c["kwargs"]); kw.setdefault("quality_threshold", 30); k = kw.get("k", 4)
        tot = np.zeros(4 ** k, np.float64)
        for contig, n in sizes.items():
            for s, e in O.genome_windows(n):
                tot = tot + O.region_end_motifs(frs[contig], seqs[contig], s, e, **kw)
        freq = tot / np.sum(tot)
        assert np.array_equal(freq, g[c["key"]])
        assert O.mds(freq, k) == c["mds"]
    for j, c in enumerate(m["interval_end_motifs"]):
        kw = dict(c["kwargs"]); kw.setdefault("quality_threshold", 30); k = kw.get("k", 4)
        rows = [O.region_end_motifs(frs[iv[0]], seqs[iv[0]], iv[1], iv[2], **kw) for iv in m["intervals"]]
        assert np.array_equal(np.array(rows), g[c["key"]])
        for r, e, emm in zip(rows, c["mds"], c["mds_mm"]):
            with np.errstate(invalid="ignore", divide="ignore"):
                got = O.mds(r / np.sum(r), k)
                gotmm = O.mds(r / np.sum(r), k, True, np.sum(r))
            assert got == e or (np.isnan(got) and np.isnan(e))
            assert gotmm == emm or (np.isnan(gotmm) and np.isnan(emm))
    bad = O.Frags([0, 50], [3, 220], [60, 60], [1, 0])
    with pytest.raises(RuntimeError):
        O.region_end_motifs(bad, seqs["chrM2"], 0, 1000)
    # reference golden: regional MDS of tests/data/end_motifs/end_motifs_intervals_dif.tsv
    f = manifest["fixture17"]
    lines = f["end_motifs_intervals_dif_tsv"].splitlines()
    for line, (iv, exp), (_, expmm) in zip(lines[1:], f["regional_mds"], f["regional_mds_mm"]):
        p = line.split("\t")
        vals = np.array([float(x) for x in p[5:]])
        assert O.mds(vals / vals.sum(), 4) == pytest.approx(exp, rel=1e-15)
        assert O.mds(vals / vals.sum(), 4, True, float(p[4])) == pytest.approx(expmm, rel=1e-15)
    assert f["regional_mds"][0][1] == pytest.approx(0.5844622669209985, rel=1e-6)


def test_breakpoint_motifs(manifest, golden):
    """oracle restatement of frag/_breakpoint_motifs.py vs outputs of the reference itself."""
    g = golden("motif")
    m = manifest["motif"]
    sizes = dict(m["contigs"])
    frs = {c: frags_of(g, c + "_") for c in sizes}
    seqs = {c: _seq_ascii(g, c, n) for c, n in sizes.items()}
    nonzero = 0
    for c in m["region_breakpoint_motifs"]:
        got = O.region_breakpoint_motifs(frs[c["contig"]], seqs[c["contig"]], c["start"], c["stop"], **c["kwargs"])
        assert np.array_equal(got, g[c["key"]]), c
        nonzero += int(got.sum() > 0)
        if c["kwargs"].get("k", 6) % 2:          # odd k: the reference's length check rejects every window
            assert got.sum() == 0
    assert nonzero >= 6
    for c in m["breakpoint_motifs"]:
        kw = dict(c["kwargs"]); k = kw.get("k", 6)
        tot = np.zeros(4 ** k, np.float64)
        for contig, n in sizes.items():
            for s, e in O.genome_windows(n):
                tot = tot + O.region_breakpoint_motifs(frs[contig], seqs[contig], s, e, **kw)
        freq = tot / np.sum(tot)
        assert np.array_equal(freq, g[c["key"]])
        assert O.mds(freq, k) == c["mds"]
    for c in m["interval_breakpoint_motifs"]:
        kw = dict(c["kwargs"])
        rows = [O.region_breakpoint_motifs(frs[iv[0]], seqs[iv[0]], iv[1], iv[2], **kw) for iv in m["intervals"]]
        assert np.array_equal(np.array(rows), g[c["key"]])


def test_delfi_windows(manifest, golden):
    """oracle restatement of _delfi_single_window (frag/_delfi.py:404-511) vs the reference's tuples."""
    from helpers import delfi_tracks
    g = golden("delfi"); m = manifest["delfi"]
    sizes = dict(m["contigs"])
    frs = {c: frags_of(g, c + "_") for c in sizes}
    seqs = {c: _seq_ascii(g, c, n) for c, n in sizes.items()}
    bl, gaps = delfi_tracks(m)
    for case in m["single_window"]:
        k = case["key"]
        for j, (c, a, b) in enumerate(m["bins_list"]):
            got = O.delfi_window(frs[c], seqs[c], c, a, b, blacklist=bl.get(c) if case["use_blacklist"] else None,
                                 gaps=gaps.get(c) if case["use_gaps"] else None, quality_threshold=case["quality_threshold"])
            exp = (c, a, b, case["arms"][j], g[k + "_short"][j], g[k + "_long"][j], g[k + "_gc"][j], g[k + "_num"][j])
            assert got[:4] == exp[:4] and got[7] == exp[7], (k, c, a, b)
            for x, y in zip(got[4:7], exp[4:7]):
                assert (np.isnan(x) and np.isnan(y)) or x == y, (k, c, a, b, got, exp)
    # the blacklist and the gap rule both removed fragments somewhere, and the invalid bin has gc 0
    assert g["single_plain_num"].sum() > g["single_q30_num"].sum()
    j = m["bins_list"].index(["chr21", 118500, 121000])
    assert g["single_plain_gc"][j] == 0.0 and g["single_plain_num"][j] > 0


def test_agg_bw(manifest, golden):
    """oracle restatement of agg_bw's accumulation (utils/_agg_bw.py:84-126) vs the reference's output."""
    from helpers import agg_fixture
    g = golden("agg"); m = manifest["agg"]
    signals, strands, _ = agg_fixture(g, m)
    assert sum(v is None for v in signals) == 2          # past-the-end and unknown-contig intervals
    for c in m["cases"]:
        with np.errstate(invalid="ignore", divide="ignore"):
            got = O.agg_bw_core(signals, strands, **c["kwargs"])
        exp = g[c["key"]]
        assert str(got.dtype) == c["dtype"] and got.shape == exp.shape
        assert np.array_equal(got, exp, equal_nan=True), c["kwargs"]


def test_adjust_core(manifest, golden):
    g = golden("adjust")
    m = manifest["adjust"]
    for c in m["core_cases"]:
        x = g[c["input"]]
        sg = c["sg"]
        got = O.adjust_core(x, c["w"], c["mean"], sg is not None, *(sg or (21, 2)))
        assert np.array_equal(got, g[c["key"] + "_out"]), c
        pre = O.local_filter(x, c["w"], c["mean"])
        np.testing.assert_allclose(pre, g[c["key"] + "_pre"], rtol=1e-13, atol=1e-13)
        if not c["mean"]:
            assert np.array_equal(pre, g[c["key"] + "_pre"])


def test_adjust_driver(manifest, golden):
    g = golden("adjust")
    m = manifest["adjust"]
    raw_pos, raw_val = g["raw_pos"], g["raw_val_f32"].astype(np.float64)
    lut = {int(p): i for i, p in enumerate(raw_pos.tolist())}
    for j, c in enumerate(m["adjust_cases"]):
        kw = {"interval_size": 5000, "median_window_size": 1000, "savgol_window_size": 21, "savgol_poly_deg": 2,
              "savgol": True, "mean": False, "subtract_edges": False, "edge_size": 500, **c["kwargs"]}
        sites = O.adjust_sites(c["bed"].splitlines(keepends=True), kw["interval_size"], kw["median_window_size"])
        pos_all, val_all = [], []
        sizes = dict(manifest["synth_small"]["contigs"])
        for contig, s, e in sites:
            if e > sizes[contig]:
                continue  # pyBigWig: "Invalid interval bounds!" -> RuntimeError -> interval skipped (:145-153)
            idx = [lut[p] for p in range(s, e) if p in lut]
            if not idx:
                continue
            x = raw_val[idx].copy()
            p = raw_pos[idx]
            if kw["subtract_edges"]:
                x = x - np.mean([np.mean(x[:kw["edge_size"]]), np.mean(x[-kw["edge_size"]:])])
            y = O.adjust_core(x, kw["median_window_size"], kw["mean"], kw["savgol"], kw["savgol_window_size"], kw["savgol_poly_deg"])
            w = kw["median_window_size"]
            pos_all.append(p[w // 2: -(w // 2)]); val_all.append(y)
        pos = np.concatenate(pos_all); val = np.concatenate(val_all).astype(np.float32)
        assert np.array_equal(pos, g[f"adj_{j}_pos"])
        assert np.array_equal(val, g[f"adj_{j}_val_f32"])


def test_cleavage(manifest, golden):
    g = golden("cleavage"); m = manifest["cleavage"]
    fx = golden("fixture17"); sy = golden("synth_small")
    for j, c in enumerate(m["cases"]):
        fr = frags_of(fx) if c["src"] == "fixture" else frags_of(sy, c["contig"] + "_")
        pos, prop = O.cleavage_profile(fr, c["chrom_size"], c["start"], c["stop"], **c["kwargs"])
        assert np.array_equal(pos, g[f"clv_{j}_pos"]) and np.array_equal(prop, g[f"clv_{j}_prop"]), c
    sizes = dict(manifest["synth_small"]["contigs"])
    assert O.cleavage_intervals(m["bed"].splitlines(), 0, 0, sizes)[:2] == [("chrA", 1000, 2500), ("chrA", 9000, 9001)]
