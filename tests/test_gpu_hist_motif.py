"""CUDA interval counts / length histograms / raw lengths / end motifs vs goldens + oracle."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from finaletoolkit_b200.device import require_cuda
    return require_cuda("cuda:0")


def _dev_frags(g, prefix, dev):
    from finaletoolkit_b200.device import ContigFragments
    return ContigFragments(g[prefix + "start"], g[prefix + "stop"], g[prefix + "mapq"], g[prefix + "strand"], device=dev)


def _parse_ivs(text):
    out = []
    for line in text.splitlines(keepends=True):
        if line.startswith(("#", "track", "browser")) or not line.strip():
            continue
        p = line.strip().split("\t")
        if len(p) >= 3:
            out.append((p[0], int(p[1]), int(p[2]), p[3] if len(p) > 3 else "."))
    return out


def _dict_from(hist_row, first_row):
    """Rebuild the reference's length->count dict in first-seen order."""
    nz = np.flatnonzero(hist_row)
    order = np.argsort(first_row[nz], kind="stable")
    return {int(nz[i]): int(hist_row[nz[i]]) for i in order}


def test_fixture_counts_and_lengths(manifest, golden, dev):
    from finaletoolkit_b200 import device as D
    g = golden("fixture17"); m = manifest["fixture17"]
    fr = _dev_frags(g, "", dev)
    for c in m["single_coverage"]:
        kw = dict(c["kwargs"]); kw.pop("contig", None)
        s, e = kw.pop("start", 0), kw.pop("stop", None)
        cnt, _, _ = D.interval_hist(fr, [s], [e], **kw)
        assert int(cnt[0]) == c["result"][4], c
    for c in m["frag_length"]:
        kw = dict(c["kwargs"]); kw.pop("contig", None)
        assert D.frag_lengths(fr, **kw).cpu().tolist() == c["lengths"]


def test_synth_coverage(manifest, golden, dev):
    from finaletoolkit_b200 import device as D
    g = golden("synth_small"); m = manifest["synth_small"]
    sizes = dict(m["contigs"])
    frs = {c: _dev_frags(g, c + "_", dev) for c in sizes}
    ivs = _parse_ivs(m["cov_intervals_bed"])
    for c in m["coverage"]:
        kw = dict(c["kwargs"]); norm = kw.pop("normalize", False); sf = kw.pop("scale_factor", 1.0)
        if norm:
            total = sum(int(D.interval_hist(fr, [0], [None], **kw)[0][0]) for fr in frs.values())
            sf /= total
        for contig in sizes:
            sel = [i for i, iv in enumerate(ivs) if iv[0] == contig]
            cnt, _, _ = D.interval_hist(frs[contig], [ivs[i][1] for i in sel], [ivs[i][2] for i in sel], **kw)
            cnt = cnt.cpu().numpy()
            for j, i in enumerate(sel):
                assert int(cnt[j]) * sf == c["results"][i][4], (ivs[i], c["kwargs"])


def test_synth_length_dists(manifest, golden, dev):
    from finaletoolkit_b200 import device as D
    g = golden("synth_small"); m = manifest["synth_small"]
    sizes = dict(m["contigs"])
    frs = {c: _dev_frags(g, c + "_", dev) for c in sizes}
    ofrs = {c: O.Frags(g[c + "_start"], g[c + "_stop"], g[c + "_mapq"], g[c + "_strand"]) for c in sizes}
    # genome-wide / region dicts incl. first-seen order -> exact reference statistics text
    for c in m["frag_length_bins"]:
        kw = dict(c["kwargs"]); contig = kw.pop("contig", None)
        bs = kw.pop("bin_size", 1); kw.pop("summary_stats", None); sfrac = kw.pop("short_fraction", None)
        kw.setdefault("min_length", 0)
        s, e = kw.pop("start", None), kw.pop("stop", None)
        dicts = []
        for cc in ([contig] if contig else sizes):
            nb = frs[cc].max_len + 1
            _, h, f = D.interval_hist(frs[cc], [s], [e], n_bins=nb, pooled=True, first_seen=True, **kw)
            d = _dict_from(h[0].cpu().numpy(), f[0].cpu().numpy())
            assert d == O.length_dist(ofrs[cc], s, e, **kw) and list(d) == list(O.length_dist(ofrs[cc], s, e, **kw))
            dicts.append(d)
        d = O.merge_dists(dicts)
        bins, counts = O.length_bins(d, bs)
        assert bins.tolist() == c["bins"] and counts == c["counts"]
    # per-interval dicts (frag_length_intervals)
    ivs = _parse_ivs(m["cov_intervals_bed"])
    for c in m["frag_length_intervals"]:
        kw = dict(c["kwargs"]); sr = kw.pop("short_reads", 150); kw.setdefault("min_length", 0)
        for contig in sizes:
            sel = [i for i, iv in enumerate(ivs) if iv[0] == contig]
            nb = frs[contig].max_len + 1
            cnt, h, f = D.interval_hist(frs[contig], [ivs[i][1] for i in sel], [ivs[i][2] for i in sel],
                                        n_bins=nb, first_seen=True, **kw)
            h = h.cpu().numpy(); f = f.cpu().numpy()
            for j, i in enumerate(sel):
                d = _dict_from(h[j], f[j])
                assert list(O.length_stats(d, sr)) == c["results"][i][4:], ivs[i]
                assert int(cnt[j]) == sum(d.values())
    for c in m["frag_length"]:
        kw = dict(c["kwargs"]); contig = kw.pop("contig", None)
        got = np.concatenate([D.frag_lengths(frs[cc], **kw).cpu().numpy() for cc in ([contig] if contig else sizes)])
        assert np.array_equal(got, g[c["key"]])


def _codes(g, name, n):
    codes = np.unpackbits(g[f"{name}_codes_packed"]).reshape(-1, 2)[:n]
    return (codes[:, 0] * 2 + codes[:, 1]).astype(np.uint8), np.unpackbits(g[f"{name}_nmask_packed"])[:n].astype(bool)


def test_motifs_golden(manifest, golden, dev):
    from finaletoolkit_b200 import device as D
    g = golden("motif"); m = manifest["motif"]
    sizes = dict(m["contigs"])
    frs = {c: _dev_frags(g, c + "_", dev) for c in sizes}
    refs = {c: D.PackedContig.from_codes(*_codes(g, c, n), device=dev) for c, n in sizes.items()}

    def mode(kw):
        return 0 if kw.get("both_strands", True) else (2 if kw.get("negative_strand") else 1)

    for c in m["region_end_motifs"]:
        kw = c["kwargs"]
        got = D.end_motif_hist(frs[c["contig"]], refs[c["contig"]], [c["start"]], [c["stop"]], k=kw.get("k", 4),
                               strand_mode=mode(kw), quality_threshold=kw.get("quality_threshold", 20))
        assert np.array_equal(got[0].cpu().numpy(), g[c["key"]]), c
    for c in m["end_motifs"]:
        kw = c["kwargs"]; k = kw.get("k", 4)
        tot = None
        for contig, n in sizes.items():
            w = O.genome_windows(n)
            tot = D.end_motif_hist(frs[contig], refs[contig], [a for a, _ in w], [b for _, b in w], k=k, strand_mode=mode(kw),
                                   quality_threshold=kw.get("quality_threshold", 30), pooled=True, counts=tot)
        cc = tot[0].cpu().numpy().astype(np.float64)
        assert np.array_equal(cc / np.sum(cc), g[c["key"]])
    for c in m["interval_end_motifs"]:
        kw = c["kwargs"]; k = kw.get("k", 4)
        ivs = m["intervals"]
        rows = np.zeros((len(ivs), 4 ** k), np.int64)
        for contig in sizes:
            sel = [i for i, iv in enumerate(ivs) if iv[0] == contig]
            got = D.end_motif_hist(frs[contig], refs[contig], [ivs[i][1] for i in sel], [ivs[i][2] for i in sel], k=k,
                                   strand_mode=mode(kw), quality_threshold=kw.get("quality_threshold", 30))
            rows[sel] = got.cpu().numpy()
        assert np.array_equal(rows, g[c["key"]])
    from finaletoolkit_b200.device import ContigFragments
    bad = ContigFragments(np.array([0, 50], np.int32), np.array([3, 220], np.int32), np.array([60, 60], np.uint8),
                          np.array([1, 0], np.uint8), device=dev)
    with pytest.raises(RuntimeError):
        D.end_motif_hist(bad, refs["chrM2"], [0], [1000])
    # breakpoint motifs (frag/_breakpoint_motifs.py): same kernel, windows centred on the breakpoints
    for c in m["region_breakpoint_motifs"]:
        kw = c["kwargs"]
        got = D.end_motif_hist(frs[c["contig"]], refs[c["contig"]], [c["start"]], [c["stop"]], k=kw.get("k", 6),
                               strand_mode=mode(kw), quality_threshold=kw.get("quality_threshold", 30), breakpoint=True)
        assert np.array_equal(got[0].cpu().numpy(), g[c["key"]]), c
    for c in m["breakpoint_motifs"]:
        kw = c["kwargs"]; k = kw.get("k", 6)
        tot = None
        for contig, n in sizes.items():
            w = O.genome_windows(n)
            tot = D.end_motif_hist(frs[contig], refs[contig], [a for a, _ in w], [b for _, b in w], k=k, strand_mode=mode(kw),
                                   quality_threshold=kw.get("quality_threshold", 30), pooled=True, counts=tot, breakpoint=True)
        cc = tot[0].cpu().numpy().astype(np.float64)
        assert np.array_equal(cc / np.sum(cc), g[c["key"]])
    for c in m["interval_breakpoint_motifs"]:
        kw = c["kwargs"]; k = kw.get("k", 6)
        ivs = m["intervals"]
        rows = np.zeros((len(ivs), 4 ** k), np.int64)
        for contig in sizes:
            sel = [i for i, iv in enumerate(ivs) if iv[0] == contig]
            got = D.end_motif_hist(frs[contig], refs[contig], [ivs[i][1] for i in sel], [ivs[i][2] for i in sel], k=k,
                                   strand_mode=mode(kw), quality_threshold=kw.get("quality_threshold", 30), breakpoint=True)
            rows[sel] = got.cpu().numpy()
        assert np.array_equal(rows, g[c["key"]])


@pytest.mark.parametrize("seed", range(4))
def test_random_counts_vs_oracle(seed, dev):
    from finaletoolkit_b200 import device as D
    from finaletoolkit_b200.device import ContigFragments
    from finaletoolkit_b200.synth import synth_fragments, synth_twobit
    rng = np.random.default_rng(100 + seed)
    clen = int(rng.integers(50_000, 400_000)); n = int(rng.integers(1, 200_000))
    st, sp, mq, sd = synth_fragments(clen, n, seed, seed_base=777)
    ofr = O.Frags(st, sp, mq, sd); dfr = ContigFragments(st, sp, mq, sd, device=dev)
    ivs = [(int(s), int(s + rng.integers(0, 60_000))) for s in rng.integers(0, clen, 30)]
    ivs += [(0, None), (None, None), (clen // 2, None), (0, 0), (0, clen), (5, 6)]
    for pol in ("midpoint", "any"):
        lo = [None, 0, 100, 167][seed]; hi = [None, 10**9, 220, 167][seed]; q = [30, 0, 60, 1][seed]
        cnt, h, f = D.interval_hist(dfr, [a for a, _ in ivs], [b for _, b in ivs], pol, lo, hi, q,
                                    n_bins=ofr.max_len + 1, first_seen=True)
        cnt = cnt.cpu().numpy(); h = h.cpu().numpy(); f = f.cpu().numpy()
        for j, (a, b) in enumerate(ivs):
            exp = O.length_dist(ofr, a, b, lo, hi, pol, q)
            got = _dict_from(h[j], f[j])
            assert got == exp and list(got) == list(exp), (pol, a, b)
            assert cnt[j] == O.single_coverage(ofr, a, b, lo, hi, pol, q)
        a, b = ivs[seed]
        assert np.array_equal(D.frag_lengths(dfr, a, b, pol, quality_threshold=q).cpu().numpy(), O.frag_lengths(ofr, a, b, pol, q))
    # motifs, all strand modes and several k, against the oracle on a random genome
    codes, nm = synth_twobit(clen, seed, seed_base=888, telomere=500, block_len=3000)
    seq = np.frombuffer(b"ACGT", np.uint8)[codes].copy(); seq[nm] = ord("N")
    ref = D.PackedContig.from_codes(codes, nm, device=dev)
    keep = sp <= clen  # the reference raises when a fragment ends past the contig
    ofr2 = O.Frags(st[keep], sp[keep], mq[keep], sd[keep]); dfr2 = ContigFragments(st[keep], sp[keep], mq[keep], sd[keep], device=dev)
    for k, mode in [(4, 0), (1, 1), (3, 2), (6, 0), (7, 1), (2, 0)]:
        got = D.end_motif_hist(dfr2, ref, [a for a, b in ivs if b is not None and a is not None],
                               [b for a, b in ivs if b is not None and a is not None], k=k, strand_mode=mode, quality_threshold=q)
        got = got.cpu().numpy()
        j = 0
        for a, b in ivs:
            if a is None or b is None:
                continue
            exp = O.region_end_motifs(ofr2, seq.tobytes(), a, b, k, mode == 0, mode == 2, q)
            assert np.array_equal(got[j], exp), (k, mode, a, b)
            j += 1
    # breakpoint variant: fragments ending past the contig are legal here (that end is skipped)
    sp2 = sp.copy(); sp2[-3:] = clen + np.array([0, 1, 40])
    ofr3 = O.Frags(st, sp2, mq, sd); dfr3 = ContigFragments(st, sp2, mq, sd, device=dev)
    sel = [(a, b) for a, b in ivs if a is not None and b is not None]
    for k, mode in [(6, 0), (2, 1), (4, 2), (5, 0), (8, 0)]:
        got = D.end_motif_hist(dfr3, ref, [a for a, _ in sel], [b for _, b in sel], k=k, strand_mode=mode,
                               quality_threshold=q, breakpoint=True).cpu().numpy()
        for j, (a, b) in enumerate(sel):
            exp = O.region_breakpoint_motifs(ofr3, seq.tobytes(), a, b, k, mode == 0, mode == 2, q)
            assert np.array_equal(got[j], exp), (k, mode, a, b)


def test_fused_coverage_and_pooled_histogram(dev):
    """pooled="hist": per-interval counts + one pooled histogram in a single pass."""
    from finaletoolkit_b200 import device as D
    from finaletoolkit_b200.device import ContigFragments
    from finaletoolkit_b200.synth import synth_fragments
    clen, n = 1_500_000, 450_000
    st, sp, mq, sd = synth_fragments(clen, n, 11)
    fr = ContigFragments(st, sp, mq, sd, device=dev)
    nb = fr.max_len + 1
    edges = np.arange(0, clen + 5000, 5000).clip(max=clen)
    for ivs, kw in [((edges[:-1].tolist(), edges[1:].tolist()), dict()),
                    (([0, 1000, 1000, 700_000], [900_000, 5000, 5000, None]), dict(intersect_policy="any", min_length=100, max_length=400, quality_threshold=0))]:
        cnt, hist, first = D.interval_hist(fr, *ivs, n_bins=nb, pooled="hist", first_seen=True, **kw)
        c0, h0, f0 = D.interval_hist(fr, *ivs, n_bins=nb, first_seen=True, **kw)
        assert np.array_equal(cnt.cpu().numpy(), c0.cpu().numpy())
        assert np.array_equal(hist[0].cpu().numpy(), h0.sum(0).cpu().numpy())
        assert np.array_equal(first[0].cpu().numpy(), f0.min(0).values.cpu().numpy())
    # tiling intervals: the pooled histogram is the whole-contig one
    cnt, hist, _ = D.interval_hist(fr, edges[:-1].tolist(), edges[1:].tolist(), n_bins=nb, pooled="hist")
    _, hall, _ = D.interval_hist(fr, [0], [None], n_bins=nb, pooled=True)
    assert np.array_equal(hist.cpu().numpy(), hall.cpu().numpy()) and int(cnt.sum()) == int(hall.sum())
