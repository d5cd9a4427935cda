"""Size-independent properties at chromosome scale (the oracle is too slow to replay everything):
additivity / tiling invariance, checksums against closed forms or plain numpy reductions,
linearity over fragment subsets, shift invariance - plus oracle spot checks on random windows."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

CLEN, NFRAG = 60_000_000, 20_000_000


@pytest.fixture(scope="module")
def big():
    from finaletoolkit_b200.device import ContigFragments, require_cuda
    from finaletoolkit_b200.synth import synth_fragments
    dev = require_cuda("cuda:0")
    st, sp, mq, sd = synth_fragments(CLEN, NFRAG, 7)
    return dict(dev=dev, st=st, sp=sp, mq=mq, sd=sd, fr=ContigFragments(st, sp, mq, sd, device=dev), ofr=O.Frags(st, sp, mq, sd))


def test_coverage_additivity_and_numpy_checksum(big):
    from finaletoolkit_b200 import device as D
    fr, st, sp, mq = big["fr"], big["st"], big["sp"], big["mq"]
    L = (sp - st).astype(np.int64); mid = st.astype(np.int64) + L // 2
    for kw, mask in [(dict(), mq >= 30), (dict(min_length=120, max_length=180, quality_threshold=0), (L >= 120) & (L <= 180))]:
        whole = int(D.interval_hist(fr, [0], [None], **kw)[0][0])
        assert whole == int(mask.sum())
        for step in (5000, 1_000_003):
            edges = np.arange(0, CLEN + step, step).clip(max=CLEN)
            cnt = D.interval_hist(fr, edges[:-1].tolist(), edges[1:].tolist(), **kw)[0].cpu().numpy()
            assert int(cnt.sum()) == int((mask & (mid < CLEN)).sum())
            # every bin equals the numpy histogram of passing midpoints
            ref = np.bincount(np.minimum(mid[mask & (mid < CLEN)] // step, len(cnt) - 1), minlength=len(cnt))
            assert np.array_equal(cnt, ref)
    # "any" policy against numpy on a handful of overlapping intervals
    ivs = [(0, 1), (10_000_000, 10_000_500), (59_999_000, None), (123_456, 40_000_000)]
    cnt = D.interval_hist(fr, [a for a, _ in ivs], [b for _, b in ivs], "any", None, None, 30)[0].cpu().numpy()
    for (a, b), c in zip(ivs, cnt):
        m = (mq >= 30) & (sp > a) & ((st < b) if b is not None else True)
        assert int(c) == int(m.sum())
    # oracle spot checks
    rng = np.random.default_rng(1)
    for s in rng.integers(0, CLEN - 20_000, 20).tolist():
        e = s + int(rng.integers(1, 20_000))
        for pol in ("midpoint", "any"):
            got = int(D.interval_hist(fr, [s], [e], pol, 100, 220, 20)[0][0])
            assert got == O.single_coverage(big["ofr"], s, e, 100, 220, pol, 20)


def test_length_histogram_checksums(big):
    from finaletoolkit_b200 import device as D
    fr, st, sp, mq = big["fr"], big["st"], big["sp"], big["mq"]
    L = (sp - st).astype(np.int64)
    nb = fr.max_len + 1
    cnt, hist, first = D.interval_hist(fr, [0], [None], n_bins=nb, pooled=True, first_seen=True, quality_threshold=30)
    h = hist[0].cpu().numpy(); f = first[0].cpu().numpy()
    ok = mq >= 30
    assert np.array_equal(h, np.bincount(L[ok], minlength=nb)) and int(cnt[0]) == int(ok.sum()) == int(h.sum())
    vals, idx = np.unique(L[ok], return_index=True)
    assert np.array_equal(f[vals], np.flatnonzero(ok)[idx]) and np.all(f[h == 0] == 2 ** 31 - 1)
    # fused per-interval coverage + pooled histogram == separate passes
    edges = np.arange(0, CLEN + 5000, 5000).clip(max=CLEN)
    c2, h2, _ = D.interval_hist(fr, edges[:-1].tolist(), edges[1:].tolist(), n_bins=nb, pooled="hist", quality_threshold=30)
    assert np.array_equal(h2[0].cpu().numpy(), h) and int(c2.sum()) == int(h.sum())
    # raw lengths in stream order
    got = D.frag_lengths(fr, 1_000_000, 31_000_000, "midpoint", quality_threshold=30).cpu().numpy()
    mid = st.astype(np.int64) + L // 2
    m = ok & (mid >= 1_000_000) & (mid < 31_000_000)
    assert np.array_equal(got, L[m].astype(np.int32))


def test_end_motif_linearity_and_numpy(big):
    from finaletoolkit_b200 import device as D
    from finaletoolkit_b200.device import ContigFragments
    from finaletoolkit_b200.synth import synth_twobit
    dev = big["dev"]
    n = 4_000_000
    st, sp, mq, sd = big["st"][::5][:n], big["sp"][::5][:n], big["mq"][::5][:n], big["sd"][::5][:n]
    codes, nm = synth_twobit(CLEN, 3)
    ref = D.PackedContig.from_codes(codes, nm, device=dev)
    fr = ContigFragments(st, sp, mq, sd, device=dev)
    w = O.genome_windows(CLEN)
    ws, we = [a for a, _ in w], [b for _, b in w]
    k = 4
    full = D.end_motif_hist(fr, ref, ws, we, k=k, strand_mode=0, quality_threshold=30, pooled=True)[0].cpu().numpy()
    # linearity over a split of the fragments
    a = ContigFragments(st[0::2], sp[0::2], mq[0::2], sd[0::2], device=dev)
    b = ContigFragments(st[1::2], sp[1::2], mq[1::2], sd[1::2], device=dev)
    parts = sum(D.end_motif_hist(x, ref, ws, we, k=k, strand_mode=0, quality_threshold=30, pooled=True)[0].cpu().numpy() for x in (a, b))
    assert np.array_equal(full, parts)
    # both strands = forward ends of all fragments + reverse ends of all fragments
    allplus = ContigFragments(st, sp, mq, np.ones_like(sd), device=dev)
    fwd = D.end_motif_hist(allplus, ref, ws, we, k=k, strand_mode=1, quality_threshold=30, pooled=True)[0].cpu().numpy()
    rev = D.end_motif_hist(fr, ref, ws, we, k=k, strand_mode=2, quality_threshold=30, pooled=True)[0].cpu().numpy()
    assert np.array_equal(full, fwd + rev)
    # plain numpy restatement of the window semantics (fragments straddling a 1 Mb edge count twice)
    ok = mq >= 30
    s64, e64 = st.astype(np.int64)[ok], sp.astype(np.int64)[ok]
    # windows overlapped by [s, e) among the VISITED ones: CLEN is a multiple of 1 Mb, so the
    # reference never visits the last full window (frag/_motif_common.py:542,560-565)
    assert w[-1] == (CLEN, CLEN) and w[-2] == (CLEN - 2_000_000, CLEN - 1_000_000)
    last_visited = len(w) - 2
    mult = np.maximum(np.minimum((e64 - 1) // 1_000_000, last_visited) - s64 // 1_000_000 + 1, 0)
    pw = 4 ** np.arange(k - 1, -1, -1)
    fidx = sum(codes[s64 + j].astype(np.int64) * pw[j] for j in range(k))
    fbad = sum(nm[s64 + j] for j in range(k)) > 0
    ridx = sum((3 - codes[e64 - 1 - j].astype(np.int64)) * pw[j] for j in range(k))
    rbad = sum(nm[e64 - 1 - j] for j in range(k)) > 0
    exp = np.bincount(fidx[~fbad], weights=mult[~fbad], minlength=256) + np.bincount(ridx[~rbad], weights=mult[~rbad], minlength=256)
    assert np.array_equal(full, exp.astype(np.int64))
    # breakpoint motifs (k = 6, windows centred on the breakpoints): same checks, numpy restatement of
    # frag/_breakpoint_motifs.py:120-186 with the same double counting across window edges
    k, h = 6, 3
    bp = dict(k=k, quality_threshold=30, pooled=True, breakpoint=True)
    full = D.end_motif_hist(fr, ref, ws, we, strand_mode=0, **bp)[0].cpu().numpy()
    parts = sum(D.end_motif_hist(x, ref, ws, we, strand_mode=0, **bp)[0].cpu().numpy() for x in (a, b))
    assert np.array_equal(full, parts)
    fwd = D.end_motif_hist(allplus, ref, ws, we, strand_mode=1, **bp)[0].cpu().numpy()
    rev = D.end_motif_hist(fr, ref, ws, we, strand_mode=2, **bp)[0].cpu().numpy()
    assert np.array_equal(full, fwd + rev)
    pw = 4 ** np.arange(k - 1, -1, -1)
    near = (s64 >= h) & (s64 + h < CLEN)                      # else the fragment is skipped entirely
    fs_, fe_, mu = s64[near], e64[near], mult[near]
    fidx = sum(codes[fs_ - h + j].astype(np.int64) * pw[j] for j in range(k))
    fbad = sum(nm[fs_ - h + j] for j in range(k)) > 0
    rin = fe_ + h <= CLEN                                     # reverse window past the contig end: that end is skipped
    fe_r, mu_r = fe_[rin], mu[rin]
    ridx = sum((3 - codes[fe_r + h - 1 - j].astype(np.int64)) * pw[j] for j in range(k))
    rbad = sum(nm[fe_r + h - 1 - j] for j in range(k)) > 0
    exp = (np.bincount(fidx[~fbad], weights=mu[~fbad], minlength=4 ** k)
           + np.bincount(ridx[~rbad], weights=mu_r[~rbad], minlength=4 ** k))
    assert np.array_equal(full, exp.astype(np.int64)) and full.sum() > 1_000_000
    assert not D.end_motif_hist(fr, ref, ws, we, k=5, strand_mode=0, quality_threshold=30, pooled=True, breakpoint=True).any()


def test_adjust_shift_invariance_mean_closed_form_and_oracle(big):
    import torch
    from finaletoolkit_b200.device import WpsPlan, adjust_segments
    fr = big["fr"]
    n = 20_000_000
    wps = WpsPlan([0], [n], CLEN, 180, big["dev"]).run(fr).cpu().numpy().astype(np.float32)
    lens = [5000] * (n // 5000)
    w = 1000
    base, off = adjust_segments(wps, lens, savgol=False)
    base = base.cpu().numpy()
    shifted, _ = adjust_segments(wps + 37.0, lens, savgol=False)
    assert np.array_equal(base, shifted.cpu().numpy())                       # median filter commutes with +c
    assert np.all(base * 2 == np.rint(base * 2))                             # half-integers only
    mean, _ = adjust_segments(wps, lens, use_mean=True, savgol=False)
    x = wps.astype(np.float64).reshape(-1, 5000)
    cs = np.concatenate([np.zeros((x.shape[0], 1)), np.cumsum(x, axis=1)], axis=1)
    exp = x[:, w // 2: 5000 - w // 2] - (cs[:, w:5000] - cs[:, :5000 - w]) / w
    np.testing.assert_allclose(mean.cpu().numpy().reshape(-1, 4000), exp, rtol=0, atol=1e-9)
    # oracle (numpy + scipy, the reference's own calls) on random segments, full pipeline
    full, _ = adjust_segments(wps, lens)
    full = full.cpu().numpy()
    rng = np.random.default_rng(3)
    for k in rng.integers(0, len(lens), 25).tolist():
        seg = wps[k * 5000:(k + 1) * 5000].astype(np.float64)
        np.testing.assert_allclose(full[off[k]:off[k + 1]], O.adjust_core(seg), rtol=1e-5, atol=1e-9)
        assert np.array_equal(base[off[k]:off[k + 1]], O.local_filter(seg, w))
    # one merged chromosome-scale segment == per-run decomposition invariance
    one, _ = adjust_segments(wps[:3_000_000], [3_000_000], savgol=False, run_len=4096)
    two, _ = adjust_segments(wps[:3_000_000], [3_000_000], savgol=False, run_len=1531)
    assert torch.equal(one, two)


def test_delfi_windows_partition_and_numpy(big):
    """DELFI bins tiling the contig: every bin equals numpy's histogram of the passing midpoints,
    G+C counts add up to the contig's, and gaps / blacklist only ever remove fragments."""
    from finaletoolkit_b200 import device as D
    from finaletoolkit_b200.synth import synth_twobit
    fr, st, sp, mq = big["fr"], big["st"], big["sp"], big["mq"]
    codes, nm = synth_twobit(CLEN, 3)
    ref = D.PackedContig.from_codes(codes, nm, device=big["dev"])
    L = (sp - st).astype(np.int64); mid = (st.astype(np.int64) + sp) // 2
    ok = (mq >= 30) & (L >= 100) & (L <= 220) & (mid < CLEN)
    step = 100_000
    ws = np.arange(0, CLEN, step, dtype=np.int64); we = np.minimum(ws + step, CLEN)
    got = D.delfi_windows(fr, ref, ws, we, quality_threshold=30).cpu().numpy()
    nb = len(ws)
    assert np.array_equal(got[:, 0], np.bincount(mid[ok & (L < 151)] // step, minlength=nb))
    assert np.array_equal(got[:, 1], np.bincount(mid[ok & (L >= 151)] // step, minlength=nb))
    assert np.array_equal(got[:, 2], got[:, 0] + got[:, 1])
    gc = ((codes == 1) | (codes == 2)) & ~nm
    assert np.array_equal(got[:, 3], np.add.reduceat(gc.astype(np.int64), ws))
    # a gap track + blacklist: monotone (never adds), and exact against the oracle on a few bins
    rng = np.random.default_rng(8)
    r0 = np.sort(rng.integers(0, CLEN - 5000, 300)); r1 = r0 + rng.integers(150, 5000, 300)
    order = np.lexsort((r1, r0)); blk = (r0[order], r1[order])
    gaps = ((29_000_000, 31_500_000), [(0, 10_000), (CLEN - 10_000, CLEN)])
    cut = D.delfi_windows(fr, ref, ws, we, blacklist=blk, gaps=gaps, quality_threshold=30).cpu().numpy()
    assert (cut[:, :3] <= got[:, :3]).all() and cut[:, 2].sum() < got[:, 2].sum() and np.array_equal(cut[:, 3], got[:, 3])
    seq = np.frombuffer(b"ACGT", np.uint8)[codes].copy(); seq[nm] = ord("N")
    for j in (0, 7, 289, 290, 314, nb - 1):
        assert tuple(cut[j]) == O.delfi_counts(big["ofr"], seq.tobytes(), int(ws[j]), int(we[j]), blk, gaps, 30)


def test_agg_signal_linearity_and_numpy():
    """Strand-aware aggregation at scale: integer rows make the fp64 sum order-free, so the kernel
    must equal numpy's column sums exactly; a flipped '-' row equals the '+' row reversed."""
    from finaletoolkit_b200 import device as D
    rng = np.random.default_rng(12)
    n_seg, row_len, mws = 20_000, 5000, 120
    rows = rng.integers(-90, 40, (n_seg, row_len)).astype(np.float32)
    strand = rng.choice(np.array([1, -1, 0], np.int8), n_seg)
    lo, isz = mws // 2, row_len - mws
    got = D.agg_signal(rows, strand, lo, isz).cpu().numpy()
    core = rows[:, lo: lo + isz].astype(np.float64)
    exp = core[strand == 1].sum(0) + core[strand == -1][:, ::-1].sum(0)
    assert np.array_equal(got, exp)
    assert np.array_equal(D.agg_signal(rows[:1], [-1], lo, isz).cpu().numpy(), core[0][::-1])


def test_fused_pass_2000_intervals_vs_oracle(big):
    """>= 2000 random intervals plus the edge intervals of a chromosome-scale shard, the fused
    pass (WPS + coverage + length histogram) against the brute-force oracle; totals against numpy."""
    import os
    import torch
    from finaletoolkit_b200 import device as D
    fr, ofr, st, sp, mq = big["fr"], big["ofr"], big["st"], big["sp"], big["mq"]
    edges = np.arange(0, CLEN + 5000, 5000).clip(max=CLEN)
    plan = D.WpsPlan(edges[:-1], edges[1:], CLEN, 180, big["dev"])
    wps, cov, hist = plan.run_fused(fr, 120, 120, 180, 30, None, None, 30, n_bins=fr.max_len + 1)
    torch.cuda.synchronize()
    rng = np.random.default_rng(11)
    pick = np.unique(np.concatenate([[0, 1, plan.n_intervals - 2, plan.n_intervals - 1],
                                     rng.choice(plan.n_intervals, 2100, replace=False)]))
    assert len(pick) >= 2000
    s_, e_ = edges[:-1][pick], edges[1:][pick]
    exp, off = O.wps_intervals(ofr, s_, e_, CLEN, 120, 120, 180, 30, threads=os.cpu_count() or 1)
    sel = torch.from_numpy(np.concatenate([np.arange(plan.offsets[i], plan.offsets[i + 1]) for i in pick])).to(big["dev"])
    assert np.array_equal(wps[sel].cpu().numpy().astype(np.int64), exp)
    exp_cov = O.interval_coverage(ofr, s_, e_, None, None, "midpoint", 30, threads=os.cpu_count() or 1)
    assert np.array_equal(cov.cpu().numpy()[pick], exp_cov)
    L = (sp - st).astype(np.int64); mid = st.astype(np.int64) + L // 2
    ok = (mq >= 30) & (mid < CLEN)
    assert np.array_equal(hist.cpu().numpy(), np.bincount(L[ok], minlength=fr.max_len + 1))
    assert int(cov.sum()) == int(ok.sum())


def test_narrow_wire_records_and_streamed_pipeline_at_scale(big):
    """20 M fragments through the 24-bit wire records: lossless round trip (column checksums + equality), and the
    streamed pipeline (packed in, unpack + fused sweep, int8 out, captured CUDA graph) == the resident int32 pass."""
    import torch
    from finaletoolkit_b200 import device as D
    from finaletoolkit_b200.packed import PackedFragments
    from finaletoolkit_b200.pipeline import StreamedContig
    dev, st, sp, mq = big["dev"], big["st"], big["sp"], big["mq"]
    pk = PackedFragments(st, sp, mq, None)
    assert pk.record_bytes == 3 and pk.n_raw < pk.n_blocks // 50 and pk.wire_bytes() < 3.07 * NFRAG
    fr = pk.to_device(dev)
    assert torch.equal(fr.start, big["fr"].start) and torch.equal(fr.stop, big["fr"].stop) and torch.equal(fr.mapq, big["fr"].mapq)
    edges = np.arange(0, CLEN + 5000, 5000).clip(max=CLEN)
    plan = D.WpsPlan(edges[:-1], edges[1:], CLEN, 180, dev)
    ref, cov, hist = plan.run_fused(big["fr"], n_bins=pk.max_len + 1)
    pipe = StreamedContig(None, None, None, edges[:-1], edges[1:], CLEN, n_chunks=8, device=dev, packed=pk, wps_dtype="int8")
    for rep in range(3):          # eager, then the captured graph twice
        pipe.h_wps.zero_()
        w, c, h, t = pipe.run()
    assert pipe._graph is not None
    assert torch.equal(w[: pipe.n_positions].to(torch.int32), ref.cpu())
    assert torch.equal(c, cov.cpu()) and torch.equal(h[0], hist.cpu()) and int(t[0]) == int(cov.sum())


def test_shard_at_chromosome_scale(big):
    """Three contigs (the 60-Mb one twice under different names + a short one) through ONE launch of the shard ==
    the per-contig fused pass, incl. the adjusted series; checksum of checksums across the two copies."""
    import torch
    from finaletoolkit_b200 import device as D
    from finaletoolkit_b200 import distributed as FD
    from finaletoolkit_b200.io.fragments import FragmentTable
    dev, st, sp, mq, sd = big["dev"], big["st"], big["sp"], big["mq"], big["sd"]
    k = 2_000_000
    cols = {"A": (st, sp, mq, sd), "S": (st[:k], sp[:k], mq[:k], sd[:k]), "B": (st, sp, mq, sd)}
    sizes = {"A": CLEN, "S": int(sp[:k].max()) + 1000, "B": CLEN}
    table = FragmentTable(cols)
    res, hist, tot = FD.multi_wps_genome(table, list(sizes.items()), None, 5000, 120, 120, 180, 30, coverage=True,
                                         length_hist=True, adjust=dict(median_window_size=1000), device=dev,
                                         contigs=list(sizes))
    assert torch.equal(res["A"].wps, res["B"].wps) and torch.equal(res["A"].cov, res["B"].cov)
    assert torch.equal(res["A"].adjusted, res["B"].adjusted)
    edges = np.arange(0, CLEN + 5000, 5000).clip(max=CLEN)
    plan = D.WpsPlan(edges[:-1], edges[1:], CLEN, 180, dev)
    wps, cov, h = plan.run_fused(big["fr"], 120, 120, 180, 30, None, None, 30, n_bins=hist.numel())
    assert torch.equal(res["A"].wps, wps) and torch.equal(res["A"].cov, cov)
    ap = D.AdjustPlan(np.diff(plan.offsets), 1000, True, 21, 2, dev, skip_short=True)
    adj, _ = D.adjust_segments(wps, None, plan=ap, median_window_size=1000)
    assert torch.equal(res["A"].adjusted, adj)
    assert tot == 2 * int(cov.sum()) + int(res["S"].cov.sum())
    assert int(hist.sum()) == tot
