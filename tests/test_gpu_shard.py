"""``distributed.GenomeShard``: the contigs of a rank laid end to end in one virtual coordinate space
(one range prepass + one persistent launch per pass) must give, contig by contig, exactly what the
per-contig plans give - and what the oracle's single stream gives (reference drivers
frag/_multi_wps.py:152-198, frag/_adjust_wps.py:229-291)."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from finaletoolkit_b200.device import require_cuda
    return require_cuda("cuda:0")


def _table(rng, contigs):
    from finaletoolkit_b200.io.fragments import FragmentTable
    from finaletoolkit_b200.synth import synth_fragments
    cols = {}
    for i, (c, n) in enumerate(contigs.items()):
        k = 0 if c == "empty" else int(rng.integers(n // 8, n // 3))
        st, sp, mq, sd = synth_fragments(n, k, i, seed_base=4100)
        if k > 50:
            # rows that reach past the contig end, zero-length rows, long outliers, no mapq cut on some
            sp[-3:] = n + np.array([5, 400, 2000], np.int32)
            z = rng.choice(k, 5, replace=False); sp[z] = st[z]
            j = rng.choice(k, 5, replace=False); sp[j] = st[j] + rng.integers(700, 2500, 5).astype(np.int32)
        cols[c] = (st, sp, mq, sd)
    return FragmentTable(cols), cols


def _sites(rng, contigs, gappy):
    sites = {}
    for c, n in contigs.items():
        if gappy:
            s = np.sort(rng.choice(max(n - 1, 1), size=min(12, max(n // 3000, 1)), replace=False)).astype(np.int64)
            e = np.minimum(s + rng.integers(1, 9000, len(s)), n + 30)
            e[-1] = s[-1]                                # a degenerate interval: no output
        else:
            edges = np.arange(0, n + 5000, 5000, dtype=np.int64).clip(max=n)
            s, e = edges[:-1].copy(), edges[1:].copy()
        sites[c] = (s, e)
    return sites


@pytest.mark.parametrize("gappy", [False, True])
def test_shard_equals_per_contig_and_oracle(gappy, dev):
    import torch
    from finaletoolkit_b200 import distributed as FD
    from finaletoolkit_b200.device import AdjustPlan, WpsPlan, adjust_segments
    rng = np.random.default_rng(11 + gappy)
    contigs = {"a": 61_003, "empty": 20_001, "b": 33_333, "c": 7_001, "d": 90_002}
    table, cols = _table(rng, contigs)
    sites = _sites(rng, contigs, gappy)
    plans = {}
    adj = dict(median_window_size=1000, savgol=True, savgol_window_size=21, savgol_poly_deg=2)
    for rep in range(2):      # second call: cached shard, reused buffers
        res, hist, tot = FD.multi_wps_genome(table, list(contigs.items()), sites, 5000, 120, 120, 180, 30, coverage=True,
                                             length_hist=True, adjust=adj, device=dev, contigs=list(contigs), plans=plans)
    assert len(plans) == 1 and len(next(iter(plans.values())).groups) == 1
    exp_hist = np.zeros(hist.numel(), np.int64)
    for c, n in contigs.items():
        fr = table.device(c, dev)
        s, e = sites[c]
        plan = WpsPlan(s, e, n, 180, dev)
        wps, cov, h = plan.run_fused(fr, 120, 120, 180, 30, None, None, 30, n_bins=hist.numel())
        r = res[c]
        assert torch.equal(r.wps, wps[: plan.n_positions]) and torch.equal(r.cov, cov)
        exp_hist += h.cpu().numpy()
        ap = AdjustPlan(np.diff(plan.offsets), 1000, True, 21, 2, dev, skip_short=True)
        assert np.array_equal(r.adj_offsets, ap.out_off)
        if ap.n_total:
            a, _ = adjust_segments(wps, None, plan=ap, **adj)
            assert torch.equal(r.adjusted, a)
        ofr = O.Frags(*cols[c])
        exp, _ = O.wps_intervals(ofr, s, e, n, 120, 120, 180, 30, threads=4)
        assert np.array_equal(r.wps.cpu().numpy().astype(np.int64), exp)
        assert np.array_equal(r.cov.cpu().numpy(), O.interval_coverage(ofr, s, e, None, None, "midpoint", 30))
    assert np.array_equal(hist.cpu().numpy(), exp_hist) and tot == sum(int(r.cov.sum()) for r in res.values())

    # plain WPS (no fused by-products) through the same shard
    res2, h2, t2 = FD.multi_wps_genome(table, list(contigs.items()), sites, 5000, 120, 120, 180, 30, device=dev,
                                       contigs=list(contigs), plans=plans)
    assert h2 is None and t2 is None
    for c in contigs:
        assert torch.equal(res2[c].wps, res[c].wps) and res2[c].cov is None


def test_shard_splits_into_groups_before_int32_runs_out(dev, monkeypatch):
    import torch
    from finaletoolkit_b200 import distributed as FD
    rng = np.random.default_rng(5)
    contigs = {"a": 40_000, "b": 50_001, "c": 30_003}
    table, cols = _table(rng, contigs)
    sites = _sites(rng, contigs, False)
    one = FD.multi_wps_genome(table, list(contigs.items()), sites, 5000, 120, 120, 180, 30, coverage=True,
                              length_hist=True, device=dev, contigs=list(contigs))
    monkeypatch.setattr(FD, "_SHARD_SPAN", 120_000)     # forces a <= 2 contigs per group
    plans = {}
    two = FD.multi_wps_genome(table, list(contigs.items()), sites, 5000, 120, 120, 180, 30, coverage=True,
                              length_hist=True, device=dev, contigs=list(contigs), plans=plans)
    assert len(next(iter(plans.values())).groups) >= 2
    assert torch.equal(one[1], two[1]) and one[2] == two[2]
    for c in contigs:
        assert torch.equal(one[0][c].wps, two[0][c].wps) and torch.equal(one[0][c].cov, two[0][c].cov)


def test_genome_bin_counts_and_length_dict_through_the_shard(dev):
    """Genome-wide bin coverage + the first-seen-ordered length dict (frag/_frag_length.py:408-421) from one
    launch over the contigs laid end to end == the oracle's per-contig stream, zero-length rows included."""
    from finaletoolkit_b200 import distributed as FD
    rng = np.random.default_rng(21)
    contigs = {"a": 61_003, "empty": 20_001, "b": 33_333, "c": 7_001}
    table, cols = _table(rng, contigs)
    bins = {c: (np.arange(0, n, 10_000, dtype=np.int64), np.minimum(np.arange(0, n, 10_000, dtype=np.int64) + 10_000, n))
            for c, n in contigs.items()}
    n_bins = 2600
    got, ldict = FD.genome_bin_counts(table, bins, n_bins=n_bins, quality_threshold=30, device=dev)
    frs = {c: O.Frags(*cols[c]) for c in contigs}
    exp_dict = O.merge_dists(O.length_dist(frs[c], int(a), int(b), None, None, "midpoint", 30)
                             for c in contigs for a, b in zip(*bins[c]))
    for c in contigs:
        assert np.array_equal(got[c], O.interval_coverage(frs[c], bins[c][0], bins[c][1], None, None, "midpoint", 30)), c
    assert {k: v for k, v in exp_dict.items() if k < n_bins} == ldict
    # whole contigs (region None, None), the frag_length_bins stream: dict AND its first-seen order
    d = FD.genome_length_distribution(table, min_length=0, max_length=None, quality_threshold=30, device=dev)
    exp = O.merge_dists(O.length_dist(fr, None, None, 0, None, "midpoint", 30) for fr in frs.values())
    assert d == exp and list(d) == list(exp)
