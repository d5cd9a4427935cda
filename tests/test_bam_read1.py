"""Read-level region selection of BAM input (io/alignment.py:242-247), CPU side.

``tests/golden/bam_read1.*`` hold what the UNMODIFIED reference yields on a synthetic short-read BAM
(oracle/make_golden_bam.py).  Here: (1) the oracle's restatement of the fetch is pinned on those outputs,
(2) the native decoder's read-1 spans equal the oracle's record walk, (3) the product's host-side selection
(``FragmentTable.read1_affected / fetch_groups / fetched_union``, ``frag/_common.per_fetch``) reproduces the
reference's per-interval numbers when the per-table computation is done by the oracle instead of a kernel,
(4) the host stream helpers (``utils.frag_generator`` / ``frag_array``) equal the reference's output.
The CUDA path replays the same vectors in ``tests/test_gpu_bam.py``.
"""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def bam(tmp_path_factory, golden):
    g = golden("bam_read1")
    with open(os.path.join(GOLDEN, "bam_read1.json")) as fh:
        m = json.load(fh)
    d = tmp_path_factory.mktemp("bam_read1")
    path = str(d / "read1.bam")
    with open(path, "wb") as fh:
        fh.write(g["bam_file"].tobytes())
    open(path + ".bai", "wb").close()
    return path, g, m


def _sorted_rows(a):
    a = np.asarray(a, np.int64).reshape(-1, 4)
    return a[np.lexsort((a[:, 3], a[:, 2], a[:, 1], a[:, 0]))]


def _frags(rows):
    a = np.array([[r[1], r[2], r[3], int(r[4])] for r in rows], np.int64).reshape(-1, 4)
    a = a[np.argsort(a[:, 0], kind="stable")]
    fr = O.Frags(a[:, 0].astype(np.int32), a[:, 1].astype(np.int32), a[:, 2].astype(np.uint8), a[:, 3].astype(np.uint8))
    return fr


def test_oracle_fetch_pinned_on_reference(bam):
    path, g, m = bam
    raw = g["bam_file"].tobytes()
    for c in m["frag_generator"]:
        rows = O.frag_stream(O.bam_fetch(raw, c["contig"], c["start"], c["stop"]), start=c["start"], stop=c["stop"],
                             **{"quality_threshold": 30, **c["kwargs"]})
        got = np.array([[r[1], r[2], r[3], int(r[4])] for r in rows], np.int64).reshape(-1, 4)
        assert np.array_equal(got, g[c["key"]]), c       # same rows, same (file) order
    assert sum(g[c["key"]].shape[0] > 0 for c in m["frag_generator"]) >= 9
    sizes = dict(m["refs"])
    for c in m["wps"]:
        kw = {"window_size": 120, "min_length": 120, "max_length": 180, "quality_threshold": 30, **c["kwargs"]}
        lo, hi = max(c["start"] - kw["max_length"], 0), min(c["stop"] + kw["max_length"], sizes[c["contig"]])
        fr = _frags(O.bam_fetch(raw, c["contig"], lo, hi))
        assert np.array_equal(O.wps_interval(fr, c["start"], c["stop"], sizes[c["contig"]], **kw), g[c["key"]]), c
        # the fragment-level selection alone is NOT what the reference computes on this file
    differs = 0
    for c in m["wps"]:
        kw = {"window_size": 120, "min_length": 120, "max_length": 180, "quality_threshold": 30, **c["kwargs"]}
        whole = _frags(O.bam_fetch(raw, c["contig"]))
        differs += not np.array_equal(O.wps_interval(whole, c["start"], c["stop"], sizes[c["contig"]], **kw), g[c["key"]])
    assert differs >= 1
    for c in m["single_coverage"]:
        if c["contig"] is None:
            continue
        fr = _frags(O.bam_fetch(raw, c["contig"], c["start"], c["stop"]))
        assert O.single_coverage(fr, c["start"], c["stop"], **c["kwargs"]) == c["result"][4], c


def test_decoder_read1_spans(bam):
    from finaletoolkit_b200.io import fragments
    path, g, m = bam
    fragments._CACHE.clear()
    tab = fragments.load_fragments(path)
    assert tab.is_sam and tab.has_read1() and tab.contigs == ["chrA", "chrB"]
    _, rows = O.bam_fragments(g["bam_file"].tobytes(), with_read1=True)
    for contig in tab.contigs:
        r = [x for x in rows if x[0] == contig]
        order = np.argsort(np.array([x[1] for x in r]), kind="stable")
        st, sp, mq, sd = tab.host(contig)
        assert st.tolist() == [r[i][1] for i in order] and sp.tolist() == [r[i][2] for i in order]
        lo, hi = tab.read1[contig]
        # stored clipped to the fragment (a dovetailed read decides nothing outside its own template)
        assert lo.tolist() == [max(r[i][5], r[i][1]) for i in order]
        assert hi.tolist() == [max(min(r[i][6], r[i][2]), max(r[i][5], r[i][1]) + 1) for i in order]
    assert any(x[6] > x[2] or x[5] < x[1] for x in rows if x[0] == "chrB")        # the dovetails are there
    fragments._CACHE.clear()


def test_host_stream_helpers_on_bam(bam):
    import finaletoolkit_b200 as F
    path, g, m = bam
    for c in m["frag_generator"]:
        rows = list(F.frag_generator(path, c["contig"], start=c["start"], stop=c["stop"], **c["kwargs"]))
        got = np.array([[r[1], r[2], r[3], int(r[4])] for r in rows], np.int64).reshape(-1, 4)
        # the table is sorted by fragment start, a BAM by read position: same rows, order within a contig may differ
        if c["contig"] is None:
            assert sorted({r[0] for r in rows}) == c["contigs"]
        assert np.array_equal(_sorted_rows(got), _sorted_rows(g[c["key"]])), c
    for c in m["frag_array"]:
        fa = F.frag_array(path, c["contig"], start=c["start"], stop=c["stop"], **c["kwargs"])
        got = np.stack([fa["start"], fa["stop"], fa["strand"].astype(np.int64)], axis=1)
        exp = g[c["key"]]
        assert np.array_equal(got[np.lexsort(got.T[::-1])], exp[np.lexsort(exp.T[::-1])])


def test_per_fetch_reproduces_reference_counts(bam):
    """``per_fetch`` with the per-table work done on the host: the grouping / union logic alone must turn a
    fragment-level counter into the reference's read-level numbers (coverage text of 110 intervals, 3 configs)."""
    from finaletoolkit_b200.frag._common import group_by_contig, per_fetch
    from finaletoolkit_b200.io import fragments
    path, g, m = bam
    fragments._CACHE.clear()
    tab = fragments.load_fragments(path)
    ivs = [ln.split("\t") for ln in m["tiles"].splitlines()]
    ivs = [(x[0], int(x[1]), int(x[2]), x[3] if len(x) > 3 else ".") for x in ivs]
    calls = []
    for case in m["coverage"]:
        kw = {"intersect_policy": "midpoint", "min_length": None, "max_length": None, "quality_threshold": 30,
              **{k: v for k, v in case["kwargs"].items() if k not in ("normalize", "scale_factor")}}
        counts = [0] * len(ivs)
        for contig, idx in group_by_contig([iv[0] for iv in ivs]).items():
            def run(t, sel, contig=contig, idx=idx):
                calls.append(len(sel))
                st, sp, mq, sd = t.host(contig)
                fr = O.Frags(st, sp, mq, sd)
                return [O.single_coverage(fr, ivs[idx[k]][1], ivs[idx[k]][2], **kw) for k in sel]
            for i, v in zip(idx, per_fetch(tab, contig, [ivs[i][1] for i in idx], [ivs[i][2] for i in idx], run)):
                counts[i] = v
        scale = case["kwargs"].get("scale_factor", 1.0)
        if case["kwargs"].get("normalize"):
            total = sum(O.single_coverage(O.Frags(*tab.host(c)), 0, None, **kw) for c in tab.contigs)
            scale /= total
        text = "".join(f"{c}\t{s}\t{e}\t{n}\t{v * scale}\n" for (c, s, e, n), v in zip(ivs, counts))
        assert text == case["text"], case["kwargs"]
    # a handful of batched calls per contig, not one per interval
    assert max(calls) > 10 and len(calls) < 40
    # and the fragment-level predicate alone gives different numbers on this file
    st, sp, mq, sd = tab.host("chrA")
    plain = [O.single_coverage(O.Frags(st, sp, mq, sd), s, e) for c, s, e, _ in ivs if c == "chrA"]
    exp = [float(ln.split("\t")[4]) for ln in m["coverage"][0]["text"].splitlines() if ln.startswith("chrA")]
    assert plain != exp
    fragments._CACHE.clear()


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_fetch_selection_brute_force(seed):
    """``read1_affected`` / ``fetched`` / ``fetch_groups`` / ``fetched_union`` against a brute-force restatement on
    random read-1 tables and random region sets (tilings, overlapping padded windows, scattered, unbounded)."""
    from finaletoolkit_b200.io.fragments import FragmentTable
    rng = np.random.default_rng(seed)
    n = 6000
    st = np.sort(rng.integers(0, 120_000, n))
    ln = rng.integers(40, 420, n)
    rl = np.minimum(rng.integers(20, 160, n), ln + rng.integers(-5, 30, n)).clip(1)      # some reads run past the template
    fwd = rng.random(n) < 0.5
    sp = st + ln
    r1s = np.where(fwd, st, sp - rl); r1e = np.where(fwd, st + rl, sp)
    tab = FragmentTable({"c": (st, sp, rng.integers(0, 61, n).astype(np.uint8), fwd.astype(np.uint8))}, is_sam=True,
                        read1={"c": (r1s, r1e)})
    fs, fe, _, _ = tab.host("c")
    lo1, hi1 = tab.read1["c"]
    assert np.all(lo1 >= fs) and np.all(hi1 <= fe) and np.all(hi1 > lo1)
    sets = [(np.arange(0, 120_000, 5_000), np.arange(0, 120_000, 5_000) + 5_000),
            (np.arange(0, 120_000, 900) - 180, np.arange(0, 120_000, 900) + 1_180),
            (rng.integers(-500, 120_000, 150), None)]
    for S, E in sets:
        if E is None:
            E = S + rng.integers(1, 6_000, len(S))
        aff = tab.read1_affected("c", S, E)
        groups = tab.fetch_groups("c", S, E)
        assert sorted(np.concatenate(groups).tolist()) == list(range(len(S)))
        for g in groups:
            a, b, _, _ = tab.fetched_union("c", S[g], E[g]).host("c")
            gs = np.sort(S[g]); ge = E[g][np.argsort(S[g], kind="stable")]
            assert np.all(gs[1:] >= ge[:-1])                                    # disjoint inside a group
            for k in g.tolist():
                fetched = (lo1 < E[k]) & (hi1 > S[k])                           # what an indexed fetch yields
                frag_any = (fs < E[k]) & (fe > S[k])
                assert aff[k] == bool((frag_any & ~fetched).any())
                one = tab.fetched("c", int(S[k]), int(E[k])).host("c")
                assert np.array_equal(one[0], fs[fetched]) and np.array_equal(one[1], fe[fetched])
                # inside the union table the kernels' own predicates see exactly the fetched rows of region k
                seen = (a < E[k]) & (b > S[k])
                assert np.array_equal(np.c_[a[seen], b[seen]], np.c_[fs[fetched & frag_any], fe[fetched & frag_any]])
                mid, cm = (a + b) // 2, (fs.astype(np.int64) + fe) // 2
                assert np.array_equal(a[(mid >= S[k]) & (mid < E[k])], fs[fetched & (cm >= S[k]) & (cm < E[k])])
        # fetch_only (no fragment-level test follows the fetch): selection by the WHOLE read, queried with bounds
        # widened by fetch_reach - every fetched row of region k passes the overlap test, no other row does
        raw_lo, raw_hi = tab.read1_raw["c"]
        reach = tab.fetch_reach("c")
        assert reach >= int((fe.astype(np.int64) - fs).max()) and reach >= int((raw_hi - raw_lo).max())
        aff = tab.read1_affected("c", S, E, fetch_only=True)
        groups = tab.fetch_groups("c", S, E, fetch_only=True)
        assert sorted(np.concatenate(groups).tolist()) == list(range(len(S)))
        for g in groups:
            a, b, _, _ = tab.fetched_union("c", S[g], E[g], fetch_only=True).host("c")
            for k in g.tolist():
                fetched = (raw_lo < E[k]) & (raw_hi > S[k])
                frag_any = (fs < E[k]) & (fe > S[k])
                assert aff[k] == bool((frag_any != fetched).any())
                one = tab.fetched("c", int(S[k]), int(E[k]), fetch_only=True).host("c")
                assert np.array_equal(one[0], fs[fetched]) and np.array_equal(one[1], fe[fetched])
                seen = (a < E[k] + reach) & (b > S[k] - reach)
                assert np.array_equal(np.c_[a[seen], b[seen]], np.c_[fs[fetched], fe[fetched]])
    assert tab.fetched("c", None, None).n_fragments("c") == n and not tab.read1_affected("c", [None], [None])[0]
    assert tab.fetched("c", 60_000, None).n_fragments("c") == int((hi1 > 60_000).sum())
    plain = FragmentTable({"c": (st, sp, np.zeros(n, np.uint8), np.ones(n, np.uint8))})
    assert not plain.has_read1("c") and plain.fetched("c", 5, 10) is plain and not plain.read1_affected("c", [5], [10]).any()


def test_alignment_wrapper(bam, tmp_path, manifest, golden):
    """``AlignmentWrapper.fetch`` (io/alignment.py:218-302) over the columnar table: BAM = the reads an indexed
    fetch returns (oracle restatement, pinned above), fragment file = the rows a tabix query returns."""
    import finaletoolkit_b200 as F
    from helpers import write_text_gz
    path, g, m = bam
    raw = g["bam_file"].tobytes()
    with F.AlignmentWrapper(path, quality_threshold=20) as w:
        assert w.is_sam and w.chroms == dict(m["refs"])
        for contig, s, e in [("chrA", 10_000, 12_000), ("chrB", 5_000, 9_000), ("chrB", None, 700), ("chrA", 59_000, None),
                             ("chrB", None, None), (None, None, None)]:
            got = sorted(tuple(f) for f in w.fetch(contig, s, e))
            exp = sorted(tuple(r) for r in O.bam_fetch(raw, contig, s, e) if r[3] >= 20)
            assert got == exp and (got or contig == "chrB"), (contig, s, e)
        f = next(w.fetch("chrA", 10_000, 12_000))
        assert isinstance(f, F.Fragment) and f.length == f.stop - f.start and isinstance(f.is_forward, bool)
    with pytest.raises(NotImplementedError):
        F.AlignmentWrapper(path, read1_only=False)
    fx = manifest["fixture17"]
    frag = write_text_gz(tmp_path / "a.frag.gz", fx["frag_gz_text"])
    rows = [ln.split("\t") for ln in fx["frag_gz_text"].splitlines()]
    with F.AlignmentWrapper(frag, quality_threshold=0) as w:
        assert not w.is_sam and w.chroms == {"12": None}
        got = [tuple(f) for f in w.fetch("12", 34443118, 34443538)]
        exp = [("12", int(r[1]), int(r[2]), int(r[3]), "+" in r[4]) for r in rows if int(r[2]) > 34443118 and int(r[1]) < 34443538]
        assert got == exp and len(got) > 0
        assert len(list(w.fetch())) == len(rows) == 17
