"""CUDA adjust_wps (median/mean + Savitzky-Golay) vs the reference goldens and the oracle.

Tolerance per BASELINE.json north_star: adjust_wps floats within 1e-5 relative
(absolute floor 1e-9 for values that are exactly 0 in the reference).
"""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 1e-9


def _run(x, lens, **kw):
    import torch
    from finaletoolkit_b200.device import adjust_segments
    out, off = adjust_segments(np.asarray(x, np.float32), lens, **kw)
    torch.cuda.synchronize()
    return out.cpu().numpy(), off


def test_core_golden(manifest, golden):
    g = golden("adjust"); m = manifest["adjust"]
    worst = 0.0
    for c in m["core_cases"]:
        x = g[c["input"]]
        sg = c["sg"]
        kw = dict(median_window_size=c["w"], use_mean=c["mean"], savgol=sg is not None)
        if sg:
            kw.update(savgol_window_size=sg[0], savgol_poly_deg=sg[1])
        got, _ = _run(x, [len(x)], **kw)
        exp = g[c["key"] + "_out"]
        np.testing.assert_allclose(got, exp, rtol=RTOL, atol=ATOL, err_msg=str(c))
        worst = max(worst, float(np.abs(got - exp).max()))
        if sg is None and not c["mean"]:
            assert np.array_equal(got, exp), c          # the pre-SG median path is exact
    assert worst < 1e-10


def test_driver_golden_and_run_splitting(manifest, golden):
    """Segments from the reference's adjust_wps driver (merged intervals), several run lengths."""
    g = golden("adjust"); m = manifest["adjust"]
    raw_pos, raw_val = g["raw_pos"], g["raw_val_f32"]
    lut = {int(p): i for i, p in enumerate(raw_pos.tolist())}
    for j, c in enumerate(m["adjust_cases"]):
        kw = {"interval_size": 5000, "median_window_size": 1000, "savgol_window_size": 21, "savgol_poly_deg": 2,
              "savgol": True, "mean": False, "subtract_edges": False, "edge_size": 500, **c["kwargs"]}
        sites = O.adjust_sites(c["bed"].splitlines(keepends=True), kw["interval_size"], kw["median_window_size"])
        segs, lens = [], []
        sizes = dict(manifest["synth_small"]["contigs"])
        for contig, s, e in sites:
            if e > sizes[contig]:
                continue  # the reference skips it (pyBigWig "Invalid interval bounds!")
            idx = [lut[p] for p in range(s, e) if p in lut]
            if idx:
                segs.append(raw_val[idx]); lens.append(len(idx))
        for run_len in (2048, 257, 64):
            got, off = _run(np.concatenate(segs), lens, median_window_size=kw["median_window_size"], use_mean=kw["mean"],
                            savgol=kw["savgol"], savgol_window_size=kw["savgol_window_size"],
                            savgol_poly_deg=kw["savgol_poly_deg"], subtract_edges=kw["subtract_edges"],
                            edge_size=kw["edge_size"], run_len=run_len)
            exp = g[f"adj_{j}_val_f32"]
            assert got.shape == exp.shape
            np.testing.assert_allclose(got, exp.astype(np.float64), rtol=2e-5, atol=1e-6, err_msg=f"{c} run={run_len}")
            # the reference stores float32 in the bigWig: after the same rounding almost all values are identical
            assert (got.astype(np.float32) == exp).mean() > 0.999


@pytest.mark.parametrize("seed", range(3))
def test_random_vs_oracle(seed):
    rng = np.random.default_rng(seed)
    lens = [int(v) for v in rng.integers(300, 9000, 6)]
    w = int(rng.choice([2, 100, 250, 300]))
    lens = [max(l, w + 40) for l in lens]
    kinds = []
    xs = []
    for i, l in enumerate(lens):
        if i % 3 == 0:      # integer WPS-like (fast path)
            v = np.cumsum(rng.integers(-2, 3, l)).astype(np.float32)
        elif i % 3 == 1:    # wide range: leaves the 256-bin window -> generic path
            v = (rng.integers(-5, 6, l) * rng.choice([1, 1, 1, 400], l)).astype(np.float32)
        else:               # non-integer floats -> generic path
            v = rng.normal(0, 3, l).astype(np.float32)
        xs.append(v)
    for use_mean in (False, True):
        for sub in (False, True):
            got, off = _run(np.concatenate(xs), lens, median_window_size=w, use_mean=use_mean, savgol=True,
                            savgol_window_size=21, savgol_poly_deg=2, subtract_edges=sub, edge_size=77, run_len=500)
            for i, v in enumerate(xs):
                x = v.astype(np.float64)
                if sub:
                    x = x - np.mean([np.mean(x[:77]), np.mean(x[-77:])])
                exp = O.adjust_core(x, w, use_mean, True, 21, 2)
                np.testing.assert_allclose(got[off[i]:off[i + 1]], exp, rtol=RTOL, atol=1e-8, err_msg=f"seg {i} mean={use_mean} sub={sub}")


def test_errors():
    from finaletoolkit_b200.device import adjust_segments
    x = np.zeros(500, np.float32)
    with pytest.raises(ValueError):
        adjust_segments(x, [500], median_window_size=1000)
    with pytest.raises(ValueError):
        adjust_segments(x, [500], median_window_size=101)
    with pytest.raises(ValueError):
        adjust_segments(x, [500], median_window_size=490)  # 10 outputs < savgol window


def _int_series(rng, l, kind):
    if kind == 0:        # WPS-like random walk
        return np.cumsum(rng.integers(-2, 3, l))
    if kind == 1:        # wide spread: the median wanders over > 32 levels -> several band passes
        return (np.cumsum(rng.integers(-9, 10, l)) + rng.integers(-40, 41, l))
    if kind == 2:        # constant + isolated spikes (upper median far above the lower one)
        v = np.full(l, 7); v[rng.choice(l, l // 50, replace=False)] = 3000
        return v
    if kind == 3:        # two plateaus 500 levels apart: the band has to jump
        v = np.where(np.arange(l) < l // 2, -250, 250) + rng.integers(-3, 4, l)
        return v
    return rng.integers(-15, 16, l) * 2   # only even values: gaps between occupied levels


@pytest.mark.parametrize("seed", range(6))
def test_rank_kernel_vs_oracle_and_hist_path(seed):
    """The fused rank-bitmap kernel (median + SG in one pass): vs the oracle (numpy median + scipy
    savgol) and bit-for-bit vs the sliding-histogram path; float32 and int32 inputs; multi-tile segments."""
    import torch
    from finaletoolkit_b200.device import adjust_segments
    rng = np.random.default_rng(50 + seed)
    w = int(rng.choice([2, 100, 250, 1000]))
    lens = [int(v) for v in rng.integers(w + 21, w + 9000, 5)] + [w + 21, w + 22, w + 4096, w + 4097, 13_000 + w]
    xs = [_int_series(rng, l, (seed + i) % 5) for i, l in enumerate(lens)]
    x = np.concatenate(xs)
    for savgol in (True, False):
        for sub in (False, True):
            kw = dict(median_window_size=w, savgol=savgol, savgol_window_size=21, savgol_poly_deg=2, subtract_edges=sub,
                      edge_size=33)
            got, off = adjust_segments(x.astype(np.float32), lens, **kw)
            got_i, _ = adjust_segments(torch.from_numpy(x.astype(np.int32)).cuda(), lens, **kw)
            old, _ = adjust_segments(x.astype(np.float32), lens, impl="hist", **kw)
            torch.cuda.synchronize()
            assert torch.equal(got, got_i)
            if savgol and not sub:
                # exact sliding-moment smoothing vs the fp64 stencil of the fallback path: they differ by the
                # stencil's own rounding only (values are O(100): 1e-11 absolute is ~1e-13 relative)
                assert float((got - old).abs().max()) < 1e-11, (seed, w, float((got - old).abs().max()))
            else:
                assert torch.equal(got, old), (seed, w, savgol, sub, float((got - old).abs().max()))
            gh = got.cpu().numpy()
            for i, v in enumerate(xs):
                xx = v.astype(np.float64)
                if sub:
                    xx = xx - np.mean([np.mean(xx[:33]), np.mean(xx[-33:])])
                exp = O.adjust_core(xx, w, False, savgol, 21, 2)
                np.testing.assert_allclose(gh[off[i]:off[i + 1]], exp, rtol=RTOL, atol=1e-8, err_msg=f"seg {i} w={w}")
                if not savgol and not sub:
                    assert np.array_equal(gh[off[i]:off[i + 1]], exp)     # pre-SG median series is exact


def test_rank_kernel_other_sg_windows_and_fallback():
    """Runtime Savitzky-Golay windows (5, 51) and the flagged-tile fallback (one non-integer sample)."""
    import torch
    from finaletoolkit_b200.device import adjust_segments
    rng = np.random.default_rng(3)
    lens = [3000, 5000, 1200]
    x = np.concatenate([_int_series(rng, l, 0) for l in lens]).astype(np.float32)
    from finaletoolkit_b200.device import savgol_rational
    for sgw, deg in ((5, 2), (51, 3), (3, 2), (7, 0), (9, 1), (11, 4), (21, 5), (101, 2)):
        got, off = adjust_segments(x, lens, median_window_size=500, savgol_window_size=sgw, savgol_poly_deg=deg)
        old, _ = adjust_segments(x, lens, median_window_size=500, savgol_window_size=sgw, savgol_poly_deg=deg, impl="hist")
        if deg >= 4:
            assert savgol_rational(sgw, deg) == (0, 0, 0) and torch.equal(got, old)      # fp64 stencil in both
        else:
            assert savgol_rational(sgw, deg)[2] > 0 and float((got - old).abs().max()) < 1e-11
        o = 0
        gh = got.cpu().numpy()
        for i, l in enumerate(lens):
            exp = O.adjust_core(x[o:o + l].astype(np.float64), 500, False, True, sgw, deg)
            np.testing.assert_allclose(gh[off[i]:off[i + 1]], exp, rtol=RTOL, atol=1e-8)
            o += l
    y = x.copy(); y[4000] += 0.5
    got, off = adjust_segments(y, lens, median_window_size=500)
    o = 0
    gh = got.cpu().numpy()
    for i, l in enumerate(lens):
        exp = O.adjust_core(y[o:o + l].astype(np.float64), 500, False, True, 21, 2)
        np.testing.assert_allclose(gh[off[i]:off[i + 1]], exp, rtol=RTOL, atol=1e-8)
        o += l
