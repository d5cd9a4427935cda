"""agg_bw: CUDA accumulation vs the reference's outputs / the oracle.

Reference: utils/_agg_bw.py:18-146; its own fixture tests/data/test.bw + known answers tests/test_agg_bw.py:15-26.
"""
import numpy as np
import pytest

from helpers import agg_fixture
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def test_agg_signal_kernel_vs_oracle():
    from finaletoolkit_b200 import device as D
    rng = np.random.default_rng(9)
    for n_seg, row_len, mws in [(1, 10, 1), (37, 501, 120), (1000, 5000, 1000), (9, 64, 0), (130, 300, 7)]:
        rows = rng.normal(0, 50, (n_seg, row_len)).astype(np.float32)
        rows[rng.random(rows.shape) < 0.01] = np.nan
        strands = rng.choice(["+", "-", "."], n_seg).tolist()
        isz = row_len - mws
        lo = mws // 2
        exp = O.agg_bw_core(list(rows), strands, mws)
        if mws == 0:          # values[0:-0] is empty: every interval is skipped by the reference
            assert exp.dtype == np.int64 and not exp.any()
            continue
        got = D.agg_signal(rows, [1 if s == "+" else -1 if s == "-" else 0 for s in strands], lo, isz).cpu().numpy()
        assert np.array_equal(got, exp.astype(np.float64)), (n_seg, row_len, mws)   # bit-exact fp64 running sum


def test_agg_bw_api_golden(tmp_path, manifest, golden, capsys):
    import finaletoolkit_b200 as F
    g = golden("agg"); m = manifest["agg"]
    _, _, write_bw = agg_fixture(g, m)
    bw = write_bw(tmp_path / "agg.bw")
    bed = str(tmp_path / "agg.bed"); open(bed, "w").write(m["bed"])
    for j, c in enumerate(m["cases"]):
        out = str(tmp_path / f"o{j}.wig")
        with np.errstate(invalid="ignore", divide="ignore"):
            r = F.agg_bw(bw, bed, out, **c["kwargs"])
        assert str(r.dtype) == c["dtype"] and np.array_equal(r, g[c["key"]], equal_nan=True), c["kwargs"]
        assert open(out).read() == c["wig"]
    assert "Invalid interval bounds!" in capsys.readouterr().out
    with pytest.raises(ValueError):
        F.agg_bw(bw, bed, str(tmp_path / "x.txt"))
    with pytest.raises(ValueError):
        F.agg_bw(bw, str(tmp_path / "agg.bw"), str(tmp_path / "x.wig"))
    # the reference's own bigWig fixture (written by libBigWig) and its known answers
    ref_bw = str(tmp_path / "test.bw"); open(ref_bw, "wb").write(g["ref_test_bw"].tobytes())
    ref_bed = str(tmp_path / "bw_test.bed"); open(ref_bed, "w").write(m["ref_bed"])
    for k in m["ref_known"]:
        r = F.agg_bw(ref_bw, ref_bed, str(tmp_path / "r.wig"), k["median_window_size"])
        assert r == pytest.approx(k["expect"])
