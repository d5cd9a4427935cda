"""The BAM read-level replay (shared by tests/test_gpu_bam.py on the GPU and, with the device layer swapped
for the oracle by tests/host_shim.py, by tests/test_bam_api_dryrun.py on the CPU): every per-region feature of the
public API against ``tests/golden/bam_read1.*`` (outputs of the unmodified reference, oracle/make_golden_bam.py).
Integer results are bit-exact; the float statistics of ``frag_length_intervals`` are compared at 1e-12 relative
(the reference sums in BAM file order, the table is sorted by fragment start: another order of the same terms)."""
import hashlib
import io
import json
import os
from contextlib import redirect_stderr, redirect_stdout

import numpy as np
import pytest

from helpers import golden_codes, read_gz, write_2bit

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def make_fixture(d, golden):
    g = golden("bam_read1")
    with open(os.path.join(GOLDEN, "bam_read1.json")) as fh:
        m = json.load(fh)
    path = str(d / "read1.bam")
    with open(path, "wb") as fh:
        fh.write(g["bam_file"].tobytes())
    open(path + ".bai", "wb").close()
    tb = write_2bit(d / "read1.2bit", [(c, *golden_codes(g, c, n)) for c, n in m["refs"]])
    cs = d / "cs"; cs.write_text(m["chrom_sizes"])
    tiles = d / "tiles.bed"; tiles.write_text(m["tiles"])
    return dict(path=path, g=g, m=m, tb=tb, cs=str(cs), tiles=str(tiles), dir=d, sizes=dict(m["refs"]))


def check_table_carries_read1(bam):
    from finaletoolkit_b200.io.fragments import load_fragments
    tab = load_fragments(bam["path"])
    assert tab.has_read1("chrA") and tab.has_read1("chrB")
    tiles = [ln.split("\t") for ln in bam["m"]["tiles"].splitlines() if ln.startswith("chrA")]
    aff = tab.read1_affected("chrA", [int(x[1]) for x in tiles], [int(x[2]) for x in tiles])
    assert aff.mean() > 0.5          # the file is built so that the read-level rule matters almost everywhere


def check_wps_and_multi_wps(bam):
    import finaletoolkit_b200 as F
    g, m = bam["g"], bam["m"]
    for c in m["wps"]:
        r = F.wps(bam["path"], c["contig"], c["start"], c["stop"], bam["sizes"][c["contig"]], **c["kwargs"])
        assert np.array_equal(r["wps"].astype(np.int64), g[c["key"]]), c
    sites = bam["dir"] / "sites.bed"; sites.write_text(m["multi_wps"]["sites"])
    out = str(bam["dir"] / "mw.bed.gz")
    with redirect_stderr(io.StringIO()):
        F.multi_wps(bam["path"], str(sites), chrom_sizes=bam["cs"], output_file=out, **m["multi_wps"]["kwargs"])
    text = read_gz(out)
    got = np.array([int(ln.split("\t")[3]) for ln in text.splitlines()], np.int64)
    assert len(got) == m["multi_wps"]["n_lines"] and np.array_equal(got, g["multi_wps_scores"])
    assert hashlib.sha256(text.encode()).hexdigest() == m["multi_wps"]["sha256"]


def check_coverage(bam):
    import finaletoolkit_b200 as F
    m = bam["m"]
    for c in m["single_coverage"]:
        got = F.single_coverage(bam["path"], c["contig"], c["start"], c["stop"], **c["kwargs"])
        assert list(got) == c["result"], c
    for j, c in enumerate(m["coverage"]):
        out = str(bam["dir"] / f"cov_{j}.bed")
        F.coverage(bam["path"], bam["tiles"], out, **c["kwargs"])
        assert open(out).read() == c["text"], c["kwargs"]


def check_fragment_lengths(bam):
    import finaletoolkit_b200 as F
    g, m = bam["g"], bam["m"]
    for c in m["frag_length"]:
        got = F.frag_length(bam["path"], c["contig"], c["start"], c["stop"], **c["kwargs"])
        # same multiset; a BAM streams in read-position order, the table in fragment-start order
        assert np.array_equal(np.sort(got.astype(np.int64)), np.sort(g[c["key"]])), c
    for c in m["frag_length_bins"]:
        bins, counts = F.frag_length_bins(bam["path"], c["contig"], c["start"], c["stop"], **c["kwargs"])
        assert np.asarray(bins).tolist() == c["bins"] and np.asarray(counts).tolist() == c["counts"], c
    for c in m["frag_length_intervals"]:
        rows = F.frag_length_intervals(bam["path"], bam["tiles"], **c["kwargs"])
        assert len(rows) == len(c["rows"])
        for got, exp in zip(rows, c["rows"]):
            assert list(got[:4]) == exp[:4]
            assert [got[5], got[7], got[8], got[9]] == [exp[5], exp[7], exp[8], exp[9]], exp      # median, min, max, count
            assert got[4] == pytest.approx(exp[4], rel=1e-12) and got[6] == pytest.approx(exp[6], rel=1e-12, abs=1e-12)
            assert got[10] == pytest.approx(exp[10], rel=1e-12)


def check_motifs(bam):
    import finaletoolkit_b200 as F
    g, m = bam["g"], bam["m"]
    for c in m["region_motifs"]:
        d = getattr(F, c["fn"])(bam["path"], c["contig"], c["start"], c["stop"], bam["tb"], **c["kwargs"])
        assert np.array_equal(np.array(list(d.values()), np.int64), g[c["key"]]), c
    for c in m["interval_motifs"]:
        ivs = [tuple(iv) for iv in m[c.get("intervals", "motif_intervals")]]
        res = getattr(F, c["fn"])(bam["path"], bam["tb"], ivs, **c["kwargs"])
        got = np.array([list(d.values()) for _, d in res.intervals], np.int64)
        assert np.array_equal(got, g[c["key"]]), c


def check_cleavage(bam):
    import finaletoolkit_b200 as F
    g, m = bam["g"], bam["m"]
    for c in m["cleavage_profile"]:
        r = F.cleavage_profile(bam["path"], bam["sizes"][c["contig"]], c["contig"], c["start"], c["stop"], **c["kwargs"])
        assert np.array_equal(r["pos"], g[c["key"] + "_pos"])
        assert np.array_equal(r["proportion"], g[c["key"]]), c          # the same two integers divided: bit-exact
    bed = bam["dir"] / "clv.bed"; bed.write_text(m["multi_cleavage_profile"]["bed"])
    out = str(bam["dir"] / "clv.bed.gz")
    with redirect_stderr(io.StringIO()), redirect_stdout(io.StringIO()):
        F.multi_cleavage_profile(bam["path"], str(bed), bam["cs"], output_file=out, **m["multi_cleavage_profile"]["kwargs"])
    text = read_gz(out)
    assert len(text.splitlines()) == m["multi_cleavage_profile"]["n_lines"]
    assert hashlib.sha256(text.encode()).hexdigest() == m["multi_cleavage_profile"]["sha256"]


def check_delfi_bins(bam):
    from finaletoolkit_b200.frag._delfi import delfi_rows
    from finaletoolkit_b200.io.fragments import load_fragments
    from finaletoolkit_b200.io.reference import open_reference
    g, m = bam["g"], bam["m"]
    tab, ref = load_fragments(bam["path"]), open_reference(bam["tb"])
    bins = m["delfi"]["bins"]
    for c in m["delfi"]["cases"]:
        got = np.zeros((len(bins), 4), np.int64)
        for contig in ("chrA", "chrB"):
            idx = [i for i, b in enumerate(bins) if b[0] == contig]
            got[idx] = delfi_rows(tab, ref, contig, [bins[i][1] for i in idx], [bins[i][2] for i in idx],
                                  quality_threshold=c["quality_threshold"])
        assert np.array_equal(got[:, :3], g[c["key"]]), c
        gc = np.where(got[:, 2] > 0, got[:, 3] / np.array([b[2] - b[1] for b in bins]), np.nan)
        assert np.array_equal(gc, g[c["key"] + "_gc"], equal_nan=True)
