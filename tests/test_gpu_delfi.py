"""DELFI bin counts + GC content: CUDA path vs the reference's outputs / the oracle.

Reference: frag/_delfi.py:404-511 (_delfi_single_window) and :129-370 (delfi table, no LOESS).
"""
import numpy as np
import pytest

from helpers import delfi_tracks, golden_codes, write_2bit, write_frag_gz
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    import torch
    return torch.device("cuda:0")


def _dev_frags(g, prefix, dev):
    from finaletoolkit_b200.device import ContigFragments
    return ContigFragments(g[prefix + "start"], g[prefix + "stop"], g[prefix + "mapq"], g[prefix + "strand"], device=dev)


def test_delfi_windows_golden(manifest, golden, dev):
    """ftk_delfi_windows_u64 == the tuples of the reference's _delfi_single_window, bin by bin."""
    from finaletoolkit_b200 import device as D
    g = golden("delfi"); m = manifest["delfi"]
    sizes = dict(m["contigs"])
    frs = {c: _dev_frags(g, c + "_", dev) for c in sizes}
    refs = {c: D.PackedContig.from_codes(*golden_codes(g, c, n), device=dev) for c, n in sizes.items()}
    bl, gaps = delfi_tracks(m)
    bins = m["bins_list"]
    for case in m["single_window"]:
        k = case["key"]
        for c in sizes:
            sel = [j for j, b in enumerate(bins) if b[0] == c and case["arms"][j] != "NOARM"]
            if not sel:
                continue
            gp = gaps.get(c) if case["use_gaps"] else None
            got = D.delfi_windows(frs[c], refs[c], [bins[j][1] for j in sel], [bins[j][2] for j in sel],
                                  blacklist=bl.get(c) if case["use_blacklist"] else None,
                                  gaps=None if gp is None else gp[:2], quality_threshold=case["quality_threshold"]).cpu().numpy()
            assert np.array_equal(got[:, 0], g[k + "_short"][sel]) and np.array_equal(got[:, 1], g[k + "_long"][sel])
            assert np.array_equal(got[:, 2], g[k + "_num"][sel])
            width = np.array([bins[j][2] - bins[j][1] for j in sel], np.float64)
            gc = np.where(got[:, 2] > 0, got[:, 3] / width, np.nan)
            exp = g[k + "_gc"][sel]
            assert np.array_equal(np.isnan(gc), np.isnan(exp)) and np.array_equal(gc[~np.isnan(gc)], exp[~np.isnan(exp)])


@pytest.mark.parametrize("seed", range(3))
def test_delfi_windows_random_vs_oracle(seed, dev):
    from finaletoolkit_b200 import device as D
    from finaletoolkit_b200.device import ContigFragments
    from finaletoolkit_b200.synth import synth_fragments, synth_twobit
    rng = np.random.default_rng(500 + seed)
    clen = int(rng.integers(200_000, 900_000)); n = int(rng.integers(1000, 300_000))
    st, sp, mq, sd = synth_fragments(clen, n, seed, seed_base=4242)
    codes, nm = synth_twobit(clen, seed, seed_base=4343, telomere=700, block_len=2500)
    seq = np.frombuffer(b"ACGT", np.uint8)[codes].copy(); seq[nm] = ord("N")
    ofr = O.Frags(st, sp, mq, sd); dfr = ContigFragments(st, sp, mq, sd, device=dev)
    ref = D.PackedContig.from_codes(codes, nm, device=dev)
    width = [100_000, 5_000, 777][seed]
    ws = list(range(0, clen, width)); we = [a + width for a in ws]          # last bin passes the contig end
    ws += [3, clen - 50, 17]; we += [4, clen, clen]                          # tiny, tail and whole-contig bins
    r0 = np.sort(rng.integers(0, clen, 60)); r1 = r0 + rng.integers(1, 3 * width, 60)
    order = np.lexsort((r1, r0)); blk = (r0[order], r1[order])
    cen = (clen // 2, clen // 2 + 4000)
    telos = [[], [(0, 900)], [(0, clen // 3), (clen // 4, clen)]][seed]      # seed 2: the `all` rule bites
    q = [30, 0, 45][seed]
    for gaps, bl in [(None, None), ((cen, telos), blk)]:
        got = D.delfi_windows(dfr, ref, ws, we, blacklist=bl, gaps=gaps, quality_threshold=q).cpu().numpy()
        for j, (a, b) in enumerate(zip(ws, we)):
            out = O.delfi_counts(ofr, seq.tobytes(), a, b, bl, gaps, q)
            assert np.array_equal(got[j], out), (seed, a, b, got[j], out)


def test_delfi_api(tmp_path, manifest, golden):
    """finaletoolkit_b200.delfi(...) writes the same table as the reference's delfi (no LOESS)."""
    import finaletoolkit_b200 as F
    g = golden("delfi"); m = manifest["delfi"]
    sizes = dict(m["contigs"])
    cols = {c: tuple(g[f"{c}_{k}"] for k in ("start", "stop", "mapq", "strand")) for c in sizes}
    frag = write_frag_gz(tmp_path / "d.frag.gz", cols)
    tb = write_2bit(tmp_path / "d.2bit", [(c, *golden_codes(g, c, n)) for c, n in sizes.items()])
    paths = {}
    for name, key in [("delfi.chrom.sizes", "chrom_sizes"), ("delfi.gaps.bed", "gaps"), ("delfi.blacklist.bed", "blacklist"),
                      ("delfi.bins.bed", "bins")]:
        paths[name] = str(tmp_path / name)
        open(paths[name], "w").write(m[key])
    for j, t in enumerate(m["delfi"]):
        kw = {k: (paths[v] if isinstance(v, str) else v) for k, v in t["kwargs"].items()}
        out = str(tmp_path / f"o{j}.tsv")
        df = F.delfi(frag, paths["delfi.chrom.sizes"], paths["delfi.bins.bed"], tb, output_file=out, no_gc_correct=True, **kw)
        assert list(df.columns) == t["columns"] and df.shape[0] == t["n_rows"]
        assert [str(x) for x in df.dtypes] == t["dtypes"]
        assert open(out).read() == t["tsv"]
        outc = str(tmp_path / f"o{j}.csv")
        F.frag._delfi._write_delfi(df, outc)
        assert open(outc).read() == t["csv"]
    # a bin on a contig the fragment file does not have: pysam's ValueError
    open(paths["delfi.bins.bed"], "a").write("chrNoBins\t0\t5000\n")
    with pytest.raises(ValueError):
        F.delfi(frag, paths["delfi.chrom.sizes"], paths["delfi.bins.bed"], tb, no_gc_correct=True, merge_bins=False)
    with pytest.raises(TypeError):
        F.delfi(frag, paths["delfi.chrom.sizes"], paths["delfi.bins.bed"], tb, gap_file=5)
