"""CUDA cleavage profile vs the reference goldens and the oracle."""
import hashlib

import numpy as np
import pytest

from helpers import read_gz, write_frag_gz, write_text_gz
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def test_golden_cases_through_the_api(tmp_path, manifest, golden):
    import finaletoolkit_b200 as F
    from finaletoolkit_b200.io import bigwig
    g = golden("cleavage"); m = manifest["cleavage"]
    sy = golden("synth_small"); ms = manifest["synth_small"]
    fx = write_text_gz(tmp_path / "fx.frag.gz", manifest["fixture17"]["frag_gz_text"])
    cols = {c: tuple(sy[f"{c}_{k}"] for k in ("start", "stop", "mapq", "strand")) for c, _ in ms["contigs"]}
    syn = write_frag_gz(tmp_path / "syn.frag.gz", cols)
    for j, c in enumerate(m["cases"]):
        r = F.cleavage_profile(fx if c["src"] == "fixture" else syn, c["chrom_size"], c["contig"], c["start"], c["stop"], **c["kwargs"])
        assert r.dtype.names == ("contig", "pos", "proportion")
        assert np.array_equal(r["pos"], g[f"clv_{j}_pos"]) and np.array_equal(r["proportion"], g[f"clv_{j}_prop"]), c
    cs = tmp_path / "cs"; cs.write_text("".join(f"{c}\t{n}\n" for c, n in ms["contigs"]))
    bed = tmp_path / "clv.bed"; bed.write_text(m["bed"])
    for j, c in enumerate(m["multi"]):
        out = str(tmp_path / f"m{j}.bed.gz")
        with pytest.warns(UserWarning):
            assert F.multi_cleavage_profile(syn, str(bed), str(cs), output_file=out, **c["kwargs"]) == out
        txt = read_gz(out)
        assert len(txt.splitlines()) == c["n_lines"] and txt.splitlines()[:3] == c["head"]
        assert hashlib.sha256(txt.encode()).hexdigest() == c["sha256_text"]
        bw = str(tmp_path / f"m{j}.bw")
        with pytest.warns(UserWarning):
            F.multi_cleavage_profile(syn, str(bed), str(cs), output_file=bw, **c["kwargs"])
        r = bigwig.open(bw)
        vals = np.concatenate([r.intervals_arrays(cc, 0, dict(ms["contigs"])[cc])[2] for cc in c["bw_contigs"]])
        assert np.array_equal(vals, g[f"multi_{j}_bw_val_f32"])
    with pytest.raises(ValueError):
        F.multi_cleavage_profile(syn, str(bed), str(cs), output_file="x.txt")


@pytest.mark.parametrize("seed", range(3))
def test_random_vs_oracle(seed):
    from finaletoolkit_b200.device import ContigFragments, cleavage_intervals
    from finaletoolkit_b200.synth import synth_fragments
    rng = np.random.default_rng(50 + seed)
    clen = int(rng.integers(30_000, 90_000)); n = int(rng.integers(100, 40_000))
    st, sp, mq, sd = synth_fragments(clen, n, seed, seed_base=321)
    ofr = O.Frags(st, sp, mq, sd); dfr = ContigFragments(st, sp, mq, sd, device="cuda:0")
    ivs = [(0, clen), (0, 1), (clen - 1, clen), (100, 5219), (100, 5220), (7, 20_500)]
    ivs += [(int(s), int(min(s + rng.integers(1, 9000), clen))) for s in rng.integers(0, clen - 1, 8)]
    lo = [None, 100, 0][seed]; hi = [None, 200, 600][seed]; q = [30, 0, 60][seed]
    out, off = cleavage_intervals(dfr, [a for a, _ in ivs], [b for _, b in ivs], clen, lo, hi, q)
    out = out.cpu().numpy()
    for k, (a, b) in enumerate(ivs):
        _, exp = O.cleavage_profile(ofr, clen, a, b, 0, 0, lo, hi, q)
        assert np.array_equal(out[off[k]:off[k + 1]], exp), (seed, a, b)


def test_full_size_depth_identity():
    """Chromosome scale: sum over positions of ends == number of in-range ends (numpy), and the
    depth implied by proportion*depth... checked through tiling invariance."""
    import torch
    from finaletoolkit_b200.device import ContigFragments, cleavage_intervals
    from finaletoolkit_b200.synth import synth_fragments
    clen, n = 30_000_000, 9_000_000
    st, sp, mq, sd = synth_fragments(clen, n, 9)
    fr = ContigFragments(st, sp, mq, sd, device="cuda:0")
    a, _ = cleavage_intervals(fr, [0], [clen], clen, None, None, 30)
    edges = np.arange(0, clen + 5000, 5000).clip(max=clen)
    b, _ = cleavage_intervals(fr, edges[:-1], edges[1:], clen, None, None, 30)
    # depth/ends are local, so tiling changes nothing - except at the first position of an interval:
    # a '-' fragment whose stop equals the interval start does not overlap the interval (stop > start
    # fails), so its end is not counted there (reference semantics of frag_array(..., "any"))
    inner = torch.ones_like(a, dtype=torch.bool); inner[::5000] = False
    assert torch.equal(a[inner], b[inner]) and bool((b[~inner] <= a[~inner]).all())
    # numpy restatement on a window
    ok = mq >= 30
    lo, hi = 12_000_000, 12_200_000
    depth = np.zeros(hi - lo + 1, np.int64)
    s = np.clip(st[ok].astype(np.int64) - lo, 0, hi - lo); e = np.clip(sp[ok].astype(np.int64) - lo, 0, hi - lo)
    np.add.at(depth, s, 1); np.add.at(depth, e, -1)
    depth = np.cumsum(depth[:-1])
    endpos = np.where(sd[ok] == 1, st[ok], sp[ok]).astype(np.int64) - lo
    ends = np.bincount(endpos[(endpos >= 0) & (endpos < hi - lo)], minlength=hi - lo)
    exp = np.zeros(hi - lo); nz = depth != 0
    exp[nz] = ends[nz] / depth[nz] * 100
    assert np.array_equal(a[lo:hi].cpu().numpy(), exp)
