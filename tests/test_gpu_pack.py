"""Packed wire format: host pack -> device unpack is lossless (escapes, ragged tails, slices), and the
packed streamed pipeline equals the resident int32 path bit for bit."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from finaletoolkit_b200.device import require_cuda
    return require_cuda("cuda:0")


def _cols(rng, n, escapes=True):
    st = np.sort(rng.integers(0, max(30 * n, 1000), n)).astype(np.int32)
    ln = rng.integers(0, 700, n).astype(np.int32)
    if escapes and n > 300:
        ln[rng.choice(n, 5, replace=False)] = 4096            # one past the 12-bit field
        ln[rng.choice(n, 5, replace=False)] = 4095            # largest length that still fits
        ln[rng.choice(n, 3, replace=False)] = -7              # malformed row: stop < start
        st[n // 2:] += 2048                                    # a gap of >= 2048 bp mid-block
        st[3 * n // 4:] += 2047                                # largest gap that still fits (if alone)
        ln[rng.choice(n, 2, replace=False)] = 70_000
    sp = (st + ln).astype(np.int32)
    mq = rng.integers(0, 256, n).astype(np.uint8)
    sd = rng.integers(0, 2, n).astype(np.uint8)
    return st, sp, mq, sd


@pytest.mark.parametrize("n", [0, 1, 2, 63, 64, 65, 127, 128, 1000, 77_777, 1_000_001])
def test_pack_unpack_roundtrip(n, dev):
    import torch
    from finaletoolkit_b200.packed import PackedFragments
    rng = np.random.default_rng(n)
    st, sp, mq, sd = _cols(rng, n)
    pk = PackedFragments(st, sp, mq, sd)
    fr = pk.to_device(dev)
    torch.cuda.synchronize()
    assert fr.n == n
    assert np.array_equal(fr.start.cpu().numpy(), st) and np.array_equal(fr.stop.cpu().numpy(), sp)
    assert np.array_equal(fr.mapq.cpu().numpy(), mq) and np.array_equal(fr.strand.cpu().numpy(), sd)
    if n > 300:
        assert 0 < pk.n_raw < pk.n_blocks          # escapes stay local to their blocks
    # no mapq / strand columns: 255 / absent
    pk2 = PackedFragments(st, sp, None, None)
    fr2 = pk2.to_device(dev)
    assert fr2.mapq is None and fr2.strand is None and np.array_equal(fr2.stop.cpu().numpy(), sp)


@pytest.mark.parametrize("n", [1, 64, 65, 129, 5000, 300_007])
def test_pack_unpack_narrow_records(n, dev):
    """24-bit records (record_bytes = 3): lossless incl. blocks that escape by gap / length, ragged tails."""
    import torch
    from finaletoolkit_b200.packed import PackedFragments
    rng = np.random.default_rng(900 + n)
    st = (np.cumsum(rng.integers(0, 9, n)) + 3).astype(np.int32)
    ln = rng.integers(0, 512, n).astype(np.int32)
    if n >= 5000:
        ln[rng.choice(n, 4, replace=False)] = 512             # one past the 9-bit field
        ln[rng.choice(n, 4, replace=False)] = 511
        st[n // 3:] += 64                                      # one past the 6-bit gap
        st[2 * n // 3:] += 63
    sp = (st + ln).astype(np.int32)
    mq = rng.integers(0, 256, n).astype(np.uint8); sd = rng.integers(0, 2, n).astype(np.uint8)
    for rb in (3, None):
        pk = PackedFragments(st, sp, mq, sd, record_bytes=rb)
        assert pk.record_bytes == 3 or rb is None          # None: whichever puts fewer bytes on the wire
        fr = pk.to_device(dev)
        torch.cuda.synchronize()
        assert np.array_equal(fr.start.cpu().numpy(), st) and np.array_equal(fr.stop.cpu().numpy(), sp)
        assert np.array_equal(fr.mapq.cpu().numpy(), mq) and np.array_equal(fr.strand.cpu().numpy(), sd)
        if n >= 5000 and pk.record_bytes == 3:
            assert 0 < pk.n_raw <= 10


def test_pack_negative_start_and_all_raw(dev):
    from finaletoolkit_b200.packed import PackedFragments
    st = np.array([-5, -1, 0, 10, 10_000_000, 10_000_001], np.int32)
    sp = st + np.array([100, 5000, 3, 0, 150, 151], np.int32)
    pk = PackedFragments(st, sp, np.arange(6, dtype=np.uint8), np.array([1, 0, 1, 0, 1, 1], np.uint8))
    assert pk.n_raw == 1
    fr = pk.to_device(dev)
    assert fr.start.cpu().tolist() == st.tolist() and fr.stop.cpu().tolist() == sp.tolist()
    assert fr.mapq.cpu().tolist() == list(range(6)) and fr.strand.cpu().tolist() == [1, 0, 1, 0, 1, 1]


def test_wire_bytes_budget():
    """chr1 at 30x: <= 330 MB for 80 M fragments (4.0625 B per fragment when nothing escapes)."""
    from finaletoolkit_b200.packed import PACK_BLOCK
    n = 80_000_000
    nb = (n + PACK_BLOCK - 1) // PACK_BLOCK
    assert nb * PACK_BLOCK * 4 + nb * 4 <= 330_000_000 and nb * PACK_BLOCK * 3 + nb * 4 <= 246_000_000


@pytest.mark.parametrize("chunks", [1, 5])
def test_packed_pipeline_matches_resident(chunks, dev):
    import torch
    from finaletoolkit_b200 import device as D
    from finaletoolkit_b200.packed import PackedFragments
    from finaletoolkit_b200.pipeline import StreamedContig
    from finaletoolkit_b200.synth import synth_fragments
    clen, n = 3_000_000, 900_000
    st, sp, mq, sd = synth_fragments(clen, n, 11)
    sp[::5000] = st[::5000] + 5000                 # a few escaped blocks inside the stream
    edges = np.arange(0, clen + 5000, 5000).clip(max=clen)
    fr = D.ContigFragments(st, sp, mq, sd, device=dev)
    plan = D.WpsPlan(edges[:-1], edges[1:], clen, 180, dev)
    ref, cov, hist = plan.run_fused(fr, n_bins=fr.max_len + 1)
    pk = PackedFragments(st, sp, mq, None)
    assert pk.n_raw > 0
    pipe = StreamedContig(None, None, None, edges[:-1], edges[1:], clen, n_chunks=chunks, device=dev, packed=pk,
                          wps_dtype="int8")
    for rep in range(3):                 # eager pass, then the captured CUDA graph twice
        pipe.h_wps.zero_(); pipe.h_cov.zero_(); pipe.h_hist.zero_()
        w, c, h, t = pipe.run()
        assert (pipe._graph is not None) == (rep > 0)
    assert np.array_equal(w.numpy()[: pipe.n_positions].astype(np.int32), ref.cpu().numpy())
    assert np.array_equal(c.numpy(), cov.cpu().numpy()) and np.array_equal(h.numpy()[0], hist.cpu().numpy())
    assert int(t[0]) == int(cov.sum())
    assert pipe.h2d_bytes < 4.2 * n + 10 * 64 * pk.n_raw + 4.2 * 64 * 2 * chunks * 20


def test_pipeline_gapped_intervals_and_empty(dev):
    """multi_wps-like windows with gaps between them: coverage / histogram are those of the intervals
    (not of their hull), whatever the chunking; an empty interval set is a no-op."""
    import torch
    from finaletoolkit_b200.packed import PackedFragments
    from finaletoolkit_b200.pipeline import StreamedContig
    from finaletoolkit_b200.synth import synth_fragments
    clen, n = 400_000, 120_000
    st, sp, mq, sd = synth_fragments(clen, n, 3, seed_base=77)
    ofr = O.Frags(st, sp, mq, sd)
    rng = np.random.default_rng(0)
    s = np.sort(rng.integers(0, clen - 6000, 25)); e = s + rng.integers(1, 6000, 25)
    pk = PackedFragments(st, sp, mq, None)
    exp_cov = O.interval_coverage(ofr, s, e, None, None, "midpoint", 30)
    exp_hist = np.zeros(pk.max_len + 1, np.int64)
    for a, b in zip(s, e):
        for L, c in O.length_dist(ofr, int(a), int(b), None, None, "midpoint", 30).items():
            exp_hist[L] += c
    for chunks in (1, 4):
        pipe = StreamedContig(None, None, None, s, e, clen, n_chunks=chunks, device=dev, packed=pk)
        w, c, h, t = pipe.run()
        assert np.array_equal(c.numpy(), exp_cov) and np.array_equal(h.numpy()[0], exp_hist)
        host = w.numpy().astype(np.int64)
        for i in range(len(s)):
            assert np.array_equal(host[pipe.offsets[i]:pipe.offsets[i + 1]], O.wps_interval(ofr, int(s[i]), int(e[i]), clen))
    empty = StreamedContig(None, None, None, [], [], clen, device=dev, packed=pk)
    w, c, h, t = empty.run()
    assert empty.n_positions == 0 and int(t[0]) == 0 and not h.numpy().any()
