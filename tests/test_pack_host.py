"""CPU-only: the host packer of the wire format (``ftk_pack_fragments_host``) against a numpy decoder
of the documented layout (include/ftk_b200.h), including escaped (raw) blocks and ragged tails."""
import numpy as np
import pytest


def _decode(pk, n):
    """Pure-numpy restatement of unpack_fragments_kernel (csrc/ftk_pack.cu)."""
    B = 64
    anchors = pk.anchors.numpy()[: pk.n_blocks].astype(np.int64)
    if pk.record_bytes == 4:
        w = pk.words.numpy().view(np.uint32)[: pk.n_blocks * B].astype(np.int64)
        d = (w & 2047).reshape(-1, B); ln = (w >> 11) & 4095
        mapq = (w >> 24) & 255; strand = (w >> 23) & 1
    else:           # 24-bit records back to back, little-endian
        b8 = pk.words.numpy().view(np.uint8)[: pk.n_blocks * B * 3].astype(np.int64).reshape(-1, 3)
        w = b8[:, 0] | (b8[:, 1] << 8) | (b8[:, 2] << 16)
        d = (w & 63).reshape(-1, B); ln = (w >> 6) & 511
        mapq = (w >> 16) & 255; strand = (w >> 15) & 1
    start = (anchors[:, None] + np.cumsum(d, axis=1)).reshape(-1)
    stop = start + ln
    raw = np.flatnonzero(anchors < 0)
    for b in raw:
        r = int(-1 - anchors[b])
        sl, rs = slice(b * B, (b + 1) * B), slice(r * B, (r + 1) * B)
        start[sl] = pk.raw_start.numpy()[rs]; stop[sl] = pk.raw_stop.numpy()[rs]
        mapq[sl] = pk.raw_mapq.numpy()[rs]; strand[sl] = pk.raw_strand.numpy()[rs]
    return start[:n], stop[:n], mapq[:n], strand[:n], len(raw)


@pytest.mark.parametrize("n", [0, 1, 64, 65, 1000, 200_003])
def test_pack_narrow_records_roundtrip(n):
    """record_bytes = 3 (24-bit records): dense short-fragment data packs without escapes, a 64-bp gap or a
    512-bp fragment sends its block to the raw columns, and the automatic choice takes the smaller format."""
    from finaletoolkit_b200.packed import PackedFragments
    rng = np.random.default_rng(n + 7)
    st = np.cumsum(rng.integers(0, 8, n)).astype(np.int32) + 5
    ln = rng.integers(30, 512, n).astype(np.int32)
    if n >= 1000:
        ln[[10, 500]] = [512, 700]          # two escapes by length
        st[n // 2 + 100:] += 64             # one by gap (a delta of 63 still fits)
    sp = (st + ln).astype(np.int32)
    mq = rng.integers(0, 256, n).astype(np.uint8); sd = rng.integers(0, 2, n).astype(np.uint8)
    pk = PackedFragments(st, sp, mq, sd, pinned=False, threads=2, record_bytes=3)
    assert pk.record_bytes == 3 and pk.wire_bytes() == pk.n_blocks * 64 * 3 + pk.n_blocks * 4
    s2, e2, q2, d2, n_raw = _decode(pk, n)
    assert np.array_equal(s2, st) and np.array_equal(e2, sp) and np.array_equal(q2, mq) and np.array_equal(d2, sd)
    assert n_raw == pk.n_raw and (pk.n_raw == (3 if n >= 1000 else 0))
    auto = PackedFragments(st, sp, mq, sd, pinned=False)
    # dense data: the narrow records win - unless the escapes outweigh them (3 raw blocks of 16 at n = 1000)
    assert auto.record_bytes == (4 if n == 1000 else 3)
    if n >= 1000:
        sparse = PackedFragments((st.astype(np.int64) * 40).astype(np.int32), (st.astype(np.int64) * 40 + ln).astype(np.int32),
                                 mq, sd, pinned=False)
        assert sparse.record_bytes == 4           # gaps of ~140 bp: nearly every block would escape


@pytest.mark.parametrize("n", [0, 1, 64, 65, 1000, 200_003])
def test_pack_layout_roundtrip(n):
    from finaletoolkit_b200.packed import PackedFragments
    rng = np.random.default_rng(n + 1)
    st = np.sort(rng.integers(0, max(20 * n, 1000), n)).astype(np.int32)
    ln = rng.integers(0, 800, n).astype(np.int32)
    if n >= 1000:
        ln[[10, 500]] = [4096, -3]
        ln[700] = 4095
        st[n // 2:] += 2048
        st[n // 2 + 300:] += 2047
    sp = (st + ln).astype(np.int32)
    mq = rng.integers(0, 256, n).astype(np.uint8); sd = rng.integers(0, 2, n).astype(np.uint8)
    pk = PackedFragments(st, sp, mq, sd, pinned=False, threads=3, record_bytes=4)
    s2, e2, q2, d2, n_raw = _decode(pk, n)
    assert np.array_equal(s2, st) and np.array_equal(e2, sp) and np.array_equal(q2, mq) and np.array_equal(d2, sd)
    assert n_raw == pk.n_raw
    if n >= 1000:
        assert 3 <= pk.n_raw <= 4           # the three bad rows' blocks (+ the 2048 gap if it is not at a block edge)
    assert pk.wire_bytes() == pk.n_blocks * 64 * 4 + pk.n_blocks * 4
    # unsorted input is sorted first (stable), like ContigFragments
    if n == 1000:
        perm = rng.permutation(n)
        pk2 = PackedFragments(st[perm], sp[perm], mq[perm], sd[perm], pinned=False, record_bytes=4)
        s3, e3, _, _, _ = _decode(pk2, n)
        assert np.array_equal(s3, st) and np.array_equal(np.sort(e3 - s3), np.sort(ln))


def test_pack_capacity_error():
    import ctypes
    from finaletoolkit_b200._lib import lib
    L = lib()
    st = np.array([0, 5000], np.int32); sp = np.array([10, 5010], np.int32)
    P = ctypes.POINTER
    words = np.zeros(64, np.uint32); anch = np.zeros(1, np.int32)
    rc = L.ftk_pack_fragments_host(st.ctypes.data_as(P(ctypes.c_int32)), sp.ctypes.data_as(P(ctypes.c_int32)), None, None,
                                   2, 1, words.ctypes.data_as(P(ctypes.c_uint32)), anch.ctypes.data_as(P(ctypes.c_int32)),
                                   None, None, None, None, 0, 4)
    assert rc == -3      # FTK_E_RANGE: one raw block needed, no capacity given
