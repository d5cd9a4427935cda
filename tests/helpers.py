"""Test helpers: materialise golden inputs as the files the public API reads."""
import gzip
import struct

import numpy as np


def write_frag_gz(path, columns, bed6=False):
    """columns: {contig: (start, stop, mapq, strand)} -> (b)gzip text + a placeholder .tbi."""
    with gzip.open(path, "wt") as fh:
        for contig, (st, sp, mq, sd) in columns.items():
            for a, b, q, s in zip(st.tolist(), sp.tolist(), mq.tolist(), sd.tolist()):
                if bed6:
                    fh.write(f"{contig}\t{a}\t{b}\t.\t{q}\t{'+' if s else '-'}\n")
                else:
                    fh.write(f"{contig}\t{a}\t{b}\t{q}\t{'+' if s else '-'}\n")
    open(str(path) + ".tbi", "wb").close()
    return str(path)


def write_text_gz(path, text):
    with gzip.open(path, "wt") as fh:
        fh.write(text)
    open(str(path) + ".tbi", "wb").close()
    return str(path)


def write_2bit(path, seqs):
    """seqs: [(name, codes A0C1G2T3 uint8, n_mask bool)] -> UCSC .2bit (T0 C1 A2 G3, MSB first)."""
    remap = np.array([2, 1, 3, 0], np.uint8)
    recs = []
    for name, codes, nm in seqs:
        n = codes.shape[0]
        u = remap[codes].copy()
        u[nm] = 0
        q = np.concatenate([u, np.zeros((-n) % 4, np.uint8)]).reshape(-1, 4)
        packed = ((q[:, 0] << 6) | (q[:, 1] << 4) | (q[:, 2] << 2) | q[:, 3]).astype(np.uint8)
        d = np.diff(np.concatenate([[0], nm.astype(np.int8), [0]]))
        starts, ends = np.flatnonzero(d == 1), np.flatnonzero(d == -1)
        body = struct.pack("<II", n, len(starts)) + starts.astype("<u4").tobytes() + (ends - starts).astype("<u4").tobytes()
        body += struct.pack("<II", 0, 0) + packed.tobytes()
        recs.append((name, body))
    off = 16 + sum(1 + len(n.encode()) + 4 for n, _ in recs)
    idx = b""
    for name, body in recs:
        idx += bytes([len(name.encode())]) + name.encode() + struct.pack("<I", off)
        off += len(body)
    with open(path, "wb") as fh:
        fh.write(struct.pack("<IIII", 0x1A412743, 0, len(recs), 0) + idx + b"".join(b for _, b in recs))
    return str(path)


def golden_codes(g, name, n):
    codes = np.unpackbits(g[f"{name}_codes_packed"]).reshape(-1, 2)[:n]
    return (codes[:, 0] * 2 + codes[:, 1]).astype(np.uint8), np.unpackbits(g[f"{name}_nmask_packed"])[:n].astype(bool)


def read_gz(path):
    with gzip.open(path, "rt") as fh:
        return fh.read()


def write_bgzf(path, text: str, block=30000):
    """Real BGZF (the container tabix/bgzip write): independent deflate members with a 'BC' extra
    field holding the block size, terminated by the 28-byte EOF block; + a placeholder .tbi."""
    import zlib
    data = text.encode()
    with open(path, "wb") as fh:
        for i in list(range(0, len(data), block)) + [None]:
            chunk = b"" if i is None else data[i:i + block]
            c = zlib.compressobj(6, zlib.DEFLATED, -15)
            payload = c.compress(chunk) + c.flush()
            bsize = 12 + 6 + len(payload) + 8 - 1
            fh.write(struct.pack("<BBBBIBBH", 31, 139, 8, 4, 0, 0, 255, 6) + b"BC" + struct.pack("<HH", 2, bsize))
            fh.write(payload + struct.pack("<II", zlib.crc32(chunk), len(chunk)))
    open(str(path) + ".tbi", "wb").close()
    return str(path)


def delfi_tracks(m):
    """Blacklist and gap tracks of the DELFI golden, parsed the way the reference parses the files
    (frag/_delfi.py:85-107; genome/gaps.py:54-62,170-182).

    Returns ``(blacklist {contig: (starts, stops)}, gaps {contig: (centromere, telomeres, has_short_arm)})``."""
    bl = {}
    for line in m["blacklist"].splitlines():
        p = line.split()
        if len(p) >= 3:
            bl.setdefault(p[0], []).append((int(p[1]), int(p[2])))
    bl = {c: (np.array([a for a, _ in sorted(r)], np.int64), np.array([b for _, b in sorted(r)], np.int64)) for c, r in bl.items()}
    rows = [ln.split() for ln in m["gaps"].splitlines() if ln.strip()]
    gaps = {}
    for c in {r[0] for r in rows}:
        cen = [(int(r[1]), int(r[2])) for r in rows if r[0] == c and r[3] == "centromere"]
        if cen:
            gaps[c] = (cen[0], [(int(r[1]), int(r[2])) for r in rows if r[0] == c and r[3] == "telomere"],
                       any(r[0] == c and r[3] == "short_arm" for r in rows))
    return bl, gaps


def agg_fixture(g, m):
    """(signals, strands, write_bw) of the agg_bw golden: per-interval ``values`` rebuilt from the stored
    signal layout (None where pyBigWig raises), and a function writing the same track as a real bigWig."""
    sizes = dict((c, n) for c, n in m["sizes"])

    def values(contig, start, stop):
        if contig not in sizes or stop <= start or stop > sizes[contig] or start < 0:
            return None
        out = np.full(stop - start, np.nan, np.float32)
        for c, pos, lo, hi in m["layout"]:
            if c != contig:
                continue
            a, b = max(pos, start), min(pos + hi - lo, stop)
            if b > a:
                out[a - start: b - start] = g[f"{c}_signal"][lo + a - pos: lo + b - pos]
        return out

    rows = [ln.split("\t") for ln in m["bed"].splitlines()]
    signals = [values(r[0], int(r[1]), int(r[2])) for r in rows]
    strands = [r[5].strip() for r in rows]

    def write_bw(path):
        from finaletoolkit_b200.io import bigwig
        with bigwig.open(str(path), "w") as w:
            w.addHeader([tuple(x) for x in m["sizes"]])
            for c, pos, lo, hi in m["layout"]:
                w.addEntries(c, pos, values=g[f"{c}_signal"][lo:hi].astype(np.float64), span=1, step=1)
        return str(path)

    return signals, strands, write_bw


def _reg2bin(beg, end):
    """UCSC / tabix binning (SAM spec 5.3), 0-based half-open [beg, end)."""
    end -= 1
    for shift, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        if beg >> shift == end >> shift:
            return base + (beg >> shift)
    return 0


def write_bgzf_indexed(path, text: str, block=30000):
    """``write_bgzf`` plus a REAL tabix index (BED-style: -s 1 -b 2 -e 3 -0) built the way htslib builds
    it: binning index with merged chunks, 16-kb linear index, the 37450 pseudo-bin, BGZF-compressed."""
    import zlib
    data = text.encode()
    coffs, sizes = [], []
    with open(path, "wb") as fh:
        for i in list(range(0, len(data), block)) + [None]:
            chunk = b"" if i is None else data[i:i + block]
            c = zlib.compressobj(6, zlib.DEFLATED, -15)
            payload = c.compress(chunk) + c.flush()
            bsize = 12 + 6 + len(payload) + 8 - 1
            coffs.append(fh.tell()); sizes.append(len(chunk))
            fh.write(struct.pack("<BBBBIBBH", 31, 139, 8, 4, 0, 0, 255, 6) + b"BC" + struct.pack("<HH", 2, bsize))
            fh.write(payload + struct.pack("<II", zlib.crc32(chunk), len(chunk)))

    def voff(t):   # virtual offset of text position t; a position at a block end is the NEXT block's start,
        k = t // block   # as bgzf_tell reports it (for the end of the data that is the EOF block)
        return (coffs[-1] << 16) if t >= len(data) else (coffs[k] << 16) | (t - k * block)

    names, per_ref = [], {}
    t = 0
    for line in data.split(b"\n")[:-1]:
        t0, t = t, t + len(line) + 1
        if not line or line.startswith(b"#"):
            continue
        f = line.split(b"\t")
        name, beg, end = f[0].decode(), int(f[1]), int(f[2])
        if name not in per_ref:
            names.append(name); per_ref[name] = dict(bins={}, lin={}, beg=voff(t0), end=None, n=0)
        r = per_ref[name]
        b = _reg2bin(beg, max(end, beg + 1))
        ch = r["bins"].setdefault(b, [])
        if ch and ch[-1][1] >> 16 == voff(t0) >> 16:      # htslib merges chunks that touch the same block
            ch[-1][1] = voff(t)
        else:
            ch.append([voff(t0), voff(t)])
        for w in range(beg >> 14, ((max(end, beg + 1) - 1) >> 14) + 1):
            r["lin"].setdefault(w, voff(t0))
        r["end"], r["n"] = voff(t), r["n"] + 1
    out = bytearray(b"TBI\x01" + struct.pack("<8i", len(names), 0x10000, 1, 2, 3, ord("#"), 0, sum(len(n) + 1 for n in names)))
    for n in names:
        out += n.encode() + b"\0"
    for n in names:
        r = per_ref[n]
        out += struct.pack("<i", len(r["bins"]) + 1)
        out += struct.pack("<IiQQQQ", 37450, 2, r["beg"], r["end"], r["n"], 0)
        for b, ch in r["bins"].items():
            out += struct.pack("<Ii", b, len(ch)) + b"".join(struct.pack("<QQ", a, e) for a, e in ch)
        n_intv = max(r["lin"]) + 1
        lin, last = [], 0
        for w in range(n_intv):
            last = r["lin"].get(w, last); lin.append(last)
        out += struct.pack("<i", n_intv) + struct.pack(f"<{n_intv}Q", *lin)
    out += struct.pack("<Q", 0)
    with open(str(path) + ".tbi", "wb") as fh:       # BGZF: one data block + the EOF block
        for chunk in (bytes(out), b""):
            c = zlib.compressobj(6, zlib.DEFLATED, -15)
            payload = c.compress(chunk) + c.flush()
            bsize = 12 + 6 + len(payload) + 8 - 1
            fh.write(struct.pack("<BBBBIBBH", 31, 139, 8, 4, 0, 0, 255, 6) + b"BC" + struct.pack("<HH", 2, bsize))
            fh.write(payload + struct.pack("<II", zlib.crc32(chunk), len(chunk)))
    return str(path)


def write_bam(path, refs, records, block=4000):
    """Minimal BAM writer for tests: ``refs`` = [(name, length)], ``records`` = dicts with ref (index or -1),
    pos, mapq, flag, tlen and cigar [(op, len)]; BGZF blocks of ``block`` bytes so records straddle them.
    Also writes a placeholder .bai (the loader only checks that it exists)."""
    import zlib
    text = b"@HD\tVN:1.6\tSO:coordinate\n" + b"".join(f"@SQ\tSN:{n}\tLN:{ln}\n".encode() for n, ln in refs)
    out = bytearray(b"BAM\x01" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(refs)))
    for n, ln in refs:
        out += struct.pack("<i", len(n) + 1) + n.encode() + b"\0" + struct.pack("<i", ln)
    for k, r in enumerate(records):
        name = f"r{k}".encode() + b"\0"
        cig = r.get("cigar", [(0, 50)])
        l_seq = sum(ln for op, ln in cig if op in (0, 1, 4, 7, 8))
        body = struct.pack("<iiBBHHHiiii", r["ref"], r["pos"], len(name), r["mapq"], 4680, len(cig), r["flag"], l_seq,
                           r["ref"], r["pos"] + 100, r["tlen"])
        body += name + b"".join(struct.pack("<I", (ln << 4) | op) for op, ln in cig)
        body += bytes((l_seq + 1) // 2) + bytes([30]) * l_seq + b"NMC\x00"
        out += struct.pack("<i", len(body)) + body
    data = bytes(out)
    with open(path, "wb") as fh:
        for i in list(range(0, len(data), block)) + [None]:
            chunk = b"" if i is None else data[i:i + block]
            c = zlib.compressobj(6, zlib.DEFLATED, -15)
            payload = c.compress(chunk) + c.flush()
            bsize = 12 + 6 + len(payload) + 8 - 1
            fh.write(struct.pack("<BBBBIBBH", 31, 139, 8, 4, 0, 0, 255, 6) + b"BC" + struct.pack("<HH", 2, bsize))
            fh.write(payload + struct.pack("<II", zlib.crc32(chunk), len(chunk)))
    open(str(path) + ".bai", "wb").close()
    return str(path)
