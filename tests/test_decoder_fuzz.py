"""Corrupted input must be rejected (or decoded to whatever well-formed records remain), never crash the
process: the native BAM and BGZF fragment-file decoders and the bigWig reader on a few hundred mutated copies of the
golden files -
byte flips, truncation, insertions, hostile values in the BAM record headers (block_size, l_read_name, n_cigar)
and in the BGZF member headers (XLEN, BSIZE, ISIZE).  Runs in a child process so that a crash is a test failure
rather than the end of the test session."""
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_CHILD = r'''
import ctypes, gzip, os, struct, sys, tempfile, zlib
import numpy as np
repo = sys.argv[1]
sys.path.insert(0, repo)
from finaletoolkit_b200._lib import lib
L = lib()
rng = np.random.default_rng(7)
d = tempfile.mkdtemp()
err = ctypes.c_int32(0)
p32 = ctypes.POINTER(ctypes.c_int32); p8 = ctypes.POINTER(ctypes.c_uint8)


def bgzf(data, block=5000):
    out = bytearray()
    for i in list(range(0, len(data), block)) + [None]:
        chunk = b"" if i is None else data[i:i + block]
        c = zlib.compressobj(1, zlib.DEFLATED, -15)
        payload = c.compress(chunk) + c.flush()
        out += struct.pack("<BBBBIBBH", 31, 139, 8, 4, 0, 0, 255, 6) + b"BC" + struct.pack("<HH", 2, 12 + 6 + len(payload) + 8 - 1)
        out += payload + struct.pack("<II", zlib.crc32(chunk), len(chunk))
    return bytes(out)


def mutate(b, header_fields):
    b = bytearray(b)
    k = int(rng.integers(0, 4))
    if k == 0:
        for _ in range(int(rng.integers(1, 12))):
            b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
    elif k == 1:
        b = b[: int(rng.integers(0, len(b)))]
    elif k == 2:
        i = int(rng.integers(0, len(b)))
        b[i:i] = bytes(rng.integers(0, 256, int(rng.integers(1, 64)), dtype=np.uint8))
    else:
        i = int(rng.integers(0, max(len(b) - 8, 1)))
        b[i:i + 4] = struct.pack("<i", int(rng.choice(header_fields)))
    return bytes(b)


def drain(ff, read1):
    for i in range(L.ftk_fragfile_n_contigs(ff)):
        n = L.ftk_fragfile_contig_count(ff, i)
        a, b = np.empty(n, np.int32), np.empty(n, np.int32)
        q, s = np.empty(n, np.uint8), np.empty(n, np.uint8)
        L.ftk_fragfile_copy(ff, i, a.ctypes.data_as(p32), b.ctypes.data_as(p32), q.ctypes.data_as(p8), s.ctypes.data_as(p8))
        if read1:
            L.ftk_fragfile_copy_read1(ff, i, a.ctypes.data_as(p32), b.ctypes.data_as(p32))


g = dict(np.load(os.path.join(repo, "tests", "golden", "bam_read1.npz")))
bam_file = g["bam_file"].tobytes()
bam_raw = gzip.decompress(bam_file)
fx = dict(np.load(os.path.join(repo, "tests", "golden", "fixture17.npz")))
frag_file = fx["frag_gz_file"].tobytes()
frag_raw = gzip.decompress(frag_file)
hostile = [0, -1, 31, 32, 33, 2 ** 31 - 1, -2 ** 31, 100000, 65535, 65536]
opened = rejected = 0
for it in range(240):
    path = os.path.join(d, "f.bam")
    body = mutate(bam_file, hostile) if it % 3 == 0 else bgzf(mutate(bam_raw, hostile))     # container / record level
    open(path, "wb").write(body)
    h = L.ftk_bamfile_open(path.encode(), 2, ctypes.byref(err))
    if h:
        opened += 1
        drain(L.ftk_bamfile_fragments(h), True)
        L.ftk_bamfile_close(h)
    else:
        rejected += 1
for it in range(160):
    path = os.path.join(d, "f.frag.gz")
    body = mutate(frag_file, hostile) if it % 2 == 0 else bgzf(mutate(frag_raw, hostile), block=700)
    open(path, "wb").write(body)
    h = L.ftk_fragfile_open(path.encode(), 2, ctypes.byref(err))
    if h:
        opened += 1
        drain(h, False)
        L.ftk_fragfile_close(h)
    else:
        rejected += 1
# bigWig reader: the R-tree / section offsets of a damaged file must never reach the native inflate as bad pointers
from finaletoolkit_b200.io import bigwig
bw = dict(np.load(os.path.join(repo, "tests", "golden", "agg.npz")))["ref_test_bw"].tobytes()
for it in range(150):
    b = bytearray(bw)
    k = int(rng.integers(0, 3))
    if k == 0:
        for _ in range(int(rng.integers(1, 10))):
            b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
    elif k == 1:
        b = b[: int(rng.integers(0, len(b)))]
    else:
        i = int(rng.integers(0, len(b) - 8))
        b[i:i + 8] = struct.pack("<q", int(rng.choice([0, -1, 2 ** 40, len(b) + 5, 2 ** 62])))
    path = os.path.join(d, "x.bw")
    open(path, "wb").write(bytes(b))
    try:
        r = bigwig.open(path)
        for c in list(r.chroms())[:3]:
            try:
                r.values(c, 0, min(int(r.chroms(c) or 1), 5000))
                r.intervals(c, 0, 0)
            except Exception:       # a Python exception is a rejection, not a crash
                pass
        opened += 1
    except Exception:
        rejected += 1
print("fuzz ok", opened, rejected)
'''


def test_native_decoders_survive_corrupted_files(tmp_path):
    script = tmp_path / "child.py"
    script.write_text(_CHILD)
    r = subprocess.run([sys.executable, str(script), REPO], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stderr[-2000:])
    last = r.stdout.strip().splitlines()[-1].split()
    assert last[:2] == ["fuzz", "ok"] and int(last[2]) > 20 and int(last[3]) > 20      # both outcomes exercised
