"""N-GPU check of the distributed genome-wide reductions and drivers (NCCL).  Spawned by
tests/test_gpu_dist.py (pytest -m gpu, needs >= 2 GPUs), or by hand:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/dist_gpu_check.py

Every rank computes the sharded result; rank 0 compares with the oracle's single-process stream.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from finaletoolkit_b200 import distributed as FD  # noqa: E402
from finaletoolkit_b200.io.fragments import FragmentTable  # noqa: E402
from finaletoolkit_b200.sharding import DistContext  # noqa: E402
from finaletoolkit_b200.synth import synth_fragments  # noqa: E402
from oracle import oracle as O  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ctx = DistContext()
contigs = {"1": 900_000, "2": 700_000, "3": 650_000, "4": 300_000, "5": 120_000, "6": 90_000, "7": 30_000}
cols = {c: synth_fragments(n, n // 4, i, seed_base=6000) for i, (c, n) in enumerate(contigs.items())}
table = FragmentTable(cols)
mine = FD.owned_contigs(table, ctx)
total = FD.genome_total_coverage(table, quality_threshold=30, ctx=ctx, device=dev)
d = FD.genome_length_distribution(table, min_length=0, max_length=None, quality_threshold=30, ctx=ctx, device=dev)
if ctx.rank == 0:
    frs = {c: O.Frags(*cols[c]) for c in contigs}
    exp_total = sum(O.single_coverage(fr, 0, None) for fr in frs.values())
    exp = O.merge_dists(O.length_dist(fr, None, None, 0, None, "midpoint", 30) for fr in frs.values())
    assert total == exp_total, (total, exp_total)
    assert d == exp and list(d) == list(exp), "genome-wide dict / first-seen order mismatch"
    print(f"dist check ok: world={ctx.world} rank0 owns {mine}; total={total}, {len(d)} distinct lengths, "
          f"first keys {list(d)[:5]}")

# motifs (end + breakpoint) and DELFI bins: sharded by contig, one all-reduce each
from finaletoolkit_b200 import device as D  # noqa: E402
from finaletoolkit_b200.synth import synth_twobit  # noqa: E402


class _Ref:
    def __init__(self):
        self.chroms = dict(contigs, extra=20_000)      # "extra": in the reference, not in the fragments
        self._packed = {c: synth_twobit(n, i, seed_base=6100, telomere=500, block_len=1000)
                        for i, (c, n) in enumerate(self.chroms.items())}

    def device_contig(self, c, device=None):
        return D.PackedContig.from_codes(*self._packed[c], device=device)

    def ascii(self, c):
        codes, nm = self._packed[c]
        seq = np.frombuffer(b"ACGT", np.uint8)[codes].copy(); seq[nm] = ord("N")
        return seq.tobytes()


ref = _Ref()
em = FD.genome_end_motif_counts(table, ref, k=3, quality_threshold=30, ctx=ctx, device=dev)
bm = FD.genome_end_motif_counts(table, ref, k=4, quality_threshold=30, ctx=ctx, device=dev, breakpoint=True)
bins = {c: (np.arange(0, n, 25_000), np.minimum(np.arange(0, n, 25_000) + 25_000, n)) for c, n in ref.chroms.items()}
blk = {"1": (np.array([100_000, 400_000]), np.array([110_000, 400_900]))}
gaps = {"2": ((300_000, 340_000), [(0, 5_000), (695_000, 700_000)])}
dw = FD.genome_delfi_windows(table, ref, bins, blk, gaps, quality_threshold=30, ctx=ctx, device=dev)
if ctx.rank == 0:
    from finaletoolkit_b200.frag._motif_common import genome_windows
    exp_em = sum(O.region_end_motifs(frs[c], ref.ascii(c), a, b, 3, True, False, 30) for c in contigs for a, b in genome_windows(contigs[c]))
    exp_bm = sum(O.region_breakpoint_motifs(frs[c], ref.ascii(c), a, b, 4, True, False, 30) for c in contigs for a, b in genome_windows(contigs[c]))
    assert np.array_equal(em, exp_em) and np.array_equal(bm, exp_bm), "motif counts mismatch"
    empty = O.Frags(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.uint8), np.zeros(0, np.uint8))
    for c in ref.chroms:
        fr = frs.get(c, empty)
        for j, (a, b) in enumerate(zip(*bins[c])):
            assert tuple(dw[c][j]) == O.delfi_counts(fr, ref.ascii(c), int(a), int(b), blk.get(c), gaps.get(c), 30), (c, a, b)
    print(f"dist check ok: motifs {int(em.sum())} / {int(bm.sum())} k-mers, DELFI {sum(int(v[:, 2].sum()) for v in dw.values())} "
          f"fragments in {sum(len(v) for v in dw.values())} bins")

# ---- genome drivers: LPT-sharded WPS + coverage + histogram, device-resident WPS -> adjust
sizes = list(contigs.items())
res, hist, tot = FD.multi_wps_genome(table, sizes, None, 5000, 120, 120, 180, 30, coverage=True, length_hist=True,
                                     adjust=dict(median_window_size=1000), ctx=ctx, device=dev)
assert sorted(res) == sorted(mine)
bad = 0
for c, r in res.items():
    fr = O.Frags(*cols[c])
    exp, off = O.wps_intervals(fr, r.starts, r.stops, contigs[c], 120, 120, 180, 30, threads=4)
    bad += int((r.wps.cpu().numpy().astype(np.int64) != exp).sum())
    bad += int((r.cov.cpu().numpy() != O.interval_coverage(fr, r.starts, r.stops, None, None, "midpoint", 30)).sum())
    for i in r.adj_segments[:3]:
        ref_adj = O.adjust_core(exp[off[i]: off[i + 1]].astype(np.float64), 1000, False, True, 21, 2)
        got = r.adjusted[r.adj_offsets[i]: r.adj_offsets[i + 1]].cpu().numpy()
        bad += int(not np.allclose(got, ref_adj, rtol=1e-5, atol=1e-9))
badt = torch.tensor([bad], device=dev); dist.all_reduce(badt)
frs_all = {c: O.Frags(*cols[c]) for c in contigs}
exp_hist = np.zeros(hist.numel(), np.int64)
for c, n in contigs.items():
    for L, v in O.length_dist(frs_all[c], 0, n, None, None, "midpoint", 30).items():
        exp_hist[L] += v
assert int(badt.item()) == 0, "multi_wps_genome mismatch vs oracle"
assert np.array_equal(hist.cpu().numpy(), exp_hist) and tot == int(exp_hist.sum())

# ---- the public API under torch.distributed: every rank gets the full answer, rank 0 writes
import tempfile  # noqa: E402
import finaletoolkit_b200 as F  # noqa: E402
tmp = tempfile.mkdtemp(prefix=f"ftk_dist_{ctx.rank}_")
bed = os.path.join(tmp, "iv.bed")
rows = [(c, a, min(a + 40_000, n)) for c, n in contigs.items() for a in range(0, n, 40_000)]
open(bed, "w").write("".join(f"{c}\t{a}\t{b}\tx\n" for c, a, b in rows))
out_bed = os.path.join(tmp, "cov.bed")
got = F.coverage(table, bed, out_bed, normalize=True, scale_factor=1e6, quality_threshold=30)
exp_total = sum(O.single_coverage(fr, 0, None) for fr in frs_all.values())
for (c, a, b), g in zip(rows, got):
    assert g.coverage == O.single_coverage(frs_all[c], a, b) * (1e6 / exp_total), (c, a, b)
assert os.path.exists(out_bed) == (ctx.rank == 0)
bins, counts = F.frag_length_bins(table, bin_size=5, quality_threshold=30)
exp_d = O.merge_dists(O.length_dist(fr, None, None, 0, None, "midpoint", 30) for fr in frs_all.values())
eb, ec = O.length_bins(exp_d, 5)
assert list(bins) == list(eb) and list(counts) == list(ec)
if ctx.rank == 0:
    print(f"dist check ok: multi_wps_genome {sum(len(r.starts) for r in res.values())} intervals on rank 0, "
          f"hist total {tot}; API coverage(normalize) + frag_length_bins equal the oracle on every rank")
dist.barrier()
dist.destroy_process_group()
