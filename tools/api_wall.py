"""Profiling helper (PYTHONPATH=.:tests python tools/api_wall.py): wall-clock of the public API on a 5 M-fragment
BGZF file (60 Mb contig, 12 000 5-kb intervals) - finds host-side bottlenecks around the kernels."""
import os, sys, time, tempfile
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from helpers import write_bgzf, write_2bit
import finaletoolkit_b200 as F
from finaletoolkit_b200.synth import synth_fragments, synth_twobit

tmp = tempfile.mkdtemp(prefix="ftk_wall_")
n, clen = int(os.environ.get("NFRAG", 5_000_000)), 60_000_000
st, sp, mq, sd = synth_fragments(clen, n, 1)
t0 = time.time()
parts = np.char.add(np.char.add(np.char.add(np.char.add(np.char.add("1\t", st.astype(str)), "\t"), sp.astype(str)), "\t"), mq.astype(str))
txt = "\n".join(np.char.add(parts, np.where(sd == 1, "\t+", "\t-")).tolist()) + "\n"
frag = write_bgzf(os.path.join(tmp, "big.frag.gz"), txt); del txt, parts
codes, nm = synth_twobit(clen, 1)
tb = write_2bit(os.path.join(tmp, "ref.2bit"), [("1", codes, nm)])
cs = os.path.join(tmp, "cs"); open(cs, "w").write(f"1\t{clen}\n")
bed = os.path.join(tmp, "sites.bed")
open(bed, "w").write("".join(f"1\t{a}\t{a + 1}\t.\t0\t+\n" for a in range(2500, clen - 2500, 5000)))
tiles = os.path.join(tmp, "tiles.bed")
open(tiles, "w").write("".join(f"1\t{a}\t{a + 5000}\n" for a in range(0, clen - 5000, 5000)))
print(f"inputs built in {time.time() - t0:.1f}s", flush=True)


def wall(name, fn, reps=2):
    for r in range(reps):
        t0 = time.time(); out = fn(); dt = time.time() - t0
        print(f"{name:34s} run {r}: {dt * 1e3:9.1f} ms", flush=True)
    return out


wall("frag_length_bins", lambda: F.frag_length_bins(frag, "1", 0, clen, bin_size=5))
wall("coverage (12k intervals) -> bed", lambda: F.coverage(frag, tiles, os.path.join(tmp, "cov.bed"), normalize=False))
wall("frag_length_intervals", lambda: F.frag_length_intervals(frag, tiles, os.path.join(tmp, "fli.bed")))
wall("multi_wps -> .bw", lambda: F.multi_wps(frag, bed, chrom_sizes=cs, output_file=os.path.join(tmp, "wps.bw")))
wall("adjust_wps .bw -> .bw", lambda: F.adjust_wps(os.path.join(tmp, "wps.bw"), bed, os.path.join(tmp, "adj.bw"), cs))
agg_bed = os.path.join(tmp, "agg.bed")
open(agg_bed, "w").write("".join(f"1\t{a + 600}\t{a + 4400}\t.\t0\t{'+-'[(a // 5000) & 1]}\n" for a in range(0, clen - 5000, 5000)))
wall("agg_bw (12k x 3800)", lambda: F.agg_bw(os.path.join(tmp, "adj.bw"), agg_bed, os.path.join(tmp, "agg.wig"), 0))
wall("end_motifs k=4", lambda: F.end_motifs(frag, tb, k=4))
wall("interval_end_motifs", lambda: F.interval_end_motifs(frag, tb, tiles, k=4))
wall("breakpoint_motifs k=6", lambda: F.breakpoint_motifs(frag, tb))
wall("multi_cleavage_profile -> .bw", lambda: F.multi_cleavage_profile(frag, tiles, cs, output_file=os.path.join(tmp, "clv.bw")))
bins = os.path.join(tmp, "bins.bed")
open(bins, "w").write("".join(f"1\t{a}\t{a + 100000}\n" for a in range(0, clen, 100000)))
wall("delfi (no LOESS, no merge)", lambda: F.delfi(frag, cs, bins, tb, no_gc_correct=True, merge_bins=False, remove_nocov=False))
