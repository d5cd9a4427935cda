"""Synthetic fragment sets and 2-bit genomes.

Host-side numpy only.  The distributions are the ones the measurement
contract fixes so that the CUDA path, its CPU checker and the CPU baseline all see
identical inputs:

* ``start ~ U{0 .. len-L}``, then sorted ascending (coordinate-sorted like a
  tabix-indexed ``.frag.gz``);
* ``L`` from ``0.85*N(167,15^2) + 0.12*N(330,25^2) + 0.03*U{50..500}``,
  rounded, clipped to ``[30, 600]``;
* ``mapq`` = 60 w.p. 0.9 else ``U{0..59}``; strand Bernoulli(0.5);
* ``numpy.random.Generator(PCG64(seed=1000 + contig_index))``.
"""
from __future__ import annotations

import numpy as np

__all__ = ["B37_CONTIGS", "synth_fragments", "synth_twobit", "pack_twobit"]

# First 24 rows of the b37 chrom.sizes used by the reference's tests
# (reference: tests/data/b37.chrom.sizes:1-24).
B37_CONTIGS = [
    ("1", 249250621), ("2", 243199373), ("3", 198022430), ("4", 191154276),
    ("5", 180915260), ("6", 171115067), ("7", 159138663), ("8", 146364022),
    ("9", 141213431), ("10", 135534747), ("11", 135006516), ("12", 133851895),
    ("13", 115169878), ("14", 107349540), ("15", 102531392), ("16", 90354753),
    ("17", 81195210), ("18", 78077248), ("19", 59128983), ("20", 63025520),
    ("21", 48129895), ("22", 51304566), ("X", 155270560), ("Y", 59373566),
]


def synth_fragments(contig_len: int, n: int, contig_index: int = 0,
                    seed_base: int = 1000):
    """Return ``(start, stop, mapq, strand)`` as (int32, int32, uint8, uint8).

    ``strand`` is 1 for '+'.  Rows are sorted by ``start`` (stable).
    """
    rng = np.random.Generator(np.random.PCG64(seed_base + contig_index))
    comp = rng.random(n)
    length = np.empty(n, dtype=np.float64)
    m0 = comp < 0.85
    m1 = (comp >= 0.85) & (comp < 0.97)
    m2 = comp >= 0.97
    length[m0] = rng.normal(167.0, 15.0, int(m0.sum()))
    length[m1] = rng.normal(330.0, 25.0, int(m1.sum()))
    length[m2] = rng.integers(50, 501, int(m2.sum()))
    length = np.clip(np.rint(length), 30, 600).astype(np.int64)
    length = np.minimum(length, max(contig_len, 1))
    start = np.floor(rng.random(n) * (contig_len - length + 1)).astype(np.int64)
    mapq = np.where(rng.random(n) < 0.9, 60, rng.integers(0, 60, n)).astype(np.uint8)
    strand = (rng.random(n) < 0.5).astype(np.uint8)
    order = np.argsort(start, kind="stable")
    start = start[order]
    stop = start + length[order]
    return (start.astype(np.int32), stop.astype(np.int32), mapq[order], strand[order])


def pack_twobit(codes: np.ndarray, n_mask: np.ndarray):
    """Pack per-base codes (A0 C1 G2 T3) and an N mask into device words.

    Layout (the one ``ftk_end_motif_hist_u64`` reads): base ``i`` lives in bits
    ``2*(i%16) .. 2*(i%16)+1`` of ``seq_words[i//16]`` (little-endian within the
    word); N flag of base ``i`` is bit ``i%32`` of ``nmask_words[i//32]``.
    Both arrays are padded with two spare words so that an unaligned 64-bit
    window read never runs past the allocation.
    """
    n = int(codes.shape[0])
    nw = (n + 15) // 16 + 2
    padded = np.zeros(nw * 16, dtype=np.uint32)
    padded[:n] = codes.astype(np.uint32) & 3
    shifts = (np.arange(16, dtype=np.uint32) * 2)[None, :]
    seq_words = np.bitwise_or.reduce(padded.reshape(nw, 16) << shifts, axis=1).astype(np.uint32)
    nm = (n + 31) // 32 + 2
    mpad = np.zeros(nm * 32, dtype=np.uint8)
    mpad[:n] = n_mask.astype(np.uint8)
    nmask_words = np.packbits(mpad.reshape(nm, 32), axis=1, bitorder="little").view("<u4").reshape(nm)
    return seq_words, nmask_words.astype(np.uint32)


def synth_twobit(contig_len: int, contig_index: int = 0, seed_base: int = 2000,
                 telomere: int = 10_000, n_blocks: int = 3, block_len: int = 50_000):
    """Random ACGT contig with N-blocks: returns ``(codes uint8 A0C1G2T3, n_mask bool)``."""
    rng = np.random.Generator(np.random.PCG64(seed_base + contig_index))
    codes = rng.integers(0, 4, contig_len, dtype=np.uint8)
    n_mask = np.zeros(contig_len, dtype=bool)
    t = min(telomere, contig_len // 4)
    n_mask[:t] = True
    n_mask[contig_len - t:] = True
    bl = min(block_len, max(contig_len // 20, 1))
    for _ in range(n_blocks):
        s = int(rng.integers(0, max(contig_len - bl, 1)))
        n_mask[s: s + bl] = True
    return codes, n_mask


def synth_fragments_device(contig_len: int, n: int, contig_index: int, device, seed_base: int = 3000):
    """Same distributions as ``synth_fragments`` drawn ON the GPU (torch Philox generator seeded with
    ``seed_base + contig_index``): returns start-sorted CUDA tensors ``(start int32, stop int32,
    mapq uint8)``.  Used for the genome-scale configuration (1e9 fragments), where drawing and sorting
    on the host would take minutes; checkers copy the columns they need back to the host."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed_base + contig_index)
    comp = torch.rand(n, generator=g, device=device)
    length = torch.empty(n, dtype=torch.float32, device=device).normal_(167.0, 15.0, generator=g)
    di = torch.empty(n, dtype=torch.float32, device=device).normal_(330.0, 25.0, generator=g)
    length = torch.where(comp >= 0.85, di, length)
    del di
    uni = torch.randint(50, 501, (n,), generator=g, device=device, dtype=torch.int32).to(torch.float32)
    length = torch.where(comp >= 0.97, uni, length)
    del uni, comp
    length = length.round_().clamp_(30, 600).to(torch.int32).clamp_(max=max(int(contig_len), 1))
    u = torch.rand(n, generator=g, device=device, dtype=torch.float64)
    start = (u * (int(contig_len) - length.to(torch.float64) + 1.0)).floor_().to(torch.int32)
    del u
    mapq = torch.where(torch.rand(n, generator=g, device=device) < 0.9,
                       torch.full((n,), 60, dtype=torch.int32, device=device),
                       torch.randint(0, 60, (n,), generator=g, device=device, dtype=torch.int32)).to(torch.uint8)
    start, order = torch.sort(start, stable=True)
    length = length[order]
    mapq = mapq[order]
    del order
    return start, start + length, mapq
