// 5' end-motif k-mer histograms on sm_100a.
//
// Replaces the per-fragment loop of region_end_motifs (frag/_end_motifs.py:115-179):
// two py2bit sequence() calls + a dict increment per fragment become two 64-bit
// window reads of a 2-bit packed contig (L2-resident: chr1 = 62 MB + 31 MB N-mask)
// and shared-memory-privatised histogram atomics.
//   forward k-mer : ref[fs, fs+k)              index = sum code_j * 4^(k-1-j)
//   reverse k-mer : revcomp(ref[fe-k, fe))     (utils/utils.py:413-437)
//   k-mers containing N are skipped (frag/_end_motifs.py:133,142)
//   forward window out of bounds -> fragment skipped entirely (:135-136)
//   reverse window out of bounds -> RuntimeError in the reference (:144-151):
//       reported through *error_flag, the host raises.
// Membership is tabix overlap only - no length filter, no midpoint policy
// (frag/_end_motifs.py:115-120).
// Index order = itertools.product("ACGT") (utils/utils.py:388-410): A0 C1 G2 T3.
// Breakpoint motifs (region_breakpoint_motifs, frag/_breakpoint_motifs.py:53-196) are the same
// kernel with windows centred on the two breakpoints, h = k/2:
//   fragment skipped when fs-h < 0 or fs+h >= contig_len (:125-133)
//   forward k-mer : ref[fs-h, fs+h)            reverse : revcomp(ref[fe-h, fe+h))
//   reverse window out of bounds -> that end is skipped (OutOfBoundsError is a ValueError, :186)
//   odd k: every window has 2h != k bases -> "length discrepancy" -> nothing is ever counted
// Roofline: HBM, 10 B per candidate fragment (start, stop, mapq, strand) + 2 x 2 B
// of packed reference per end served from L2.
#include "ftk_common.cuh"

namespace ftk {

constexpr int kMotifThreads = 256;
constexpr int kMotifSmemBins = 4096;  // k <= 6 privatised in shared memory


// 2k-bit window starting at base `pos` (base i at bits 2*(i%16) of word i/16): one funnel shift
// over two consecutive words; k <= 12, so 24 bits at a shift <= 30 always fit.
__device__ __forceinline__ uint32_t window2(const uint32_t *__restrict__ seq, int pos, uint32_t mask2k) {
    const int w = pos >> 4;
    return __funnelshift_r(__ldg(seq + w), __ldg(seq + w + 1), (pos & 15) * 2) & mask2k;
}
__device__ __forceinline__ bool has_n(const uint32_t *__restrict__ nmask, int pos, uint32_t maskk) {
    const int w = pos >> 5;
    return (__funnelshift_r(__ldg(nmask + w), __ldg(nmask + w + 1), pos & 31) & maskk) != 0u;
}
// reverse the order of the k 2-bit digits of x (first base becomes the most significant digit)
__device__ __forceinline__ uint32_t digit_reverse(uint32_t x, int k) {
    x = __brev(x);
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    return x >> (32 - 2 * k);
}

__global__ void motif_ranges_kernel(const int32_t *__restrict__ frag_start, int64_t n_frag,
                                    const int32_t *__restrict__ ivl_start,
                                    const int32_t *__restrict__ ivl_stop, int64_t n_ivl,
                                    int halo, int64_t *__restrict__ ranges) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * n_ivl) return;
    const int64_t k = t >> 1;
    // overlap: fe > S (=> fs > S - maxL) and fs < E
    const int64_t key = (t & 1) ? (int64_t)ivl_stop[k] : (int64_t)ivl_start[k] - halo;
    ranges[t] = lower_bound(frag_start, n_frag, key);
}

// Fragments are fetched four per lane with 128-bit loads (two vectors = eight fragments in flight per
// thread); the k-mer windows are read from the L2-resident packed contig.  (A variant that staged the
// contig span of every 4096-fragment sub-chunk in shared memory was measured in round 2 and lost:
// 0.59 ms against 0.39 ms - the staging serialises bounds -> span -> fragments behind two barriers
// per sub-chunk and its 75 registers cut the residency to three CTAs per SM.)
template <bool SMEM, bool BREAKPOINT>
__global__ void __launch_bounds__(kMotifThreads, 6)
end_motif_kernel(const int32_t *__restrict__ frag_start, const int32_t *__restrict__ frag_stop,
                 const uint8_t *__restrict__ frag_mapq, const uint8_t *__restrict__ frag_strand,
                 const uint32_t *__restrict__ seq, const uint32_t *__restrict__ nmask, int64_t contig_len,
                 const int32_t *__restrict__ ivl_start, const int32_t *__restrict__ ivl_stop,
                 const int64_t *__restrict__ ranges, int k, int strand_mode, int min_mapq,
                 int pooled, int splits, unsigned long long *__restrict__ counts,
                 int32_t *__restrict__ error_flag) {
    __shared__ int s_cnt[SMEM ? kMotifSmemBins : 1];
    const int tid = threadIdx.x;
    const int n_bins = 1 << (2 * k);
    const int64_t ivl = blockIdx.x / splits;
    const int split = blockIdx.x % splits;
    unsigned long long *__restrict__ row = counts + (pooled ? 0 : ivl * (int64_t)n_bins);
    const int S = ivl_start[ivl], E = ivl_stop[ivl];
    const int64_t lo_all = ranges[2 * ivl], hi_all = ranges[2 * ivl + 1];
    int64_t chunk = (hi_all - lo_all + splits - 1) / splits;
    chunk = (chunk + 3) & ~(int64_t)3;
    const int64_t lo = lo_all + (int64_t)split * chunk;
    const int64_t hi = min(hi_all, lo + chunk);

    if (SMEM) {
        for (int b = tid; b < n_bins; b += kMotifThreads) s_cnt[b] = 0;
        __syncthreads();
    }
    auto bump = [&](uint32_t idx) {
        if (SMEM) atomicAdd(&s_cnt[idx], 1); else atomicAdd(&row[idx], 1ull);
    };

    // 32-bit everything inside the slice: contig positions are int32 and a slice is < 2^31 fragments
    const uint32_t mask2k = (1u << (2 * k)) - 1u, maskk = (1u << k) - 1u;
    const int len32 = (int)min(contig_len, (int64_t)INT32_MAX);
    const int h = k >> 1;
    auto visit = [&](int fs, int fe, int q, int sd) {
        if (q < min_mapq || !(fe > S && fs < E)) return;
        if (BREAKPOINT) {
            if (fs < h || fs >= len32 - h) return;                       // too close to a contig end
            if ((k & 1) != 0) return;                                    // 2h != k: never counted
            if (strand_mode == 0 || (strand_mode == 1 && sd)) {
                if (!has_n(nmask, fs - h, maskk)) bump(digit_reverse(window2(seq, fs - h, mask2k), k));
            }
            if (strand_mode != 1) {
                if (fe < h || fe > len32 - h) return;                    // OutOfBoundsError -> skipped
                if (!has_n(nmask, fe - h, maskk)) bump((~window2(seq, fe - h, mask2k)) & mask2k);
            }
            return;
        }
        if (strand_mode == 1 && !sd) return;         // forward-only: '+' fragments only
        if (strand_mode != 2) {
            if (fs < 0 || fs > len32 - k) return;    // ValueError -> `continue`
            if (!has_n(nmask, fs, maskk)) bump(digit_reverse(window2(seq, fs, mask2k), k));
        }
        if (strand_mode != 1) {
            if (fe < k || fe > len32) {
                if (strand_mode == 0) atomicOr(error_flag, 1);  // RuntimeError in the reference
                return;
            }
            if (!has_n(nmask, fe - k, maskk)) bump((~window2(seq, fe - k, mask2k)) & mask2k);
        }
    };
    // the slice is widened to a 16-byte boundary on the left so that every lane loads 4 fragments per
    // 128-bit load; fragments left of lo are masked by index
    const int64_t lo_al = lo & ~(int64_t)3;
    const int skip = (int)(lo - lo_al);
    const int cnt = (hi > lo) ? (int)(hi - lo_al) : 0;
    const int nvec = cnt >> 2;
    const int4 *__restrict__ vs = reinterpret_cast<const int4 *>(frag_start + lo_al);
    const int4 *__restrict__ ve = reinterpret_cast<const int4 *>(frag_stop + lo_al);
    const uchar4 *__restrict__ vq = reinterpret_cast<const uchar4 *>(frag_mapq ? frag_mapq + lo_al : nullptr);
    const uchar4 *__restrict__ vd = reinterpret_cast<const uchar4 *>(frag_strand ? frag_strand + lo_al : nullptr);
    constexpr int kU = 2;
    for (int v0 = tid; v0 < nvec; v0 += kU * kMotifThreads) {
        int4 s4[kU], e4[kU];
        uchar4 q4[kU], d4[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int v = v0 + u * kMotifThreads;
            if (v < nvec) {
                s4[u] = __ldcs(vs + v); e4[u] = __ldcs(ve + v);
                q4[u] = vq ? __ldcs(vq + v) : make_uchar4(255, 255, 255, 255);
                d4[u] = vd ? __ldcs(vd + v) : make_uchar4(1, 1, 1, 1);
            } else {
                s4[u] = make_int4(0, 0, 0, 0); e4[u] = make_int4(0, 0, 0, 0);
                q4[u] = make_uchar4(0, 0, 0, 0); d4[u] = make_uchar4(1, 1, 1, 1);
            }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int v = v0 + u * kMotifThreads;
            if (v >= nvec) continue;
            const bool head = (v == 0);
            if (!(head && skip > 0)) visit(s4[u].x, e4[u].x, q4[u].x, d4[u].x);
            if (!(head && skip > 1)) visit(s4[u].y, e4[u].y, q4[u].y, d4[u].y);
            if (!(head && skip > 2)) visit(s4[u].z, e4[u].z, q4[u].z, d4[u].z);
            visit(s4[u].w, e4[u].w, q4[u].w, d4[u].w);
        }
    }
    {   // tail: at most 3 fragments
        const int i = nvec * 4 + tid;
        if (i < cnt && i >= skip)
            visit(__ldcs(frag_start + lo_al + i), __ldcs(frag_stop + lo_al + i),
                  frag_mapq ? (int)__ldcs(frag_mapq + lo_al + i) : 255,
                  frag_strand ? (int)__ldcs(frag_strand + lo_al + i) : 1);
    }
    if (SMEM) {
        __syncthreads();
        for (int b = tid; b < n_bins; b += kMotifThreads) {
            const int c = s_cnt[b];
            if (c) atomicAdd(&row[b], (unsigned long long)c);
        }
    }
}

}  // namespace ftk

using namespace ftk;

template <bool BREAKPOINT>
static int motif_hist_launch(const int32_t *frag_start, const int32_t *frag_stop,
                             const uint8_t *frag_mapq, const uint8_t *frag_strand,
                             int64_t n_frag, int32_t max_frag_len,
                             const uint32_t *seq_words, const uint32_t *nmask_words, int64_t contig_len,
                             const int32_t *ivl_start, const int32_t *ivl_stop, int64_t n_ivl,
                             int32_t k, int32_t strand_mode, int32_t min_mapq,
                             int32_t pooled, int32_t splits,
                             int64_t *scratch, uint64_t *counts, int32_t *error_flag,
                             ftk_stream_t stream_) {
    if (n_ivl == 0) return FTK_OK;
    if (n_frag < 0 || n_ivl < 0 || splits < 1 || k < 1 || k > 12) return FTK_E_INVALID;
    if (strand_mode < 0 || strand_mode > 2) return FTK_E_INVALID;
    if (!seq_words || !nmask_words || !ivl_start || !ivl_stop || !scratch || !counts) return FTK_E_INVALID;
    if (!BREAKPOINT && !error_flag) return FTK_E_INVALID;
    if (n_frag > 0 && (!frag_start || !frag_stop)) return FTK_E_INVALID;
    if (n_ivl * (int64_t)splits > INT32_MAX) return FTK_E_RANGE;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    {
        const int64_t n = 2 * n_ivl;
        motif_ranges_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(
            frag_start, n_frag, ivl_start, ivl_stop, n_ivl, max_frag_len < 0 ? 0 : max_frag_len, scratch);
        FTK_CHECK_LAUNCH("motif_ranges_kernel");
    }
    const unsigned grid = (unsigned)(n_ivl * splits);
    auto *c = reinterpret_cast<unsigned long long *>(counts);
    if ((1 << (2 * k)) <= kMotifSmemBins)
        end_motif_kernel<true, BREAKPOINT><<<grid, kMotifThreads, 0, stream>>>(
            frag_start, frag_stop, frag_mapq, frag_strand, seq_words, nmask_words, contig_len,
            ivl_start, ivl_stop, scratch, k, strand_mode, min_mapq, pooled, splits, c, error_flag);
    else
        end_motif_kernel<false, BREAKPOINT><<<grid, kMotifThreads, 0, stream>>>(
            frag_start, frag_stop, frag_mapq, frag_strand, seq_words, nmask_words, contig_len,
            ivl_start, ivl_stop, scratch, k, strand_mode, min_mapq, pooled, splits, c, error_flag);
    FTK_CHECK_LAUNCH("end_motif_kernel");
    return FTK_OK;
}

extern "C" int ftk_end_motif_hist_u64(const int32_t *frag_start, const int32_t *frag_stop,
                                      const uint8_t *frag_mapq, const uint8_t *frag_strand,
                                      int64_t n_frag, int32_t max_frag_len,
                                      const uint32_t *seq_words, const uint32_t *nmask_words,
                                      int64_t contig_len,
                                      const int32_t *ivl_start, const int32_t *ivl_stop, int64_t n_ivl,
                                      int32_t k, int32_t strand_mode, int32_t min_mapq,
                                      int32_t pooled, int32_t splits,
                                      int64_t *scratch, uint64_t *counts, int32_t *error_flag,
                                      ftk_stream_t stream_) {
    return motif_hist_launch<false>(frag_start, frag_stop, frag_mapq, frag_strand, n_frag, max_frag_len,
                                    seq_words, nmask_words, contig_len, ivl_start, ivl_stop, n_ivl, k,
                                    strand_mode, min_mapq, pooled, splits, scratch, counts, error_flag, stream_);
}

extern "C" int ftk_breakpoint_motif_hist_u64(const int32_t *frag_start, const int32_t *frag_stop,
                                             const uint8_t *frag_mapq, const uint8_t *frag_strand,
                                             int64_t n_frag, int32_t max_frag_len,
                                             const uint32_t *seq_words, const uint32_t *nmask_words,
                                             int64_t contig_len,
                                             const int32_t *ivl_start, const int32_t *ivl_stop, int64_t n_ivl,
                                             int32_t k, int32_t strand_mode, int32_t min_mapq,
                                             int32_t pooled, int32_t splits,
                                             int64_t *scratch, uint64_t *counts, ftk_stream_t stream_) {
    return motif_hist_launch<true>(frag_start, frag_stop, frag_mapq, frag_strand, n_frag, max_frag_len,
                                   seq_words, nmask_words, contig_len, ivl_start, ivl_stop, n_ivl, k,
                                   strand_mode, min_mapq, pooled, splits, scratch, counts, nullptr, stream_);
}
