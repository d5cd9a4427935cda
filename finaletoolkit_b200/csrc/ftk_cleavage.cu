// Cleavage profile (Zhou et al. 2022) on sm_100a: per-base fragment depth + fragment-end counts.
//
// Replaces _coverage_and_ends (frag/_cleavage_profile.py:33-90: np.add.at difference array +
// cumsum for depth, bincount of strand-selected ends) and the proportion step of
// cleavage_profile (frag/_cleavage_profile.py:190-217):
//   depth[p]  = #{fragments with start <= p < stop}       (start/stop clipped to the interval)
//   ends[p]   = #{'+' fragments with start == p} + #{'-' fragments with stop == p}
//   out[p]    = depth ? ends / depth * 100 : 0             (fp64, same operation order)
// over the fragments of frag_array(..., intersect_policy="any") for the interval: mapq >= q,
// inclusive length window, stop > interval_start and start < interval_stop.
// Same tile machinery as WPS: one CTA per <= 5120-position tile, events scattered with shared
// atomics into a difference array (+1 at start, -1 at stop, clamped onto the first slot /
// dropped past the last) and an end-count array, block prefix scan, fp64 store.
// Roofline: HBM, 10 B per candidate fragment (start, stop, mapq, strand) + 8 B per position.
// Measured and dropped in round 2: ONE packed array (ends * 65536 + depth difference in a word: two
// atomics per fragment instead of three, 20 KB instead of 40 KB per tile, WPS-style shuffle scan) -
// 905 us with direct 16-byte stores, 1.08 ms with the results staged through the array, against
// 827 us for the two-array form below; ATOMS.ADD vs ATOMS.POPC.INC makes no difference (834 / 839 us).
#include "ftk_common.cuh"

namespace ftk {

constexpr int kClvThreads = 256;
constexpr int kClvCap = 5120;
constexpr int kClvPer = kClvCap / kClvThreads;   // 20 contiguous positions per thread

__global__ void clv_tile_ranges_kernel(const int32_t *__restrict__ frag_start, int64_t n_frag,
                                       const int32_t *__restrict__ tile_p0, const int32_t *__restrict__ tile_len,
                                       int64_t n_tiles, int halo, int64_t *__restrict__ ranges) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * n_tiles) return;
    const int64_t tile = t >> 1;
    const int64_t p0 = tile_p0[tile];
    // a covering fragment has stop > p0 (=> start > p0 - max_frag_len) and start < p0 + len
    const int64_t key = (t & 1) ? p0 + tile_len[tile] : p0 - halo;
    ranges[t] = lower_bound(frag_start, n_frag, key);
}

__global__ void __launch_bounds__(kClvThreads)
cleavage_tile_kernel(const int32_t *__restrict__ frag_start, const int32_t *__restrict__ frag_stop,
                     const uint8_t *__restrict__ frag_mapq, const uint8_t *__restrict__ frag_strand,
                     const int32_t *__restrict__ tile_p0, const int32_t *__restrict__ tile_len,
                     const int32_t *__restrict__ tile_ivl_lo, const int32_t *__restrict__ tile_ivl_hi,
                     const int64_t *__restrict__ tile_out_off, const int64_t *__restrict__ ranges,
                     int min_len, int max_len, int min_mapq, double *__restrict__ out) {
    __shared__ __align__(16) int DE[2 * kClvCap];
    int *D = DE;              // depth difference array
    int *E = DE + kClvCap;    // fragment-end counts (the pair is recycled as double[kClvCap] for the store)
    __shared__ int warp_tot[kClvThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t tile = blockIdx.x;
    const int p0 = tile_p0[tile], len = tile_len[tile];
    const int ivl_lo = tile_ivl_lo[tile], ivl_hi = tile_ivl_hi[tile];
    const int64_t lo = ranges[2 * tile], hi = ranges[2 * tile + 1];
    for (int i = tid; i < kClvCap / 4; i += kClvThreads) {
        reinterpret_cast<int4 *>(D)[i] = make_int4(0, 0, 0, 0);
        reinterpret_cast<int4 *>(E)[i] = make_int4(0, 0, 0, 0);
    }
    __syncthreads();
    const uint32_t d_addr = smem_addr_once(D), e_addr = smem_addr_once(E);
    auto visit = [&](int fs, int fe, int q, int sd) {
        const int L = fe - fs;
        // frag_array(..., "any") of the INTERVAL: tabix overlap + mapq + inclusive length window
        if (q < min_mapq || L < 0 || !frag_len_ok(L, min_len, max_len) || !(fe > ivl_lo && fs < ivl_hi)) return;
        const int s_idx = fs - p0, e_idx = fe - p0;
        if (e_idx > 0 && s_idx < len) {                 // covers at least one tile position
            red_shared_inc(d_addr + 4u * (unsigned)max(s_idx, 0));
            if (e_idx < len) red_shared_add(d_addr + 4u * (unsigned)e_idx, -1);
        }
        const int end_idx = sd ? s_idx : e_idx;
        if (end_idx >= 0 && end_idx < len) red_shared_inc(e_addr + 4u * (unsigned)end_idx);
    };
    // the candidate slice is widened to a 16-byte boundary on the left (extra fragments are harmless:
    // the range is only a superset) so every lane streams 4 fragments per 128-bit load
    constexpr int kU = 2;   // independent vector loads in flight per thread (8 fragments)
    const int64_t lo_al = lo & ~(int64_t)3;
    const int cnt = (hi > lo_al) ? (int)(hi - lo_al) : 0;
    const int nvec = cnt >> 2;
    const int4 *__restrict__ vs = reinterpret_cast<const int4 *>(frag_start + lo_al);
    const int4 *__restrict__ ve = reinterpret_cast<const int4 *>(frag_stop + lo_al);
    const uchar4 *__restrict__ vq = reinterpret_cast<const uchar4 *>(frag_mapq ? frag_mapq + lo_al : nullptr);
    const uchar4 *__restrict__ vd = reinterpret_cast<const uchar4 *>(frag_strand ? frag_strand + lo_al : nullptr);
    for (int v0 = tid; v0 < nvec; v0 += kU * kClvThreads) {
        int4 s4[kU], e4[kU];
        uchar4 q4[kU], d4[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int v = v0 + u * kClvThreads;
            if (v < nvec) {
                s4[u] = __ldcs(vs + v); e4[u] = __ldcs(ve + v);
                q4[u] = vq ? __ldcs(vq + v) : make_uchar4(255, 255, 255, 255);
                d4[u] = vd ? __ldcs(vd + v) : make_uchar4(1, 1, 1, 1);
            } else {
                s4[u] = make_int4(0, 0, 0, 0); e4[u] = make_int4(-1, -1, -1, -1);   // L < 0: skipped
                q4[u] = make_uchar4(0, 0, 0, 0); d4[u] = make_uchar4(1, 1, 1, 1);
            }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            visit(s4[u].x, e4[u].x, q4[u].x, d4[u].x);
            visit(s4[u].y, e4[u].y, q4[u].y, d4[u].y);
            visit(s4[u].z, e4[u].z, q4[u].z, d4[u].z);
            visit(s4[u].w, e4[u].w, q4[u].w, d4[u].w);
        }
    }
    {   // tail: at most 3 fragments
        const int i = nvec * 4 + tid;
        if (i < cnt)
            visit(__ldcs(frag_start + lo_al + i), __ldcs(frag_stop + lo_al + i),
                  frag_mapq ? (int)__ldcs(frag_mapq + lo_al + i) : 255,
                  frag_strand ? (int)__ldcs(frag_strand + lo_al + i) : 1);
    }
    __syncthreads();
    // block prefix scan: each thread owns kClvPer contiguous positions (int4 loads, conflict-free)
    int d[kClvPer];
    const int base = tid * kClvPer;
#pragma unroll
    for (int j = 0; j < kClvPer / 4; ++j) {
        const int4 v = *reinterpret_cast<const int4 *>(&D[base + 4 * j]);
        d[4 * j] = v.x; d[4 * j + 1] = v.y; d[4 * j + 2] = v.z; d[4 * j + 3] = v.w;
    }
#pragma unroll
    for (int j = 1; j < kClvPer; ++j) d[j] += d[j - 1];
    int t = d[kClvPer - 1];
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, t, off);
        if (lane >= off) t += n;
    }
    if (lane == 31) warp_tot[warp] = t;
    __syncthreads();
    int offset = t - d[kClvPer - 1];
#pragma unroll
    for (int w = 0; w < kClvThreads / 32; ++w) offset += (w < warp) ? warp_tot[w] : 0;
    int e[kClvPer];
#pragma unroll
    for (int j = 0; j < kClvPer / 4; ++j) {
        const int4 v = *reinterpret_cast<const int4 *>(&E[base + 4 * j]);
        e[4 * j] = v.x; e[4 * j + 1] = v.y; e[4 * j + 2] = v.z; e[4 * j + 3] = v.w;
    }
    __syncthreads();   // every thread holds its depth/end counts: D and E can be recycled
    double *stage = reinterpret_cast<double *>(DE);
#pragma unroll
    for (int j = 0; j < kClvPer; ++j) {
        const int depth = d[j] + offset;
        // proportions[mask] = ends[mask] / depth[mask] * 100   (frag/_cleavage_profile.py:206-208);
        // 0 / depth * 100 is exactly 0.0, so the fp64 division only runs where an end was counted
        stage[base + j] = (depth != 0 && e[j] != 0) ? (double)e[j] / (double)depth * 100.0 : 0.0;
    }
    __syncthreads();
    double *__restrict__ dst = out + tile_out_off[tile];
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {               // coalesced 16-byte streaming stores
        const int pairs = len >> 1;
        for (int p = tid; p < pairs; p += kClvThreads) {
            const double2 v = reinterpret_cast<const double2 *>(stage)[p];
            asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" :: "l"(dst + 2 * p), "d"(v.x), "d"(v.y) : "memory");
        }
        if ((len & 1) && tid == 0) dst[len - 1] = stage[len - 1];
    } else {
        for (int p = tid; p < len; p += kClvThreads) dst[p] = stage[p];
    }
}

}  // namespace ftk

using namespace ftk;

extern "C" int ftk_cleavage_tiles_f64(const int32_t *frag_start, const int32_t *frag_stop,
                                      const uint8_t *frag_mapq, const uint8_t *frag_strand,
                                      int64_t n_frag, int32_t max_frag_len,
                                      const int32_t *tile_p0, const int32_t *tile_len,
                                      const int32_t *tile_ivl_lo, const int32_t *tile_ivl_hi,
                                      const int64_t *tile_out_off, int64_t n_tiles,
                                      int32_t min_len, int32_t max_len, int32_t min_mapq,
                                      int64_t *scratch, double *out, ftk_stream_t stream_) {
    if (n_tiles == 0) return FTK_OK;
    if (n_frag < 0 || n_tiles < 0) return FTK_E_INVALID;
    if (!tile_p0 || !tile_len || !tile_ivl_lo || !tile_ivl_hi || !tile_out_off || !scratch || !out) return FTK_E_INVALID;
    if (n_frag > 0 && (!frag_start || !frag_stop)) return FTK_E_INVALID;
    if (n_tiles > INT32_MAX / 2) return FTK_E_RANGE;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int64_t n = 2 * n_tiles;
    clv_tile_ranges_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(
        frag_start, n_frag, tile_p0, tile_len, n_tiles, max_frag_len < 0 ? 0 : max_frag_len, scratch);
    FTK_CHECK_LAUNCH("clv_tile_ranges_kernel");
    cleavage_tile_kernel<<<(unsigned)n_tiles, kClvThreads, 0, stream>>>(
        frag_start, frag_stop, frag_mapq, frag_strand, tile_p0, tile_len, tile_ivl_lo, tile_ivl_hi, tile_out_off,
        scratch, min_len, max_len, min_mapq, out);
    FTK_CHECK_LAUNCH("cleavage_tile_kernel");
    return FTK_OK;
}
