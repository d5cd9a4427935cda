// Windowed Protection Score on sm_100a.
//
// Replaces the reference's O(len x n_frag) per-position loop
// (frag/_wps.py:176-188 -> _single_nt_wps, frag/_wps.py:25-53) with a
// fragment-event scatter into a shared-memory difference array followed by a
// block prefix scan.  For window [c-a, c+b] (even W: a=W/2, b=W/2-1; odd W:
// a=b=(W-1)/2 around c' = c - ((c-a)&1), which is what np.rint's
// round-half-even does at frag/_wps.py:177-178) a fragment (fs, fe), L=fe-fs:
//   L <= W : contributes -1 on [fs-b, fe+a]               (either end inside)
//   L >  W : -1 on [fs-b, fs+a], +1 on [fs+a+1, fe-b-1], -1 on [fe-b, fe+a]
// i.e. D[fs-b]-=1, D[fe+a+1]+=1 and, iff L>W, D[fs+a+1]+=2, D[fe-b]-=2;
// WPS = inclusive prefix sum of D.  Events left of the tile are clamped onto
// its first slot (they belong to the prefix), events right of it are dropped,
// so tiles are independent: no inter-CTA carry, no collective.
//
// Roofline: HBM.  Algorithmic bytes = 9 B per fragment (start, stop int32 +
// mapq uint8) + 4 B per output position.
#include "ftk_common.cuh"

namespace ftk {

constexpr int kWpsThreads = 256;
constexpr int kWpsWarps = kWpsThreads / 32;
constexpr int kWpsCap = 5120;                    // smem slots per tile
constexpr int kWpsIters = kWpsCap / (kWpsThreads * 4);  // int4 groups per lane
constexpr int kWpsSpan = kWpsIters * 128;        // positions per warp
constexpr int kWpsUnroll = 4;                    // fragment loads in flight per thread (x3 columns)
static_assert(kWpsIters * kWpsThreads * 4 == kWpsCap, "tile must split evenly");
static_assert(FTK_WPS_TILE < kWpsCap, "one guard slot for odd windows");

// Per-tile fragment index range [lo, hi): a superset of the fragments with any
// event inside the tile.  One thread per bound; the upper levels of the search
// tree stay in L2, so this costs a few microseconds per launch.
__global__ void wps_tile_ranges_kernel(const int32_t *__restrict__ frag_start, int64_t n_frag,
                                       const int32_t *__restrict__ tile_p0,
                                       const int32_t *__restrict__ tile_len, int64_t n_tiles,
                                       int a, int b, int max_len, int64_t *__restrict__ ranges) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * n_tiles) return;
    int64_t tile = t >> 1;
    int64_t p0 = tile_p0[tile];
    int64_t key;
    if ((t & 1) == 0) {
        // last event of a fragment sits at fe+a+1 <= fs+max_len+a+1; it must be > g0 >= p0-1
        key = p0 - 1 - (int64_t)a - (int64_t)max_len;
    } else {
        // first event sits at fs-b; it must be < p0+len
        key = p0 + (int64_t)tile_len[tile] + (int64_t)b;
    }
    ranges[t] = lower_bound(frag_start, n_frag, key);
}

template <bool ODD>
__global__ void __launch_bounds__(kWpsThreads)
wps_tile_kernel(const int32_t *__restrict__ frag_start, const int32_t *__restrict__ frag_stop,
                const uint8_t *__restrict__ frag_mapq,
                const int32_t *__restrict__ tile_p0, const int32_t *__restrict__ tile_len,
                const int32_t *__restrict__ tile_mid_lo, const int32_t *__restrict__ tile_mid_hi,
                const int64_t *__restrict__ tile_out_off, const int64_t *__restrict__ ranges,
                int window, int a, int b, int min_len, int max_len, int min_mapq,
                int32_t *__restrict__ out) {
    __shared__ __align__(16) int D[kWpsCap];
    __shared__ int warp_tot[kWpsWarps];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int64_t tile = blockIdx.x;

    const int p0 = tile_p0[tile];
    const int len = tile_len[tile];
    const int mid_lo = tile_mid_lo[tile];
    const int mid_hi = tile_mid_hi[tile];
    const int64_t lo = ranges[2 * tile];
    const int64_t hi = ranges[2 * tile + 1];
    // grid of the symmetric-window score G: one extra slot on the left for odd W
    const int g0 = ODD ? p0 - 1 : p0;
    const int glen = ODD ? len + 1 : len;

    // One fragment's events -> shared atomics (2, or 4 when L > W).
    auto scatter = [&](int fs, int fe, int q) {
        const int L = fe - fs;
        const int mid = fs + (L >> 1);  // (fs+fe)//2 for L >= 0
        const bool pass = (q >= min_mapq) && (L >= 0) && frag_len_ok(L, min_len, max_len) &&
                          (mid >= mid_lo) && (mid < mid_hi);
        if (!pass) return;
        const int e_last = fe + a + 1 - g0;  // +1
        const int e_first = fs - b - g0;     // -1
        if (e_last <= 0 || e_first >= glen) return;  // cancels on slot 0 / entirely right of the tile
        atomicAdd(&D[max(e_first, 0)], -1);
        if (e_last < glen) atomicAdd(&D[e_last], 1);
        if (L > window) {
            const int e1 = fs + a + 1 - g0;  // +2
            const int e2 = fe - b - g0;      // -2
            if (e1 < glen) atomicAdd(&D[max(e1, 0)], 2);
            if (e2 < glen) atomicAdd(&D[max(e2, 0)], -2);
        }
    };

    // ---- software-pipelined scatter: kWpsUnroll x 3 coalesced streaming loads stay in
    // flight per thread while the previous batch is turned into shared atomics; the
    // first batch is issued before the tile is zeroed so its latency hides behind that.
    int fs_r[kWpsUnroll], fe_r[kWpsUnroll], q_r[kWpsUnroll];
    auto load_batch = [&](int64_t i0) {
#pragma unroll
        for (int u = 0; u < kWpsUnroll; ++u) {
            const int64_t i = i0 + (int64_t)u * kWpsThreads;
            const bool in = i < hi;
            fs_r[u] = in ? __ldcs(frag_start + i) : 0;
            fe_r[u] = in ? __ldcs(frag_stop + i) : 0;
            q_r[u] = in ? (frag_mapq ? (int)__ldcs(frag_mapq + i) : 255) : -1;  // -1 never passes
        }
    };
    int64_t i0 = lo + tid;
    load_batch(i0);

#pragma unroll
    for (int j = 0; j < kWpsIters; ++j)
        reinterpret_cast<int4 *>(D)[j * kWpsThreads + tid] = make_int4(0, 0, 0, 0);
    __syncthreads();

    while (i0 < hi) {  // block-uniform trip count not required: no barrier inside
        int fs_c[kWpsUnroll], fe_c[kWpsUnroll], q_c[kWpsUnroll];
#pragma unroll
        for (int u = 0; u < kWpsUnroll; ++u) { fs_c[u] = fs_r[u]; fe_c[u] = fe_r[u]; q_c[u] = q_r[u]; }
        i0 += (int64_t)kWpsUnroll * kWpsThreads;
        if (i0 < hi) load_batch(i0);
#pragma unroll
        for (int u = 0; u < kWpsUnroll; ++u) scatter(fs_c[u], fe_c[u], q_c[u]);
    }
    __syncthreads();

    // ---- block prefix scan: lane owns int4 groups, warp owns a contiguous span
    int4 v[kWpsIters];
    int carry = 0;
    const int span0 = warp * kWpsSpan;
#pragma unroll
    for (int j = 0; j < kWpsIters; ++j) {
        const int base = span0 + j * 128 + lane * 4;
        int4 d = *reinterpret_cast<const int4 *>(&D[base]);
        d.y += d.x; d.z += d.y; d.w += d.z;
        int t = d.w;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            int n = __shfl_up_sync(0xffffffffu, t, off);
            if (lane >= off) t += n;
        }
        const int excl = t - d.w + carry;
        v[j] = make_int4(d.x + excl, d.y + excl, d.z + excl, d.w + excl);
        carry += __shfl_sync(0xffffffffu, t, 31);
    }
    if (lane == 0) warp_tot[warp] = carry;
    __syncthreads();
    int offset = 0;
#pragma unroll
    for (int w = 0; w < kWpsWarps; ++w) offset += (w < warp) ? warp_tot[w] : 0;

    int32_t *__restrict__ dst = out + tile_out_off[tile];
    if (!ODD) {
        const bool aligned = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
#pragma unroll
        for (int j = 0; j < kWpsIters; ++j) {
            const int base = span0 + j * 128 + lane * 4;
            const int4 r = make_int4(v[j].x + offset, v[j].y + offset, v[j].z + offset, v[j].w + offset);
            if (aligned && base + 3 < len) {
                st_stream4(reinterpret_cast<int4 *>(dst + base), r);
            } else {
                if (base + 0 < len) dst[base + 0] = r.x;
                if (base + 1 < len) dst[base + 1] = r.y;
                if (base + 2 < len) dst[base + 2] = r.z;
                if (base + 3 < len) dst[base + 3] = r.w;
            }
        }
    } else {
        // odd W: out[c] = G[c - ((c - a) & 1)]; stage G in shared memory (own span only)
#pragma unroll
        for (int j = 0; j < kWpsIters; ++j) {
            const int base = span0 + j * 128 + lane * 4;
            *reinterpret_cast<int4 *>(&D[base]) =
                make_int4(v[j].x + offset, v[j].y + offset, v[j].z + offset, v[j].w + offset);
        }
        __syncthreads();
        for (int k = tid; k < len; k += kWpsThreads) {
            const int c = p0 + k;
            dst[k] = D[c - ((c - a) & 1) - g0];
        }
    }
}

}  // namespace ftk

extern "C" int64_t ftk_wps_plan_tiles(const int64_t *ivl_start, const int64_t *ivl_stop,
                                      const int64_t *ivl_out_off, int64_t n_ivl,
                                      int64_t chrom_size, int32_t max_len,
                                      int32_t *tile_p0, int32_t *tile_len,
                                      int32_t *tile_mid_lo, int32_t *tile_mid_hi,
                                      int64_t *tile_out_off) {
    if (n_ivl < 0 || (n_ivl > 0 && (!ivl_start || !ivl_stop || !ivl_out_off))) return FTK_E_INVALID;
    if (chrom_size < 0 || chrom_size > INT32_MAX || max_len < 0) return FTK_E_RANGE;
    constexpr int64_t cap4 = FTK_WPS_TILE & ~3;  // keep int4-aligned pieces
    int64_t n_tiles = 0;
    for (int64_t k = 0; k < n_ivl; ++k) {
        const int64_t S = ivl_start[k], E = ivl_stop[k];
        if (E <= S) continue;  // degenerate interval: no output (frag/_wps.py:145-152)
        if (S < INT32_MIN / 2 || E > INT32_MAX) return FTK_E_RANGE;
        // frag/_wps.py:156-157 : padded fetch window = midpoint predicate of this interval
        int64_t mlo = S - max_len; if (mlo < 0) mlo = 0;
        int64_t mhi = E + max_len; if (mhi > chrom_size) mhi = chrom_size;
        const int64_t len = E - S;
        const int64_t pieces = (len + cap4 - 1) / cap4;
        int64_t piece = (len + pieces - 1) / pieces;
        piece = (piece + 3) & ~(int64_t)3;
        for (int64_t s = 0; s < len; s += piece) {
            if (tile_p0) {
                tile_p0[n_tiles] = (int32_t)(S + s);
                tile_len[n_tiles] = (int32_t)((len - s < piece) ? (len - s) : piece);
                tile_mid_lo[n_tiles] = (int32_t)mlo;
                tile_mid_hi[n_tiles] = (int32_t)mhi;
                tile_out_off[n_tiles] = ivl_out_off[k] + s;
            }
            ++n_tiles;
        }
    }
    return n_tiles;
}

static int wps_check_args(const int32_t *frag_start, const int32_t *frag_stop, int64_t n_frag,
                          const int32_t *tile_p0, const int32_t *tile_len, int64_t n_tiles,
                          int32_t window_size, int32_t max_len, const int64_t *scratch) {
    if (n_frag < 0 || n_tiles < 0 || window_size < 1) return FTK_E_INVALID;
    if (!tile_p0 || !tile_len || !scratch) return FTK_E_INVALID;
    if (n_frag > 0 && (!frag_start || !frag_stop)) return FTK_E_INVALID;
    // the reference requires an integer max_length (frag/_wps.py:156 round(start - max_length))
    if (max_len == FTK_NONE || max_len < 0) return FTK_E_INVALID;
    if (n_tiles > INT32_MAX / 2) return FTK_E_RANGE;
    return FTK_OK;
}

extern "C" int ftk_wps_tile_ranges(const int32_t *frag_start, int64_t n_frag,
                                   const int32_t *tile_p0, const int32_t *tile_len, int64_t n_tiles,
                                   int32_t window_size, int32_t max_len, int64_t *scratch,
                                   ftk_stream_t stream_) {
    using namespace ftk;
    if (n_tiles == 0) return FTK_OK;
    int rc = wps_check_args(frag_start, frag_start, n_frag, tile_p0, tile_len, n_tiles, window_size, max_len, scratch);
    if (rc != FTK_OK) return rc;
    const bool odd = (window_size & 1) != 0;
    const int a = odd ? (window_size - 1) / 2 : window_size / 2;
    const int b = odd ? a : a - 1;
    const int64_t n = 2 * n_tiles;
    const int threads = 128;
    wps_tile_ranges_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0,
                             static_cast<cudaStream_t>(stream_)>>>(
        frag_start, n_frag, tile_p0, tile_len, n_tiles, a, b, max_len, scratch);
    FTK_CHECK_LAUNCH("wps_tile_ranges_kernel");
    return FTK_OK;
}

extern "C" int ftk_wps_tiles_i32(const int32_t *frag_start, const int32_t *frag_stop,
                                 const uint8_t *frag_mapq, int64_t n_frag,
                                 const int32_t *tile_p0, const int32_t *tile_len,
                                 const int32_t *tile_mid_lo, const int32_t *tile_mid_hi,
                                 const int64_t *tile_out_off, int64_t n_tiles,
                                 int32_t window_size, int32_t min_len, int32_t max_len, int32_t min_mapq,
                                 int32_t ranges_ready, int64_t *scratch, int32_t *out,
                                 ftk_stream_t stream_) {
    using namespace ftk;
    if (n_tiles == 0) return FTK_OK;
    int rc = wps_check_args(frag_start, frag_stop, n_frag, tile_p0, tile_len, n_tiles, window_size, max_len, scratch);
    if (rc != FTK_OK) return rc;
    if (!tile_mid_lo || !tile_mid_hi || !tile_out_off || !out) return FTK_E_INVALID;
    if (!ranges_ready) {
        rc = ftk_wps_tile_ranges(frag_start, n_frag, tile_p0, tile_len, n_tiles, window_size, max_len, scratch, stream_);
        if (rc != FTK_OK) return rc;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const bool odd = (window_size & 1) != 0;
    const int a = odd ? (window_size - 1) / 2 : window_size / 2;
    const int b = odd ? a : a - 1;
    if (odd)
        wps_tile_kernel<true><<<(unsigned)n_tiles, kWpsThreads, 0, stream>>>(
            frag_start, frag_stop, frag_mapq, tile_p0, tile_len, tile_mid_lo, tile_mid_hi, tile_out_off,
            scratch, window_size, a, b, min_len, max_len, min_mapq, out);
    else
        wps_tile_kernel<false><<<(unsigned)n_tiles, kWpsThreads, 0, stream>>>(
            frag_start, frag_stop, frag_mapq, tile_p0, tile_len, tile_mid_lo, tile_mid_hi, tile_out_off,
            scratch, window_size, a, b, min_len, max_len, min_mapq, out);
    FTK_CHECK_LAUNCH("wps_tile_kernel");
    return FTK_OK;
}
