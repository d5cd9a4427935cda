// Strand-aware aggregation of a per-base signal over equal-length intervals on sm_100a.
//
// Replaces the accumulation loop of agg_bw (utils/_agg_bw.py:84-123): for every interval the
// reference trims the signal, flips it for '-' strand intervals and adds it to a running fp64
// vector.  Here the trimmed signals of all accepted intervals form one row-major float32 matrix
// in HBM and each thread owns one output position, adding the intervals in file order - the same
// order of fp64 additions as the reference, so the result is bit-identical for any signal (not
// only integer WPS).  NaN (uncovered bases) counts as 0 (np.nan_to_num, :98).
// Roofline: HBM, 4 B per (interval, position) sample read once.
#include "ftk_common.cuh"

namespace ftk {

constexpr int kAggThreads = 128;
constexpr int kAggUnroll = 8;

__global__ void __launch_bounds__(kAggThreads)
agg_signal_kernel(const float *__restrict__ x, int64_t n_seg, int64_t row_len, int trim_lo, int out_len,
                  const int8_t *__restrict__ strand, double *__restrict__ out) {
    const int p = blockIdx.x * kAggThreads + threadIdx.x;
    if (p >= out_len) return;
    const int64_t fwd = trim_lo + p, rev = trim_lo + (out_len - 1 - p);
    double acc = 0.0;
    int64_t s = 0;
    for (; s + kAggUnroll <= n_seg; s += kAggUnroll) {
        float v[kAggUnroll];
#pragma unroll
        for (int u = 0; u < kAggUnroll; ++u) {
            const int sd = strand[s + u];
            v[u] = sd ? __ldcs(x + (s + u) * row_len + (sd > 0 ? fwd : rev)) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < kAggUnroll; ++u)
            if (strand[s + u]) acc += (v[u] != v[u]) ? 0.0 : (double)v[u];
    }
    for (; s < n_seg; ++s) {
        const int sd = strand[s];
        if (!sd) continue;
        const float v = __ldcs(x + s * row_len + (sd > 0 ? fwd : rev));
        acc += (v != v) ? 0.0 : (double)v;
    }
    out[p] = acc;
}

}  // namespace ftk

using namespace ftk;

extern "C" int ftk_agg_signal_f64(const float *signal, int64_t n_seg, int64_t row_len, int32_t trim_lo,
                                  int32_t out_len, const int8_t *strand, double *out, ftk_stream_t stream_) {
    if (out_len == 0) return FTK_OK;
    if (n_seg < 0 || row_len < 0 || trim_lo < 0 || out_len < 0 || (int64_t)trim_lo + out_len > row_len)
        return FTK_E_INVALID;
    if (!out || (n_seg > 0 && (!signal || !strand))) return FTK_E_INVALID;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    agg_signal_kernel<<<(unsigned)((out_len + kAggThreads - 1) / kAggThreads), kAggThreads, 0, stream>>>(
        signal, n_seg, row_len, trim_lo, out_len, strand, out);
    FTK_CHECK_LAUNCH("agg_signal_kernel");
    return FTK_OK;
}
