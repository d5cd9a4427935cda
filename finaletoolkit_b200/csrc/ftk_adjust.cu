// adjust_wps on sm_100a: running median / mean subtraction + Savitzky-Golay smoothing.
//
// Replaces _local_filter / _running_stat (frag/_adjust_wps.py:25-45: numpy
// sliding_window_view + np.median / np.mean per window, O(n w log w)) and the
// scipy.signal.savgol_filter call (frag/_adjust_wps.py:135-138, mode='interp') of
// _single_adjust_wps (frag/_adjust_wps.py:63-163).
//
//   adj[j] = s[j + w/2] - stat(s[j : j+w]),  j in [0, n-w),  s = x - shift
//   out    = savgol(adj)  : interior = 21-tap (sg_w) fp64 stencil, first/last sg_w/2
//            outputs = polynomial fit of the first/last sg_w samples (edge matrices)
//
// Raw WPS is integer-valued with a small local range, so the running median is a
// sliding HISTOGRAM median (add one sample, drop one, walk the median bin): O(1) per
// output instead of a sort per window.  Work unit = "run": kAdjRun consecutive outputs
// of one segment, owned by ONE thread that slides its private shared-memory histogram
// (int16 bins, bank-conflict-free layout) along the run; the run also carries its own
// Savitzky-Golay ring buffer, so median, subtraction and smoothing are fused and the
// adjusted series never round-trips through HBM.  Runs whose samples are not integers
// or leave the kAdjBins-wide local window are flagged and redone by the generic
// kernel (sorted-window insertion, any float input).
// fp64 arithmetic follows numpy's: median of an even window = (lo + hi) / 2 on the
// shifted values; mean = exact integer sum / w.
// Roofline: HBM, 4 B in (float32 sample) + 8 B out (float64) per position.
#include "ftk_common.cuh"

namespace ftk {

constexpr int kAdjThreads = 128;
constexpr int kAdjBins = 256;       // local value window of the fast path
constexpr int kAdjMaxSg = 127;      // Savitzky-Golay window limit (ring buffer)
constexpr int kAdjRing = 128;

struct AdjParams {
    int w;            // median / mean window (even)
    int use_mean;
    int savgol;       // 0: out = adj
    int sg_w;         // odd, <= kAdjMaxSg
    int run;          // outputs per run
};

// Emits SG outputs for the run [ja, jb) while adj values arrive in order.
struct SgEmitter {
    const double *coef;        // [sg_w] interior stencil (correlation order)
    const double *edge_first;  // [half][sg_w]
    const double *edge_last;   // [half][sg_w]
    double *out;               // segment output base
    double ring[kAdjRing];
    int sg_w, half;
    long long ja, jb, n_out;

    __device__ __forceinline__ void push(long long c, double adj) {
        ring[c & (kAdjRing - 1)] = adj;
        // interior output j = c - half
        const long long j = c - half;
        if (j >= ja && j < jb && j >= half && j < n_out - half) {
            double acc = 0.0;
            for (int i = 0; i < sg_w; ++i) acc += coef[i] * ring[(j - half + i) & (kAdjRing - 1)];
            out[j] = acc;
        }
        if (c == sg_w - 1) {  // first sg_w samples complete: left edge outputs
            const long long hi = jb < half ? jb : half;
            for (long long e = ja; e < hi; ++e) {
                double acc = 0.0;
                for (int i = 0; i < sg_w; ++i) acc += edge_first[e * sg_w + i] * ring[i & (kAdjRing - 1)];
                out[e] = acc;
            }
        }
        if (c == n_out - 1) {  // last sg_w samples complete: right edge outputs
            const long long lo = ja > n_out - half ? ja : n_out - half;
            for (long long e = lo; e < jb; ++e) {
                double acc = 0.0;
                for (int i = 0; i < sg_w; ++i)
                    acc += edge_last[(e - (n_out - half)) * sg_w + i] * ring[(n_out - sg_w + i) & (kAdjRing - 1)];
                out[e] = acc;
            }
        }
    }
};

// which adj indices [ca, cb) a run must compute to emit outputs [ja, jb)
__device__ __forceinline__ void run_extent(const AdjParams &P, long long ja, long long jb, long long n_out,
                                           long long &ca, long long &cb) {
    if (!P.savgol) { ca = ja; cb = jb; return; }
    const int half = P.sg_w / 2;
    ca = ja - half; cb = jb + half;
    if (ja < half) { ca = 0; if (cb < P.sg_w) cb = P.sg_w; }
    if (jb > n_out - half) { if (ca > n_out - P.sg_w) ca = n_out - P.sg_w; cb = n_out; }
    if (ca < 0) ca = 0;
    if (cb > n_out) cb = n_out;
}

__device__ __forceinline__ long long find_segment(const long long *__restrict__ seg_run_off, int n_seg, long long run) {
    int lo = 0, hi = n_seg;  // last s with seg_run_off[s] <= run
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (seg_run_off[mid] <= run) lo = mid; else hi = mid;
    }
    return lo;
}

// ---------------------------------------------------------------- fast path
// hist layout: bin-major, slot(tid) = 2*lane + (warp&1) + 64*(warp>>1) so that the 32 lanes
// of a warp touch 32 distinct banks (two warps share each 32-bit word, 16 bits each).
__global__ void __launch_bounds__(kAdjThreads)
adjust_hist_kernel(const float *__restrict__ x, const long long *__restrict__ seg_off,
                   const long long *__restrict__ seg_out_off, const long long *__restrict__ seg_run_off,
                   const double *__restrict__ seg_shift, int n_seg, long long n_runs, AdjParams P,
                   const double *__restrict__ coef, const double *__restrict__ edge_first,
                   const double *__restrict__ edge_last, double *__restrict__ out,
                   unsigned char *__restrict__ fallback) {
    extern __shared__ short hist_smem[];  // [kAdjBins][kAdjThreads]
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    short *__restrict__ h = hist_smem + (2 * lane + (warp & 1) + 64 * (warp >> 1));
#define H(b) h[(b) * kAdjThreads]
    for (int b = 0; b < kAdjBins; ++b) H(b) = 0;

    const long long run = (long long)blockIdx.x * kAdjThreads + tid;
    if (run >= n_runs) return;
    const long long s = find_segment(seg_run_off, n_seg, run);
    const long long n = seg_off[s + 1] - seg_off[s];
    const long long n_out = n - P.w;
    const float *__restrict__ xs = x + seg_off[s];
    const double shift = seg_shift ? seg_shift[s] : 0.0;
    const long long ja = (run - seg_run_off[s]) * P.run;
    const long long jb = (ja + P.run < n_out) ? ja + P.run : n_out;
    long long ca, cb;
    run_extent(P, ja, jb, n_out, ca, cb);

    SgEmitter sg;
    sg.coef = coef; sg.edge_first = edge_first; sg.edge_last = edge_last;
    sg.out = out + seg_out_off[s];
    sg.sg_w = P.sg_w; sg.half = P.sg_w / 2; sg.ja = ja; sg.jb = jb; sg.n_out = n_out;

    const int w = P.w;
    const int kl = (w - 1) >> 1, ku = w >> 1;  // 0-based ranks of the two middle order statistics
    // bin = value - base; centre the window on the first sample
    const float x0 = xs[ca];
    if (x0 != rintf(x0) || fabsf(x0) > 1.0e9f) { fallback[run] = 1; return; }
    const int base = (int)x0 - kAdjBins / 2;
    long long isum = 0;
    bool bad = false;
    auto bin_of = [&](float v) -> int {
        const int b = (int)v - base;
        if (v != rintf(v) || (unsigned)b >= (unsigned)kAdjBins) { bad = true; return 0; }
        return b;
    };
    for (int t = 0; t < w; ++t) {
        const float v = xs[ca + t];
        const int b = bin_of(v);
        H(b) = H(b) + 1;
        isum += (long long)v;
    }
    if (bad) { fallback[run] = 1; return; }
    // locate the lower median bin: c_lt = #samples below bin m
    int m = 0, c_lt = 0;
    while (c_lt + H(m) <= kl) { c_lt += H(m); ++m; }

    for (long long c = ca; c < cb; ++c) {
        double stat;
        if (P.use_mean) {
            stat = (double)isum / (double)w - shift;
        } else {
            // upper median: same bin if it still covers rank ku, else the next occupied bin
            int mu = m;
            if (c_lt + H(m) <= ku) { do { ++mu; } while (H(mu) == 0); }
            const double lo = (double)(m + base) - shift, hi = (double)(mu + base) - shift;
            stat = (lo + hi) / 2.0;
        }
        const double centre = (double)xs[c + (w >> 1)] - shift;
        const double adj = centre - stat;
        if (P.savgol) sg.push(c, adj); else sg.out[c] = adj;
        if (c + 1 < cb) {  // slide: drop xs[c], add xs[c + w]
            const float vo = xs[c], vi = xs[c + w];
            const int bo = (int)vo - base;  // validated when it entered
            const int bi = bin_of(vi);
            if (bad) { fallback[run] = 1; return; }
            H(bo) = H(bo) - 1;
            H(bi) = H(bi) + 1;
            isum += (long long)vi - (long long)vo;
            c_lt += (bi < m) - (bo < m);
            while (c_lt > kl) { --m; c_lt -= H(m); }
            while (c_lt + H(m) <= kl) { c_lt += H(m); ++m; }
        }
    }
#undef H
}

// ------------------------------------------------------------ generic path
// Any float input: the window is kept sorted in a per-thread scratch array (binary
// search + shift on every step, O(w) per output).  Only flagged runs come here.
__global__ void __launch_bounds__(64)
adjust_generic_kernel(const float *__restrict__ x, const long long *__restrict__ seg_off,
                      const long long *__restrict__ seg_out_off, const long long *__restrict__ seg_run_off,
                      const double *__restrict__ seg_shift, int n_seg,
                      const long long *__restrict__ run_list, long long n_list, AdjParams P,
                      const double *__restrict__ coef, const double *__restrict__ edge_first,
                      const double *__restrict__ edge_last, double *__restrict__ out,
                      float *__restrict__ scratch /* [n_list][w] */) {
    const long long li = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= n_list) return;
    const long long run = run_list[li];
    float *__restrict__ win = scratch + li * (long long)P.w;
    const long long s = find_segment(seg_run_off, n_seg, run);
    const long long n = seg_off[s + 1] - seg_off[s];
    const long long n_out = n - P.w;
    const float *__restrict__ xs = x + seg_off[s];
    const double shift = seg_shift ? seg_shift[s] : 0.0;
    const long long ja = (run - seg_run_off[s]) * P.run;
    const long long jb = (ja + P.run < n_out) ? ja + P.run : n_out;
    long long ca, cb;
    run_extent(P, ja, jb, n_out, ca, cb);
    SgEmitter sg;
    sg.coef = coef; sg.edge_first = edge_first; sg.edge_last = edge_last;
    sg.out = out + seg_out_off[s];
    sg.sg_w = P.sg_w; sg.half = P.sg_w / 2; sg.ja = ja; sg.jb = jb; sg.n_out = n_out;
    const int w = P.w;
    // insertion sort of the first window
    double sum = 0.0;
    for (int t = 0; t < w; ++t) {
        const float v = xs[ca + t];
        int k = t;
        while (k > 0 && win[k - 1] > v) { win[k] = win[k - 1]; --k; }
        win[k] = v;
    }
    for (long long c = ca; c < cb; ++c) {
        double stat;
        if (P.use_mean) {
            sum = 0.0;  // fp64 left-to-right sum of the shifted window (numpy: pairwise; <= 1e-13 rel apart)
            for (int t = 0; t < w; ++t) sum += (double)xs[c + t] - shift;
            stat = sum / (double)w;
        } else {
            const double lo = (double)win[(w - 1) >> 1] - shift, hi = (double)win[w >> 1] - shift;
            stat = (lo + hi) / 2.0;
        }
        const double adj = ((double)xs[c + (w >> 1)] - shift) - stat;
        if (P.savgol) sg.push(c, adj); else sg.out[c] = adj;
        if (c + 1 < cb && !P.use_mean) {
            const float vo = xs[c], vi = xs[c + w];
            int lo = 0, hi = w;  // position of one copy of vo
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (win[mid] < vo) lo = mid + 1; else hi = mid; }
            int k = lo;
            // remove win[k], insert vi keeping order
            if (vi >= vo) { while (k + 1 < w && win[k + 1] < vi) { win[k] = win[k + 1]; ++k; } }
            else { while (k > 0 && win[k - 1] > vi) { win[k] = win[k - 1]; --k; } }
            win[k] = vi;
        }
    }
}

// subtract_edges (frag/_adjust_wps.py:119-123): shift[s] = mean(mean(x[:e]), mean(x[-e:]))
__global__ void adjust_edge_shift_kernel(const float *__restrict__ x, const long long *__restrict__ seg_off,
                                         int n_seg, int edge, double *__restrict__ shift) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seg) return;
    const long long n = seg_off[s + 1] - seg_off[s];
    const float *xs = x + seg_off[s];
    const long long e = edge < n ? edge : n;
    if (e <= 0) { shift[s] = 0.0; return; }
    double a = 0.0, b = 0.0;  // exact for integer-valued samples (|sum| < 2^53)
    for (long long t = 0; t < e; ++t) a += (double)xs[t];
    for (long long t = n - e; t < n; ++t) b += (double)xs[t];
    shift[s] = ((a / (double)e) + (b / (double)e)) / 2.0;
}

}  // namespace ftk

using namespace ftk;

extern "C" int ftk_adjust_edge_shift_f64(const float *x, const int64_t *seg_off, int32_t n_seg,
                                         int32_t edge_size, double *shift, ftk_stream_t stream_) {
    if (n_seg == 0) return FTK_OK;
    if (!x || !seg_off || !shift || n_seg < 0 || edge_size < 0) return FTK_E_INVALID;
    adjust_edge_shift_kernel<<<(n_seg + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream_)>>>(
        x, reinterpret_cast<const long long *>(seg_off), n_seg, edge_size, shift);
    FTK_CHECK_LAUNCH("adjust_edge_shift_kernel");
    return FTK_OK;
}

static int adjust_check(const float *x, const int64_t *seg_off, const int64_t *seg_out_off,
                        const int64_t *seg_run_off, int32_t n_seg, int64_t n_runs, int32_t w,
                        int32_t savgol, int32_t sg_w, int32_t run_len, const double *coef,
                        const double *edge_first, const double *edge_last, double *out) {
    if (!x || !seg_off || !seg_out_off || !seg_run_off || !out || n_seg < 0 || n_runs < 0) return FTK_E_INVALID;
    if (w < 2 || (w & 1) || w > 32767 || run_len < 1) return FTK_E_INVALID;
    if (savgol && (!coef || !edge_first || !edge_last || sg_w < 1 || !(sg_w & 1) || sg_w > kAdjMaxSg))
        return FTK_E_INVALID;
    return FTK_OK;
}

extern "C" int ftk_adjust_wps_f64(const float *x, const int64_t *seg_off, const int64_t *seg_out_off,
                                  const int64_t *seg_run_off, const double *seg_shift, int32_t n_seg,
                                  int64_t n_runs, int32_t w, int32_t use_mean, int32_t savgol,
                                  int32_t sg_w, int32_t run_len, const double *coef,
                                  const double *edge_first, const double *edge_last, double *out,
                                  uint8_t *fallback, ftk_stream_t stream_) {
    if (n_runs == 0 || n_seg == 0) return FTK_OK;
    int rc = adjust_check(x, seg_off, seg_out_off, seg_run_off, n_seg, n_runs, w, savgol, sg_w, run_len,
                          coef, edge_first, edge_last, out);
    if (rc != FTK_OK) return rc;
    if (!fallback) return FTK_E_INVALID;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    AdjParams P{w, use_mean, savgol, savgol ? sg_w : 1, run_len};
    const int smem = kAdjBins * kAdjThreads * (int)sizeof(short);
    static thread_local bool attr_set = false;
    if (!attr_set) {
        FTK_CUDA_TRY(cudaFuncSetAttribute(adjust_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    FTK_CUDA_TRY(cudaMemsetAsync(fallback, 0, (size_t)n_runs, stream));
    const unsigned grid = (unsigned)((n_runs + kAdjThreads - 1) / kAdjThreads);
    adjust_hist_kernel<<<grid, kAdjThreads, smem, stream>>>(
        x, reinterpret_cast<const long long *>(seg_off), reinterpret_cast<const long long *>(seg_out_off),
        reinterpret_cast<const long long *>(seg_run_off), seg_shift, n_seg, n_runs, P, coef, edge_first,
        edge_last, out, fallback);
    FTK_CHECK_LAUNCH("adjust_hist_kernel");
    return FTK_OK;
}

extern "C" int ftk_adjust_wps_generic_f64(const float *x, const int64_t *seg_off, const int64_t *seg_out_off,
                                          const int64_t *seg_run_off, const double *seg_shift, int32_t n_seg,
                                          const int64_t *run_list, int64_t n_list, int32_t w,
                                          int32_t use_mean, int32_t savgol, int32_t sg_w, int32_t run_len,
                                          const double *coef, const double *edge_first,
                                          const double *edge_last, double *out, float *scratch,
                                          ftk_stream_t stream_) {
    if (n_list == 0 || n_seg == 0) return FTK_OK;
    int rc = adjust_check(x, seg_off, seg_out_off, seg_run_off, n_seg, n_list, w, savgol, sg_w, run_len,
                          coef, edge_first, edge_last, out);
    if (rc != FTK_OK) return rc;
    if (!run_list || !scratch) return FTK_E_INVALID;
    AdjParams P{w, use_mean, savgol, savgol ? sg_w : 1, run_len};
    adjust_generic_kernel<<<(unsigned)((n_list + 63) / 64), 64, 0, static_cast<cudaStream_t>(stream_)>>>(
        x, reinterpret_cast<const long long *>(seg_off), reinterpret_cast<const long long *>(seg_out_off),
        reinterpret_cast<const long long *>(seg_run_off), seg_shift, n_seg,
        reinterpret_cast<const long long *>(run_list), n_list, P, coef, edge_first, edge_last, out, scratch);
    FTK_CHECK_LAUNCH("adjust_generic_kernel");
    return FTK_OK;
}
