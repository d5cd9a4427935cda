// Host-side statistics over per-interval length histograms (frag_length_intervals).
//
// The reference turns every interval's fragment stream into a dict length -> count and does its
// statistics in the dict's INSERTION order with Python float arithmetic
// (frag/_frag_length.py:156-172 _find_median, :204-238).  The CUDA kernel delivers, per interval,
// the histogram and the first fragment index of every length; this routine redoes the reference's
// arithmetic on those rows - same operation order, same libm calls CPython makes (float ** 2 and
// ** 0.5 are pow(), not x*x / sqrt(): they differ in the last bit for ~0.1 % of inputs) - for all
// intervals at once on the host cores, replacing a per-interval Python loop.  "Python float
// arithmetic" includes sum(): since CPython 3.12 it is Neumaier-compensated, and the goldens under
// tests/golden were produced by the reference running on this image's Python 3.12.
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "ftk_b200.h"

// libm's pow through a volatile pointer: the host compiler must not fold pow(x, 2.0) into x * x or
// pow(x, 0.5) into sqrt(x) - CPython calls the library function, and the two differ in the last bit
static double (*volatile libm_pow)(double, double) = pow;

extern "C" int ftk_length_stats_host(const int32_t *hist, const int32_t *first_seen, int64_t n_rows,
                                     int32_t n_bins, int32_t short_reads, int32_t threads,
                                     double *mean, double *median, double *stdev, int64_t *vmin,
                                     int64_t *vmax, int64_t *count, double *frac_short) {
    if (n_rows < 0 || n_bins < 0) return FTK_E_INVALID;
    if (n_rows == 0) return FTK_OK;
    if (!hist || !first_seen || !mean || !median || !stdev || !vmin || !vmax || !count || !frac_short)
        return FTK_E_INVALID;
    int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    nt = (int)std::max<int64_t>(1, std::min<int64_t>(nt, (n_rows + 255) / 256));
    auto work = [&](int64_t r0, int64_t r1) {
        std::vector<std::pair<int32_t, int32_t>> order;   // (first-seen index, length)
        for (int64_t r = r0; r < r1; ++r) {
            const int32_t *h = hist + r * n_bins, *f = first_seen + r * n_bins;
            order.clear();
            int64_t total = 0, num = 0, n_short = 0;
            int lo = -1, hi = -1;
            for (int L = 0; L < n_bins; ++L) {
                const int64_t c = h[L];
                if (c <= 0) continue;
                order.emplace_back(f[L], L);
                total += c; num += c * L;
                if (L <= short_reads) n_short += c;
                if (lo < 0) lo = L;
                hi = L;
            }
            if (total == 0) {   // frag/_frag_length.py:206-208: every field -1
                mean[r] = median[r] = stdev[r] = frac_short[r] = -1.0;
                vmin[r] = vmax[r] = count[r] = -1;
                continue;
            }
            const double m = (double)num / (double)total;
            // quirky median: cdf over ascending values, searchsorted(side='left') for total//2 (and +1)
            auto first_reaching = [&](int64_t x) {
                int64_t cdf = 0;
                for (int L = 0; L < n_bins; ++L) {
                    if (h[L] > 0) { cdf += h[L]; if (cdf >= x) return L; }
                }
                return hi;
            };
            double med;
            if (total % 2 == 1) med = (double)first_reaching(total / 2);
            else med = ((double)first_reaching(total / 2) + (double)first_reaching(total / 2 + 1)) / 2.0;
            std::sort(order.begin(), order.end());
            // Python's sum() over floats (3.12+): Neumaier-compensated, compensation added at the end
            double acc = 0.0, comp = 0.0;
            bool started = false;
            for (const auto &e : order) {
                const double d = (double)e.second - m;
                const double x = (double)h[e.second] * libm_pow(d, 2.0);
                if (!started) { acc = x; started = true; continue; }   // int 0 + first float
                const double t = acc + x;
                if (fabs(acc) >= fabs(x)) comp += (acc - t) + x;
                else comp += (x - t) + acc;
                acc = t;
            }
            if (comp != 0.0 && isfinite(comp)) acc += comp;
            const double var = acc / (double)total;
            mean[r] = m; median[r] = med; stdev[r] = libm_pow(var, 0.5);
            vmin[r] = lo; vmax[r] = hi; count[r] = total;
            frac_short[r] = (double)n_short / (double)total;
        }
    };
    if (nt == 1) { work(0, n_rows); return FTK_OK; }
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; ++t) pool.emplace_back(work, n_rows * t / nt, n_rows * (t + 1) / nt);
    for (auto &th : pool) th.join();
    return FTK_OK;
}
