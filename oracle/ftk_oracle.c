/*
 * ftk_oracle.c - CPU restatement of FinaleToolkit's per-fragment interval
 * feature path.  TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this.  The
 * product (finaletoolkit_b200) never links or calls it.
 *
 * Each function restates the reference's algorithm literally (brute force,
 * same comparisons, same order) and cites the reference file:line it follows
 * (paths relative to /root/reference/src/finaletoolkit).  It deliberately does
 * NOT use the difference-array / prefix-scan formulation of the CUDA kernels,
 * so the two implementations are independent.
 *
 * Parity pin: checked against tests/golden/* (outputs of the unmodified
 * reference, produced by oracle/make_golden.py) in tests/test_oracle_golden.py.
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC -o oracle/libftk_oracle.so oracle/ftk_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_NONE INT64_MIN /* stands for Python None (unbounded) */

/* utils/_comparison.py:13-24 : None-tolerant comparisons */
static inline int none_leq(int64_t a, int64_t b) { return b == ORC_NONE ? 1 : a <= b; }
static inline int none_geq(int64_t a, int64_t b) { return b == ORC_NONE ? 1 : a >= b; }

/* utils/_frag_generator.py:21-55 : intersect policy; policy 0 = midpoint, 1 = any */
static inline int check_intersect(int policy, int64_t r_start, int64_t r_stop, int64_t fs, int64_t fe) {
    if (policy == 0) {
        /* Python floor division; fs+fe >= 0 for genomic coordinates, but keep floor semantics */
        int64_t s = fs + fe;
        int64_t mid = (s >= 0) ? s / 2 : -((-s + 1) / 2);
        return (r_start == ORC_NONE || mid >= r_start) && (r_stop == ORC_NONE || mid < r_stop);
    }
    return (r_start == ORC_NONE || fe > r_start) && (r_stop == ORC_NONE || fs < r_stop);
}

/* first index with starts[i] >= key (starts ascending) */
static int64_t lower_bound_i32(const int32_t *a, int64_t n, int64_t key) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = lo + (hi - lo) / 2;
        if ((int64_t)a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

/*
 * The fragment stream of one (contig, start, stop) query:
 *   io/alignment.py:270-302  tabix fetch = rows overlapping [start, stop) in file
 *                            order (rec.stop > start && rec.start < stop), mapq < q dropped
 *   utils/_frag_generator.py:117-130  inclusive length filter + intersect policy
 * Fragments are start-sorted (tabix requirement); max_frag_len bounds how far
 * left of `start` an overlapping row can begin, so the scan is a bisect + walk
 * instead of a whole-file pass.  Calls cb(i) for each passing row index, in order.
 */
typedef void (*frag_cb)(int64_t i, void *ctx);

static void frag_stream(const int32_t *fs, const int32_t *fe, const uint8_t *mapq, int64_t n,
                        int64_t max_frag_len, int64_t r_start, int64_t r_stop,
                        int64_t min_len, int64_t max_len, int policy, int64_t q,
                        frag_cb cb, void *ctx) {
    int64_t fetch_lo = (r_start == ORC_NONE) ? 0 : r_start; /* tabix: None -> 0 */
    int64_t i0 = (r_start == ORC_NONE) ? 0 : lower_bound_i32(fs, n, fetch_lo - max_frag_len);
    for (int64_t i = i0; i < n; ++i) {
        int64_t s = fs[i], e = fe[i];
        if (r_stop != ORC_NONE && s >= r_stop) break;       /* rec.start < stop  */
        if (!(e > fetch_lo)) continue;                       /* rec.stop  > start */
        if ((int64_t)mapq[i] < q) continue;                  /* alignment.py:291  */
        int64_t len = e - s;
        if (none_geq(len, min_len) && none_leq(len, max_len) &&
            check_intersect(policy, r_start, r_stop, s, e))
            cb(i, ctx);
    }
}

/* ------------------------------------------------------------------ WPS */
typedef struct { int64_t *s, *e; int64_t n, cap; const int32_t *fs, *fe; } sel_t;
static void sel_push(int64_t i, void *ctx) {
    sel_t *t = (sel_t *)ctx;
    if (t->n == t->cap) {
        t->cap = t->cap ? t->cap * 2 : 1024;
        t->s = (int64_t *)realloc(t->s, t->cap * sizeof(int64_t));
        t->e = (int64_t *)realloc(t->e, t->cap * sizeof(int64_t));
    }
    t->s[t->n] = t->fs[i]; t->e[t->n] = t->fe[i]; t->n++;
}

/*
 * frag/_wps.py:142-188 (wps) with the kernel frag/_wps.py:25-53 (_single_nt_wps).
 * out[stop-start] int64.  Returns number of positions written (0 for a
 * degenerate interval, _wps.py:145-152).
 */
int64_t orc_wps_interval(const int32_t *fs, const int32_t *fe, const uint8_t *mapq, int64_t n,
                         int64_t max_frag_len, int64_t start, int64_t stop, int64_t chrom_size,
                         int64_t window_size, int64_t min_len, int64_t max_len, int64_t q,
                         int64_t *out) {
    if (stop <= start) return 0;
    /* _wps.py:156-157 */
    int64_t minimum = start - max_len; if (minimum < 0) minimum = 0;
    int64_t maximum = stop + max_len;  if (maximum > chrom_size) maximum = chrom_size;
    sel_t sel = {0, 0, 0, 0, fs, fe};
    /* _wps.py:159-169 : frag_array(..., start=minimum, stop=maximum, policy midpoint) */
    frag_stream(fs, fe, mapq, n, max_frag_len, minimum, maximum, min_len, max_len, 0, q, sel_push, &sel);
    for (int64_t c = start; c < stop; ++c) {
        /* _wps.py:176-178 : np.rint == C rint() under the default round-half-even mode */
        double ws = rint((double)c - (double)window_size * 0.5);
        double we = rint((double)c + (double)window_size * 0.5 - 1.0);
        int64_t num_spanning = 0, num_end_in = 0;
        for (int64_t j = 0; j < sel.n; ++j) {               /* _wps.py:39-53 */
            double s = (double)sel.s[j], e = (double)sel.e[j];
            int is_spanning = (s < ws) * (e > we);
            int is_start_in = (s >= ws) * (s <= we);
            int is_stop_in = (e >= ws) * (e <= we);
            num_spanning += is_spanning;
            num_end_in += (is_start_in || is_stop_in);
        }
        out[c - start] = num_spanning - num_end_in;
    }
    free(sel.s); free(sel.e);
    return stop - start;
}

/*
 * frag/_multi_wps.py:196-198 : Pool(workers).imap(wps) over intervals.  The
 * multiprocessing pool becomes an OpenMP loop over intervals (same unit of
 * parallelism).  out_off[k] = offset of interval k in `out`.
 */
void orc_wps_intervals(const int32_t *fs, const int32_t *fe, const uint8_t *mapq, int64_t n,
                       int64_t max_frag_len, const int64_t *ivl_start, const int64_t *ivl_stop,
                       const int64_t *out_off, int64_t n_ivl, int64_t chrom_size,
                       int64_t window_size, int64_t min_len, int64_t max_len, int64_t q,
                       int64_t *out, int n_threads) {
#pragma omp parallel for schedule(dynamic, 4) num_threads(n_threads)
    for (int64_t k = 0; k < n_ivl; ++k)
        orc_wps_interval(fs, fe, mapq, n, max_frag_len, ivl_start[k], ivl_stop[k], chrom_size,
                         window_size, min_len, max_len, q, out + out_off[k]);
}

/* ------------------------------------------------------- coverage / lengths */
static void count_cb(int64_t i, void *ctx) { (void)i; (*(int64_t *)ctx)++; }

/* frag/_coverage.py:117-130 (single_coverage): count of the fragment stream */
int64_t orc_single_coverage(const int32_t *fs, const int32_t *fe, const uint8_t *mapq, int64_t n,
                            int64_t max_frag_len, int64_t start, int64_t stop,
                            int64_t min_len, int64_t max_len, int policy, int64_t q) {
    int64_t c = 0;
    frag_stream(fs, fe, mapq, n, max_frag_len, start, stop, min_len, max_len, policy, q, count_cb, &c);
    return c;
}

/* frag/_coverage.py:244-248 : Pool.imap(single_coverage) over intervals */
void orc_interval_coverage(const int32_t *fs, const int32_t *fe, const uint8_t *mapq, int64_t n,
                           int64_t max_frag_len, const int64_t *ivl_start, const int64_t *ivl_stop,
                           int64_t n_ivl, int64_t min_len, int64_t max_len, int policy, int64_t q,
                           int64_t *counts, int n_threads) {
#pragma omp parallel for schedule(dynamic, 4) num_threads(n_threads)
    for (int64_t k = 0; k < n_ivl; ++k)
        counts[k] = orc_single_coverage(fs, fe, mapq, n, max_frag_len, ivl_start[k], ivl_stop[k],
                                        min_len, max_len, policy, q);
}

/*
 * frag/_frag_length.py:147-153 (_distribution_from_gen): dict length -> count.
 * A Python dict iterates in first-insertion order, which the reference's fp
 * sums depend on (_frag_length.py:213-217, 435-438), so the oracle returns the
 * distinct lengths in first-seen order: keys[j], vals[j], j < returned count.
 * frag/_frag_length.py:303 (frag_length): also the raw lengths in stream order
 * when `lengths` != NULL (capacity cap_lengths; returns -1 on overflow).
 */
typedef struct { int64_t *keys, *vals; int64_t nk, cap; const int32_t *fs, *fe; int32_t *lengths; int64_t nl, cap_l; int overflow; } dist_t;
static void dist_cb(int64_t i, void *ctx) {
    dist_t *d = (dist_t *)ctx;
    int64_t len = (int64_t)d->fe[i] - (int64_t)d->fs[i];
    if (d->lengths) { if (d->nl < d->cap_l) d->lengths[d->nl] = (int32_t)len; else d->overflow = 1; d->nl++; }
    if (!d->keys) return;
    for (int64_t j = 0; j < d->nk; ++j) if (d->keys[j] == len) { d->vals[j]++; return; }
    if (d->nk < d->cap) { d->keys[d->nk] = len; d->vals[d->nk] = 1; } else d->overflow = 1;
    d->nk++;
}

int64_t orc_length_dist(const int32_t *fs, const int32_t *fe, const uint8_t *mapq, int64_t n,
                        int64_t max_frag_len, int64_t start, int64_t stop,
                        int64_t min_len, int64_t max_len, int policy, int64_t q,
                        int64_t *keys, int64_t *vals, int64_t cap) {
    dist_t d = {keys, vals, 0, cap, fs, fe, 0, 0, 0, 0};
    frag_stream(fs, fe, mapq, n, max_frag_len, start, stop, min_len, max_len, policy, q, dist_cb, &d);
    return d.overflow ? -1 : d.nk;
}

int64_t orc_frag_lengths(const int32_t *fs, const int32_t *fe, const uint8_t *mapq, int64_t n,
                         int64_t max_frag_len, int64_t start, int64_t stop,
                         int64_t min_len, int64_t max_len, int policy, int64_t q,
                         int32_t *lengths, int64_t cap) {
    dist_t d = {0, 0, 0, 0, fs, fe, lengths, 0, cap, 0};
    frag_stream(fs, fe, mapq, n, max_frag_len, start, stop, min_len, max_len, policy, q, dist_cb, &d);
    return d.overflow ? -1 : d.nl;
}

/* ------------------------------------------------------------- end motifs */
/*
 * frag/_end_motifs.py:115-179 (region_end_motifs), io/reference.py:155-176
 * (bounds check -> OutOfBoundsError(ValueError)), utils/utils.py:388-437
 * (gen_kmers order A<C<G<T; reverse_complement).
 * seq: upper-case ASCII contig (N for unknown).  strand_mode: 0 both strands,
 * 1 forward only (is_forward fragments), 2 negative only.
 * counts[4^k] int64 is ADDED to.  Returns 0, or 1 if the reference would raise
 * RuntimeError (reverse k-mer out of bounds, _end_motifs.py:144-151).
 */
static inline int base_code(char b) {
    switch (b) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return -1; }
}
typedef struct { const int32_t *fs, *fe; const uint8_t *strand; const char *seq; int64_t seq_len; int k; int strand_mode; int64_t *counts; int err; } motif_t;

static void motif_cb(int64_t i, void *ctx) {
    motif_t *m = (motif_t *)ctx;
    if (m->err) return;
    int k = m->k;
    int64_t s = m->fs[i], e = m->fe[i];
    int do_fwd = (m->strand_mode == 0) || (m->strand_mode == 1 && m->strand[i]);
    int do_rev = (m->strand_mode == 0) || (m->strand_mode == 2);
    if (m->strand_mode == 1 && !m->strand[i]) return;        /* _end_motifs.py:154 */
    if (do_fwd) {
        /* refseq.sequence(contig, s, s+k): OOB -> ValueError -> `continue` (skips reverse too) */
        if (s < 0 || s + k > m->seq_len) return;
        int64_t idx = 0; int ok = 1;
        for (int j = 0; j < k; ++j) { int c = base_code(m->seq[s + j]); if (c < 0) { ok = 0; break; } idx = idx * 4 + c; }
        if (ok) m->counts[idx]++;
    }
    if (do_rev) {
        /* refseq.sequence(contig, e-k, e) */
        if (e - k < 0 || e > m->seq_len) {
            if (m->strand_mode == 0) m->err = 1;             /* RuntimeError */
            return;                                           /* negative-only: continue (:172-179) */
        }
        int64_t idx = 0; int ok = 1;
        for (int j = 0; j < k; ++j) {                        /* reverse complement: 3-code, reversed */
            int c = base_code(m->seq[e - 1 - j]); if (c < 0) { ok = 0; break; } idx = idx * 4 + (3 - c);
        }
        if (ok) m->counts[idx]++;
    }
}

int orc_region_end_motifs(const int32_t *fs, const int32_t *fe, const uint8_t *mapq, const uint8_t *strand,
                          int64_t n, int64_t max_frag_len, int64_t start, int64_t stop,
                          const char *seq, int64_t seq_len, int k, int strand_mode, int64_t q,
                          int64_t *counts) {
    motif_t m = {fs, fe, strand, seq, seq_len, k, strand_mode, counts, 0};
    /* AlignmentWrapper.fetch only: no length filter, membership = tabix overlap (policy any) */
    frag_stream(fs, fe, mapq, n, max_frag_len, start, stop, ORC_NONE, ORC_NONE, 1, q, motif_cb, &m);
    return m.err;
}

/* region_breakpoint_motifs, frag/_breakpoint_motifs.py:120-186: k-mers centred on the breakpoints. */
static void bkpt_cb(int64_t i, void *ctx) {
    motif_t *m = (motif_t *)ctx;
    int k = m->k, h = k / 2;
    int64_t s = m->fs[i], e = m->fe[i];
    if (s - h < 0 || s + h >= m->seq_len) return;            /* :125-133 too close to a contig end */
    int do_fwd = (m->strand_mode == 0) || (m->strand_mode == 1 && m->strand[i]);   /* :135-138 */
    int do_rev = (m->strand_mode == 0) || (m->strand_mode == 2);
    if (do_fwd) {
        if (2 * h != k) return;                              /* :145-152 len(kmer) != k -> continue */
        int64_t idx = 0; int ok = 1;
        for (int j = 0; j < k; ++j) { int c = base_code(m->seq[s - h + j]); if (c < 0) { ok = 0; break; } idx = idx * 4 + c; }
        if (ok) m->counts[idx]++;
    }
    if (do_rev) {
        if (e - h < 0 || e + h > m->seq_len || e - h > e + h) return;   /* OutOfBoundsError (ValueError) -> continue */
        if (2 * h != k) return;
        int64_t idx = 0; int ok = 1;
        for (int j = 0; j < k; ++j) {
            int c = base_code(m->seq[e + h - 1 - j]); if (c < 0) { ok = 0; break; } idx = idx * 4 + (3 - c);
        }
        if (ok) m->counts[idx]++;
    }
}

int orc_region_breakpoint_motifs(const int32_t *fs, const int32_t *fe, const uint8_t *mapq, const uint8_t *strand,
                                 int64_t n, int64_t max_frag_len, int64_t start, int64_t stop,
                                 const char *seq, int64_t seq_len, int k, int strand_mode, int64_t q,
                                 int64_t *counts) {
    motif_t m = {fs, fe, strand, seq, seq_len, k, strand_mode, counts, 0};
    frag_stream(fs, fe, mapq, n, max_frag_len, start, stop, ORC_NONE, ORC_NONE, 1, q, bkpt_cb, &m);
    return 0;
}

/* ------------------------------------------------------- DELFI window */
/*
 * frag/_delfi.py:404-511 (_delfi_single_window), the counting part: short / long / num_frags of one
 * bin and the number of G + C bases of the bin.  The NOARM test and NaN conventions are in oracle.py.
 *   bl_start/bl_stop : ALL blacklist regions of the contig sorted by (start, stop) (:85-107)
 *   telo             : n_telo (start, stop) pairs; gaps_use = 0 when the contig has no ContigGaps
 * out[4] = {short, long, num_frags, num_gc}
 */
typedef struct {
    const int32_t *fs, *fe; int64_t ws, we;
    const int64_t *r0, *r1; int64_t n_r;            /* blacklist regions contained in the bin */
    int gaps_use; int64_t c0, c1; const int64_t *telo; int64_t n_telo;
    int64_t n_short, n_long;
} delfi_t;

static void delfi_cb(int64_t i, void *ctx) {
    delfi_t *d = (delfi_t *)ctx;
    int64_t s = d->fs[i], e = d->fe[i], len = e - s;
    if (len < 100 || len > 220) return;                                   /* :442-443 */
    int64_t mid = (s + e) / 2;                                            /* s, e >= 0 */
    if (mid < d->ws || mid >= d->we) return;                              /* :445-447 */
    int blacklisted = 0;
    for (int64_t j = 0; j < d->n_r; ++j)                                  /* :449-457 */
        if (s >= d->r0[j] && s < d->r1[j] && e >= d->r0[j] && e < d->r1[j]) { blacklisted = 1; break; }
    if (d->gaps_use) {                                                    /* genome/gaps.py:226-248 */
        int in_c = e > d->c0 && s < d->c1;
        int in_t = 0;
        if (d->n_telo > 0) {
            in_t = 1;
            for (int64_t t = 0; t < d->n_telo; ++t)
                if (!(e > d->telo[2 * t] && s < d->telo[2 * t + 1])) { in_t = 0; break; }   /* all() */
        }
        if (in_c || in_t) return;                                         /* :459-460 */
    }
    if (blacklisted) return;
    if (len >= 151) d->n_long++; else d->n_short++;                       /* :462-467 */
}

void orc_delfi_window(const int32_t *fs, const int32_t *fe, const uint8_t *mapq, int64_t n, int64_t max_frag_len,
                      int64_t ws, int64_t we, const int64_t *bl_start, const int64_t *bl_stop, int64_t n_bl,
                      int gaps_use, int64_t c0, int64_t c1, const int64_t *telo, int64_t n_telo, int64_t q,
                      const char *seq, int64_t seq_len, int64_t *out) {
    /* _blacklist_in_window (:110-127): start >= ws by binary search, then stop <= we */
    int64_t lo = 0, hi = n_bl;
    while (lo < hi) { int64_t m = (lo + hi) / 2; if (bl_start[m] < ws) lo = m + 1; else hi = m; }
    int64_t cap = n_bl - lo, n_r = 0;
    int64_t *r0 = (int64_t *)malloc((size_t)(cap > 0 ? cap : 1) * sizeof(int64_t));
    int64_t *r1 = (int64_t *)malloc((size_t)(cap > 0 ? cap : 1) * sizeof(int64_t));
    for (int64_t j = lo; j < n_bl; ++j)
        if (bl_stop[j] <= we) { r0[n_r] = bl_start[j]; r1[n_r] = bl_stop[j]; n_r++; }
    delfi_t d = {fs, fe, ws, we, r0, r1, n_r, gaps_use, c0, c1, telo, n_telo, 0, 0};
    /* AlignmentWrapper.fetch(contig, ws, we): tabix overlap + mapq only */
    frag_stream(fs, fe, mapq, n, max_frag_len, ws, we, ORC_NONE, ORC_NONE, 1, q, delfi_cb, &d);
    free(r0); free(r1);
    int64_t gc = 0;
    /* valid_interval (utils/validation.py:146-166), else ref_bases = "" (:472-482) */
    if (seq && ws >= 0 && ws < seq_len && we >= 0 && we <= seq_len)
        for (int64_t p = ws; p < we; ++p) gc += (seq[p] == 'G' || seq[p] == 'C');
    out[0] = d.n_short; out[1] = d.n_long; out[2] = d.n_short + d.n_long; out[3] = gc;
}

/* ------------------------------------------------------- cleavage profile */
/*
 * frag/_cleavage_profile.py:188-217 (cleavage_profile) with _coverage_and_ends (:33-90) restated
 * as the definition it implements: for every position p of [adj_start, adj_stop), over the
 * fragments of frag_array(..., start=adj_start, stop=adj_stop, intersect_policy="any"):
 *   depth = #{start <= p < stop},  ends = #{'+' and start == p} + #{'-' and stop == p},
 *   proportion = depth ? ends / depth * 100 : 0.
 */
typedef struct { sel_t s; const uint8_t *strand; uint8_t *sd; } clv_t;
static void clv_push(int64_t i, void *ctx) {
    clv_t *c = (clv_t *)ctx;
    int64_t before = c->s.cap;
    sel_push(i, &c->s);
    if (c->s.cap != before) c->sd = (uint8_t *)realloc(c->sd, (size_t)c->s.cap);
    c->sd[c->s.n - 1] = c->strand[i];
}
int64_t orc_cleavage_interval(const int32_t *fs, const int32_t *fe, const uint8_t *mapq, const uint8_t *strand,
                              int64_t n, int64_t max_frag_len, int64_t start, int64_t stop, int64_t left,
                              int64_t right, int64_t chrom_size, int64_t min_len, int64_t max_len, int64_t q,
                              double *out) {
    int64_t adj_start = start - left; if (adj_start < 0) adj_start = 0;
    int64_t adj_stop = stop + right;  if (adj_stop > chrom_size) adj_stop = chrom_size;
    if (adj_stop <= adj_start) return 0;
    clv_t c; memset(&c, 0, sizeof(c)); c.s.fs = fs; c.s.fe = fe; c.strand = strand;
    frag_stream(fs, fe, mapq, n, max_frag_len, adj_start, adj_stop, min_len, max_len, 1, q, clv_push, &c);
    for (int64_t p = adj_start; p < adj_stop; ++p) {
        int64_t depth = 0, ends = 0;
        for (int64_t j = 0; j < c.s.n; ++j) {
            depth += (c.s.s[j] <= p && p < c.s.e[j]);
            ends += c.sd[j] ? (c.s.s[j] == p) : (c.s.e[j] == p);
        }
        out[p - adj_start] = depth ? (double)ends / (double)depth * 100.0 : 0.0;
    }
    free(c.s.s); free(c.s.e); free(c.sd);
    return adj_stop - adj_start;
}

/* ------------------------------------------------------------- adjust_wps */
static int cmp_double(const void *a, const void *b) {
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}
/*
 * frag/_adjust_wps.py:25-45 (_running_stat + _local_filter) with np.median /
 * np.mean semantics: out[j] = data[j + w/2] - stat(data[j : j+w]), j in [0, n-w).
 * Sort-based median (mean of the two middle values for even w), mean as a
 * pairwise-free left-to-right sum is NOT numpy's pairwise sum, so the mean
 * variant is compared with a tolerance in the tests; the median is exact.
 * Returns the number of outputs.
 */
int64_t orc_local_filter(const double *data, int64_t n, int64_t w, int use_mean, double *out) {
    int64_t nw = n - w;
    if (nw <= 0) return 0;
    double *buf = (double *)malloc(sizeof(double) * (size_t)w);
    for (int64_t j = 0; j < nw; ++j) {
        double stat;
        if (use_mean) {
            double acc = 0.0; for (int64_t t = 0; t < w; ++t) acc += data[j + t];
            stat = acc / (double)w;
        } else {
            memcpy(buf, data + j, sizeof(double) * (size_t)w);
            qsort(buf, (size_t)w, sizeof(double), cmp_double);
            stat = (w % 2) ? buf[w / 2] : 0.5 * (buf[w / 2 - 1] + buf[w / 2]);
        }
        out[j] = data[j + w / 2] - stat;
    }
    free(buf);
    return nw;
}
