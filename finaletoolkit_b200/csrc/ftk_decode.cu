// Host-side fragment-file decoder: BGZF/gzip text (.frag.gz, BED6 .bed.gz) -> columnar arrays.
// (Host code only; it lives in a .cu file so the one-command nvcc build picks it up.)
//
// Replaces the reference's per-interval text stream - pysam.TabixFile.fetch + int() per field,
// io/alignment.py:270-302, re-opened for every interval by utils/_frag_generator.py:112 - with a
// single multi-threaded pass: BGZF blocks are independent deflate members (BSIZE in the 'BC'
// extra field), so they are inflated in parallel into one text buffer, which is then cut at line
// boundaries and parsed by the same threads into per-contig int32/uint8 columns.
//   * 5 columns `chrom start stop mapq strand`, or BED6 (`mapq` = column 5, `strand` = column 6)
//     when the first data line has more than 5 columns (io/alignment.py:143-156);
//   * strand = '+' anywhere in the strand field (io/alignment.py:286,289);
//   * malformed rows are skipped (io/alignment.py:301-302); '#' lines are ignored;
//   * rows keep file order per contig, contigs keep order of first appearance.
// The mapq filter is NOT applied here - it is a kernel predicate.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "ftk_b200.h"

namespace {

struct Columns {
    std::string name;
    std::vector<int32_t> start, stop;
    std::vector<uint8_t> mapq, strand;
};

struct FragFile {
    std::vector<Columns> contigs;
    int bed6 = 0;
    int64_t skipped = 0;
};

struct Block { size_t off, csize, usize, uoff; };

bool read_file(const char *path, std::vector<unsigned char> &buf) {
    FILE *fh = fopen(path, "rb");
    if (!fh) return false;
    fseek(fh, 0, SEEK_END);
    long n = ftell(fh);
    fseek(fh, 0, SEEK_SET);
    buf.resize(n > 0 ? (size_t)n : 0);
    size_t got = n > 0 ? fread(buf.data(), 1, (size_t)n, fh) : 0;
    fclose(fh);
    return got == buf.size();
}

// BGZF block list, or empty if the file is not BGZF (plain gzip)
std::vector<Block> scan_bgzf(const std::vector<unsigned char> &b) {
    std::vector<Block> blocks;
    size_t p = 0, uoff = 0;
    while (p + 18 <= b.size()) {
        if (b[p] != 31 || b[p + 1] != 139 || b[p + 2] != 8 || !(b[p + 3] & 4)) return {};
        const unsigned xlen = b[p + 10] | (b[p + 11] << 8);
        size_t q = p + 12, xend = q + xlen;
        long bsize = -1;
        while (q + 4 <= xend && xend <= b.size()) {
            const unsigned slen = b[q + 2] | (b[q + 3] << 8);
            if (b[q] == 'B' && b[q + 1] == 'C' && slen == 2) bsize = (b[q + 4] | (b[q + 5] << 8)) + 1;
            q += 4 + slen;
        }
        if (bsize < 0 || p + (size_t)bsize > b.size()) return {};
        const size_t end = p + (size_t)bsize;
        const size_t usize = (size_t)b[end - 4] | ((size_t)b[end - 3] << 8) | ((size_t)b[end - 2] << 16) | ((size_t)b[end - 1] << 24);
        blocks.push_back({xend, end - 8 - xend, usize, uoff});
        uoff += usize;
        p = end;
    }
    if (p != b.size()) return {};
    return blocks;
}

bool inflate_raw(const unsigned char *src, size_t n, unsigned char *dst, size_t cap) {
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, -15) != Z_OK) return false;
    zs.next_in = const_cast<unsigned char *>(src); zs.avail_in = (uInt)n;
    zs.next_out = dst; zs.avail_out = (uInt)cap;
    const int rc = inflate(&zs, Z_FINISH);
    const bool ok = (rc == Z_STREAM_END) && zs.total_out == cap;
    inflateEnd(&zs);
    return ok;
}

// plain (multi-member) gzip fallback, single-threaded
bool inflate_gzip_all(const std::vector<unsigned char> &b, std::vector<unsigned char> &out) {
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, 15 + 32) != Z_OK) return false;
    zs.next_in = const_cast<unsigned char *>(b.data()); zs.avail_in = (uInt)b.size();
    std::vector<unsigned char> chunk(1 << 20);
    for (;;) {
        zs.next_out = chunk.data(); zs.avail_out = (uInt)chunk.size();
        const int rc = inflate(&zs, Z_NO_FLUSH);
        out.insert(out.end(), chunk.data(), chunk.data() + (chunk.size() - zs.avail_out));
        if (rc == Z_STREAM_END) {
            if (zs.avail_in == 0) break;
            if (inflateReset(&zs) != Z_OK) { inflateEnd(&zs); return false; }   // next member
        } else if (rc != Z_OK) { inflateEnd(&zs); return false; }
    }
    inflateEnd(&zs);
    return true;
}

inline bool parse_int(const char *s, const char *e, long long &v) {
    if (s == e) return false;
    bool neg = false;
    if (*s == '-' || *s == '+') { neg = (*s == '-'); ++s; if (s == e) return false; }
    long long x = 0;
    for (; s < e; ++s) {
        if (*s < '0' || *s > '9') return false;
        x = x * 10 + (*s - '0');
        if (x > (1LL << 40)) return false;
    }
    v = neg ? -x : x;
    return true;
}

struct Segment { std::string name; std::vector<int32_t> start, stop; std::vector<uint8_t> mapq, strand; };

void parse_range(const char *p, const char *end, int bed6, std::vector<Segment> &segs, int64_t &skipped) {
    Segment *cur = nullptr;
    while (p < end) {
        const char *eol = (const char *)memchr(p, '\n', (size_t)(end - p));
        if (!eol) eol = end;
        const char *le = (eol > p && eol[-1] == '\r') ? eol - 1 : eol;
        if (le > p && *p != '#') {
            const char *f[8]; const char *fe[8];
            int nf = 0;
            const char *s = p;
            while (nf < 8) {
                const char *t = (const char *)memchr(s, '\t', (size_t)(le - s));
                f[nf] = s; fe[nf] = t ? t : le; ++nf;
                if (!t) break;
                s = t + 1;
            }
            long long a, b, q;
            const int qi = bed6 ? 4 : 3, si = bed6 ? 5 : 4;
            if (nf > si && parse_int(f[1], fe[1], a) && parse_int(f[2], fe[2], b) && parse_int(f[qi], fe[qi], q) &&
                a >= INT32_MIN && a <= INT32_MAX && b >= INT32_MIN && b <= INT32_MAX) {
                const size_t nl = (size_t)(fe[0] - f[0]);
                if (!cur || cur->name.size() != nl || memcmp(cur->name.data(), f[0], nl) != 0) {
                    segs.emplace_back();
                    cur = &segs.back();
                    cur->name.assign(f[0], nl);
                }
                cur->start.push_back((int32_t)a);
                cur->stop.push_back((int32_t)b);
                cur->mapq.push_back((uint8_t)(q < 0 ? 0 : (q > 255 ? 255 : q)));
                cur->strand.push_back(memchr(f[si], '+', (size_t)(fe[si] - f[si])) ? 1 : 0);
            } else {
                ++skipped;
            }
        }
        p = eol + 1;
    }
}

}  // namespace

extern "C" void *ftk_fragfile_open(const char *path, int32_t n_threads, int32_t *err) {
    auto fail = [&](int code) -> void * { if (err) *err = code; return nullptr; };
    if (!path) return fail(FTK_E_INVALID);
    std::vector<unsigned char> raw;
    if (!read_file(path, raw)) return fail(FTK_E_IO);
    if (n_threads < 1) n_threads = (int32_t)std::max(1u, std::thread::hardware_concurrency());
    std::vector<unsigned char> text;
    std::vector<Block> blocks = scan_bgzf(raw);
    if (!blocks.empty()) {
        const size_t total = blocks.back().uoff + blocks.back().usize;
        text.resize(total);
        std::vector<std::thread> th;
        std::vector<int> ok((size_t)n_threads, 1);
        for (int t = 0; t < n_threads; ++t)
            th.emplace_back([&, t] {
                for (size_t i = (size_t)t; i < blocks.size(); i += (size_t)n_threads) {
                    const Block &b = blocks[i];
                    if (b.usize && !inflate_raw(raw.data() + b.off, b.csize, text.data() + b.uoff, b.usize)) ok[(size_t)t] = 0;
                }
            });
        for (auto &x : th) x.join();
        for (int v : ok) if (!v) return fail(FTK_E_IO);
    } else if (!raw.empty()) {
        if (raw.size() >= 2 && raw[0] == 31 && raw[1] == 139) {
            if (!inflate_gzip_all(raw, text)) return fail(FTK_E_IO);
        } else {
            text.swap(raw);   // uncompressed text
        }
    }
    std::vector<unsigned char>().swap(raw);
    const char *base = reinterpret_cast<const char *>(text.data());
    const char *end = base + text.size();
    // BED6 detection on the first data line (io/alignment.py:143-156)
    int bed6 = 0;
    for (const char *p = base; p < end;) {
        const char *eol = (const char *)memchr(p, '\n', (size_t)(end - p));
        if (!eol) eol = end;
        if (eol > p && *p != '#') {
            int tabs = 0;
            for (const char *c = p; c < eol; ++c) tabs += (*c == '\t');
            bed6 = (tabs + 1) > 5;
            break;
        }
        p = eol + 1;
    }
    // cut at line boundaries, parse in parallel
    const int T = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_threads, text.size() / (1 << 20) + 1));
    std::vector<const char *> cut((size_t)T + 1);
    cut[0] = base; cut[(size_t)T] = end;
    for (int t = 1; t < T; ++t) {
        const char *p = base + text.size() * (size_t)t / (size_t)T;
        const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
        cut[(size_t)t] = nl ? nl + 1 : end;
    }
    std::vector<std::vector<Segment>> parts((size_t)T);
    std::vector<int64_t> skipped((size_t)T, 0);
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t] { parse_range(cut[(size_t)t], cut[(size_t)t + 1], bed6, parts[(size_t)t], skipped[(size_t)t]); });
        for (auto &x : th) x.join();
    }
    FragFile *ff = new FragFile();
    ff->bed6 = bed6;
    for (int64_t s : skipped) ff->skipped += s;
    for (auto &segs : parts)
        for (auto &sg : segs) {
            Columns *dst = nullptr;
            for (auto &c : ff->contigs) if (c.name == sg.name) { dst = &c; break; }
            if (!dst) { ff->contigs.emplace_back(); dst = &ff->contigs.back(); dst->name = sg.name; }
            dst->start.insert(dst->start.end(), sg.start.begin(), sg.start.end());
            dst->stop.insert(dst->stop.end(), sg.stop.begin(), sg.stop.end());
            dst->mapq.insert(dst->mapq.end(), sg.mapq.begin(), sg.mapq.end());
            dst->strand.insert(dst->strand.end(), sg.strand.begin(), sg.strand.end());
        }
    if (err) *err = FTK_OK;
    return ff;
}

// Decode only the BGZF blocks [coffset_beg, coffset_end] of a tabix-indexed file: the text from
// virtual offset (coffset_beg, uoffset_beg) up to (coffset_end, uoffset_end), which is where the
// .tbi index says one contig's records live.  Neighbouring contigs that share the boundary blocks
// come along as extra entries; the caller picks the contig it asked for.
extern "C" void *ftk_fragfile_open_slice(const char *path, int64_t coffset_beg, int32_t uoffset_beg,
                                         int64_t coffset_end, int32_t uoffset_end, int32_t bed6,
                                         int32_t n_threads, int32_t *err) {
    auto fail = [&](int code) -> void * { if (err) *err = code; return nullptr; };
    if (!path || coffset_beg < 0 || coffset_end < coffset_beg || uoffset_beg < 0 || uoffset_end < 0)
        return fail(FTK_E_INVALID);
    FILE *fh = fopen(path, "rb");
    if (!fh) return fail(FTK_E_IO);
    // a BGZF block is at most 64 KiB: read through the end of the block that starts at coffset_end
    const size_t want = (size_t)(coffset_end - coffset_beg) + (uoffset_end > 0 ? 65536 + 64 : 0);
    std::vector<unsigned char> raw(want);
    if (fseeko(fh, (off_t)coffset_beg, SEEK_SET) != 0) { fclose(fh); return fail(FTK_E_IO); }
    raw.resize(want ? fread(raw.data(), 1, want, fh) : 0);
    fclose(fh);
    if (n_threads < 1) n_threads = (int32_t)std::max(1u, std::thread::hardware_concurrency());
    // block list up to and including the block at coffset_end (only needed when uoffset_end > 0)
    std::vector<Block> blocks;
    size_t p = 0, uoff = 0, last_uoff = 0;
    const size_t last_rel = (size_t)(coffset_end - coffset_beg);
    while (p + 18 <= raw.size() && (p < last_rel || (p == last_rel && uoffset_end > 0))) {
        if (raw[p] != 31 || raw[p + 1] != 139 || raw[p + 2] != 8 || !(raw[p + 3] & 4)) return fail(FTK_E_IO);
        const unsigned xlen = raw[p + 10] | (raw[p + 11] << 8);
        size_t q = p + 12;
        const size_t xend = q + xlen;
        long bsize = -1;
        while (q + 4 <= xend && xend <= raw.size()) {
            const unsigned slen = raw[q + 2] | (raw[q + 3] << 8);
            if (raw[q] == 'B' && raw[q + 1] == 'C' && slen == 2) bsize = (raw[q + 4] | (raw[q + 5] << 8)) + 1;
            q += 4 + slen;
        }
        if (bsize < 0 || p + (size_t)bsize > raw.size()) return fail(FTK_E_IO);
        const size_t end = p + (size_t)bsize;
        const size_t usize = (size_t)raw[end - 4] | ((size_t)raw[end - 3] << 8) | ((size_t)raw[end - 2] << 16) |
                             ((size_t)raw[end - 1] << 24);
        if (p == last_rel) last_uoff = uoff;
        blocks.push_back({xend, end - 8 - xend, usize, uoff});
        uoff += usize;
        p = end;
    }
    if (p < last_rel) return fail(FTK_E_IO);          // the index points past what the file holds
    std::vector<unsigned char> text(uoff);
    {
        const int T = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_threads, blocks.size()));
        std::vector<std::thread> th;
        std::vector<int> ok((size_t)T, 1);
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t] {
                for (size_t i = (size_t)t; i < blocks.size(); i += (size_t)T) {
                    const Block &b = blocks[i];
                    if (b.usize && !inflate_raw(raw.data() + b.off, b.csize, text.data() + b.uoff, b.usize)) ok[(size_t)t] = 0;
                }
            });
        for (auto &x : th) x.join();
        for (int v : ok) if (!v) return fail(FTK_E_IO);
    }
    const size_t t_end = (uoffset_end > 0) ? last_uoff + (size_t)uoffset_end : uoff;
    if ((size_t)uoffset_beg > t_end || t_end > text.size()) return fail(FTK_E_IO);
    const char *base = reinterpret_cast<const char *>(text.data()) + uoffset_beg;
    const char *end = reinterpret_cast<const char *>(text.data()) + t_end;
    const size_t len = (size_t)(end - base);
    const int T = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_threads, len / (1 << 20) + 1));
    std::vector<const char *> cut((size_t)T + 1);
    cut[0] = base; cut[(size_t)T] = end;
    for (int t = 1; t < T; ++t) {
        const char *c = base + len * (size_t)t / (size_t)T;
        const char *nl = (const char *)memchr(c, '\n', (size_t)(end - c));
        cut[(size_t)t] = nl ? nl + 1 : end;
    }
    std::vector<std::vector<Segment>> parts((size_t)T);
    std::vector<int64_t> skipped((size_t)T, 0);
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t] { parse_range(cut[(size_t)t], cut[(size_t)t + 1], bed6, parts[(size_t)t], skipped[(size_t)t]); });
        for (auto &x : th) x.join();
    }
    FragFile *ff = new FragFile();
    ff->bed6 = bed6;
    for (int64_t sk : skipped) ff->skipped += sk;
    for (auto &segs : parts)
        for (auto &sg : segs) {
            Columns *dst = nullptr;
            for (auto &c : ff->contigs) if (c.name == sg.name) { dst = &c; break; }
            if (!dst) { ff->contigs.emplace_back(); dst = &ff->contigs.back(); dst->name = sg.name; }
            dst->start.insert(dst->start.end(), sg.start.begin(), sg.start.end());
            dst->stop.insert(dst->stop.end(), sg.stop.begin(), sg.stop.end());
            dst->mapq.insert(dst->mapq.end(), sg.mapq.begin(), sg.mapq.end());
            dst->strand.insert(dst->strand.end(), sg.strand.begin(), sg.strand.end());
        }
    if (err) *err = FTK_OK;
    return ff;
}

extern "C" int32_t ftk_fragfile_is_bed6(void *h) { return h ? static_cast<FragFile *>(h)->bed6 : 0; }
extern "C" int64_t ftk_fragfile_skipped(void *h) { return h ? static_cast<FragFile *>(h)->skipped : 0; }
extern "C" int32_t ftk_fragfile_n_contigs(void *h) { return h ? (int32_t)static_cast<FragFile *>(h)->contigs.size() : 0; }
extern "C" const char *ftk_fragfile_contig_name(void *h, int32_t i) {
    FragFile *ff = static_cast<FragFile *>(h);
    return (ff && i >= 0 && (size_t)i < ff->contigs.size()) ? ff->contigs[(size_t)i].name.c_str() : "";
}
extern "C" int64_t ftk_fragfile_contig_count(void *h, int32_t i) {
    FragFile *ff = static_cast<FragFile *>(h);
    return (ff && i >= 0 && (size_t)i < ff->contigs.size()) ? (int64_t)ff->contigs[(size_t)i].start.size() : -1;
}
extern "C" int ftk_fragfile_copy(void *h, int32_t i, int32_t *start, int32_t *stop, uint8_t *mapq, uint8_t *strand) {
    FragFile *ff = static_cast<FragFile *>(h);
    if (!ff || i < 0 || (size_t)i >= ff->contigs.size() || !start || !stop || !mapq || !strand) return FTK_E_INVALID;
    const Columns &c = ff->contigs[(size_t)i];
    const size_t n = c.start.size();
    if (n) {
        memcpy(start, c.start.data(), n * 4); memcpy(stop, c.stop.data(), n * 4);
        memcpy(mapq, c.mapq.data(), n); memcpy(strand, c.strand.data(), n);
    }
    return FTK_OK;
}
extern "C" void ftk_fragfile_close(void *h) { delete static_cast<FragFile *>(h); }

// ---------------------------------------------------------------- bigWig section codec
// Sections are independent zlib streams, so a batch is an embarrassingly parallel loop; the
// threads take members round-robin (sections are near-uniform in size).
namespace {
template <typename F>
int run_members(int64_t n, int32_t n_threads, F &&one) {
    if (n_threads < 1) n_threads = (int32_t)std::max(1u, std::thread::hardware_concurrency());
    const int T = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads, n));
    std::vector<int> rc((size_t)T, 0);
    auto work = [&](int t) {
        for (int64_t i = t; i < n; i += T) {
            int r = one(i);
            if (r != 0) { rc[(size_t)t] = r; return; }
        }
    };
    if (T == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t) th.emplace_back(work, t);
        for (auto &x : th) x.join();
    }
    for (int r : rc) if (r != 0) return r;
    return 0;
}
}  // namespace

extern "C" int ftk_zlib_compress_batch(const uint8_t *in, const int64_t *in_off, int64_t n, int32_t level,
                                       int32_t n_threads, uint8_t *out, const int64_t *out_off,
                                       int64_t *out_size) {
    if (n < 0 || level < -1 || level > 9) return FTK_E_INVALID;
    if (n == 0) return 0;
    if (!in || !in_off || !out || !out_off || !out_size) return FTK_E_INVALID;
    return run_members(n, n_threads, [&](int64_t i) -> int {
        const int64_t len = in_off[i + 1] - in_off[i], cap = out_off[i + 1] - out_off[i];
        if (len < 0 || cap < 0 || (uLong)cap < compressBound((uLong)len)) return FTK_E_INVALID;
        uLongf got = (uLongf)cap;
        if (compress2(out + out_off[i], &got, in + in_off[i], (uLong)len, level) != Z_OK) return FTK_E_IO;
        out_size[i] = (int64_t)got;
        return 0;
    });
}

extern "C" int ftk_zlib_uncompress_batch(const uint8_t *in, const int64_t *in_off, const int64_t *in_size,
                                         int64_t n, int32_t n_threads, uint8_t *out, const int64_t *out_off,
                                         int64_t *out_size) {
    if (n < 0) return FTK_E_INVALID;
    if (n == 0) return 0;
    if (!in || !in_off || !in_size || !out || !out_off || !out_size) return FTK_E_INVALID;
    return run_members(n, n_threads, [&](int64_t i) -> int {
        const int64_t cap = out_off[i + 1] - out_off[i];
        if (in_size[i] < 0 || cap < 0) return FTK_E_INVALID;
        uLongf got = (uLongf)cap;
        if (uncompress(out + out_off[i], &got, in + in_off[i], (uLong)in_size[i]) != Z_OK) return FTK_E_IO;
        out_size[i] = (int64_t)got;
        return 0;
    });
}

// ---------------------------------------------------------------- bedGraph text + gzip members
// `contig\tpos\tpos+1\tscore\n` for consecutive positions (frag/_multi_wps.py:328-341 writes these
// lines one f-string at a time).  Two passes per thread slice: measure, then print.
namespace {
inline int dec_len(long long v) {
    int n = v < 0 ? 1 : 0;
    unsigned long long u = v < 0 ? 0ull - (unsigned long long)v : (unsigned long long)v;
    do { ++n; u /= 10; } while (u);
    return n;
}
inline char *put_dec(char *p, long long v) {
    char tmp[24];
    int n = 0;
    unsigned long long u = v < 0 ? 0ull - (unsigned long long)v : (unsigned long long)v;
    do { tmp[n++] = (char)('0' + u % 10); u /= 10; } while (u);
    if (v < 0) *p++ = '-';
    while (n) *p++ = tmp[--n];
    return p;
}
}  // namespace

extern "C" int64_t ftk_format_bedgraph_i64(const char *contig, int64_t start, const int64_t *scores, int64_t n,
                                           int32_t n_threads, char *out, int64_t out_cap) {
    if (!contig || n < 0 || (n > 0 && !scores)) return FTK_E_INVALID;
    if (n == 0) return 0;
    const size_t cl = strlen(contig);
    if (n_threads < 1) n_threads = (int32_t)std::max(1u, std::thread::hardware_concurrency());
    const int T = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads, n / 65536 + 1));
    std::vector<int64_t> bytes((size_t)T + 1, 0);
    auto slice = [&](int t, int64_t &a, int64_t &b) { a = n * t / T; b = n * (t + 1) / T; };
    auto measure = [&](int t) {
        int64_t a, b; slice(t, a, b);
        int64_t tot = 0;
        for (int64_t i = a; i < b; ++i)
            tot += (int64_t)cl + 4 + dec_len(start + i) + dec_len(start + i + 1) + dec_len(scores[i]);
        bytes[(size_t)t + 1] = tot;
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < T; ++t) th.emplace_back(measure, t);
        measure(0);
        for (auto &x : th) x.join();
    }
    for (int t = 0; t < T; ++t) bytes[(size_t)t + 1] += bytes[(size_t)t];
    const int64_t total = bytes[(size_t)T];
    if (!out) return total;                     // size query
    if (out_cap < total) return FTK_E_INVALID;
    auto print = [&](int t) {
        int64_t a, b; slice(t, a, b);
        char *p = out + bytes[(size_t)t];
        for (int64_t i = a; i < b; ++i) {
            memcpy(p, contig, cl); p += cl;
            *p++ = '\t'; p = put_dec(p, start + i);
            *p++ = '\t'; p = put_dec(p, start + i + 1);
            *p++ = '\t'; p = put_dec(p, scores[i]);
            *p++ = '\n';
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < T; ++t) th.emplace_back(print, t);
        print(0);
        for (auto &x : th) x.join();
    }
    return total;
}

// gzip members (RFC 1952) of independent chunks: the concatenation is one valid .gz file that gzip,
// zcat and Python's gzip read as a single stream - the multi-threaded stand-in for gzip.open(..., "wt").
extern "C" int ftk_gzip_compress_batch(const uint8_t *in, const int64_t *in_off, int64_t n, int32_t level,
                                       int32_t n_threads, uint8_t *out, const int64_t *out_off,
                                       int64_t *out_size) {
    if (n < 0 || level < -1 || level > 9) return FTK_E_INVALID;
    if (n == 0) return 0;
    if (!in || !in_off || !out || !out_off || !out_size) return FTK_E_INVALID;
    return run_members(n, n_threads, [&](int64_t i) -> int {
        const int64_t len = in_off[i + 1] - in_off[i], cap = out_off[i + 1] - out_off[i];
        if (len < 0 || cap < 0) return FTK_E_INVALID;
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (deflateInit2(&zs, level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) return FTK_E_IO;
        if ((uLong)cap < deflateBound(&zs, (uLong)len)) { deflateEnd(&zs); return FTK_E_INVALID; }
        zs.next_in = const_cast<uint8_t *>(in + in_off[i]); zs.avail_in = (uInt)len;
        zs.next_out = out + out_off[i]; zs.avail_out = (uInt)cap;
        const int rc = deflate(&zs, Z_FINISH);
        out_size[i] = (int64_t)zs.total_out;
        deflateEnd(&zs);
        return rc == Z_STREAM_END ? 0 : FTK_E_IO;
    });
}
