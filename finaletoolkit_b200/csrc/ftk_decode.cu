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
    std::vector<int32_t> r1_start, r1_end;   // BAM only: reference span of read 1 (what an indexed fetch tests)
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

// One BGZF member header at b[p...] (size bytes in the buffer).  Returns 1 and fills xend (first
// byte of the deflate stream), end (one past the member) and usize (ISIZE) for a well-formed,
// complete member; 0 when more bytes are needed; -1 for anything malformed: not a BGZF member, a
// subfield that overruns the extra area, BSIZE smaller than header + trailer (would underflow the
// compressed size), ISIZE beyond the 64 KiB a BGZF member may hold.
int bgzf_member(const unsigned char *b, size_t size, size_t p, size_t &xend, size_t &end, size_t &usize) {
    if (p + 18 > size) return 0;
    if (b[p] != 31 || b[p + 1] != 139 || b[p + 2] != 8 || !(b[p + 3] & 4)) return -1;
    const unsigned xlen = b[p + 10] | (b[p + 11] << 8);
    size_t q = p + 12;
    xend = q + xlen;
    if (xend > size) return 0;
    long bsize = -1;
    while (q + 4 <= xend) {
        const unsigned slen = b[q + 2] | (b[q + 3] << 8);
        if (q + 4 + slen > xend) return -1;
        if (b[q] == 'B' && b[q + 1] == 'C' && slen == 2) bsize = (b[q + 4] | (b[q + 5] << 8)) + 1;
        q += 4 + slen;
    }
    if (bsize < 0 || (size_t)bsize < (xend - p) + 8) return -1;
    end = p + (size_t)bsize;
    if (end > size) return 0;
    usize = (size_t)b[end - 4] | ((size_t)b[end - 3] << 8) | ((size_t)b[end - 2] << 16) | ((size_t)b[end - 1] << 24);
    if (usize > 65536) return -1;
    return 1;
}

// BGZF block list, or empty if the file is not BGZF (plain gzip)
std::vector<Block> scan_bgzf(const std::vector<unsigned char> &b) {
    std::vector<Block> blocks;
    size_t p = 0, uoff = 0;
    while (p + 18 <= b.size()) {
        size_t xend, end, usize;
        if (bgzf_member(b.data(), b.size(), p, xend, end, usize) != 1) return {};
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
        size_t xend, end, usize;
        if (bgzf_member(raw.data(), raw.size(), p, xend, end, usize) != 1) return fail(FTK_E_IO);
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

// ---------------------------------------------------------------- BAM -> fragment columns
// The reference reconstructs fragments from BAM reads in AlignmentWrapper._fetch_sam
// (io/alignment.py:242-268) behind the flag filter of _read_is_low_quality (:60-71, the mapq test is
// left to the kernels): keep reads that are paired, proper pair, mapped with a mapped mate, primary,
// not duplicate / QC-fail / supplementary and not read2; tlen > 0 -> [pos, pos + tlen), tlen < 0 ->
// [reference_end + tlen, reference_end), tlen == 0 dropped; mapq and strand are the read's own.
// The file is streamed: BGZF blocks are read in batches, inflated in parallel, and the BAM records
// (which ignore block boundaries) are walked in order with the unfinished tail carried over.
namespace {
struct BamState {
    bool header_done = false;
    std::vector<std::string> ref_names;
    std::vector<int64_t> ref_lens;
    std::vector<int> ref_to_col;       // refID -> index in FragFile::contigs, -1 = not seen yet
};

inline int32_t rd_i32(const unsigned char *p) { int32_t v; memcpy(&v, p, 4); return v; }
inline uint32_t rd_u32(const unsigned char *p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline uint16_t rd_u16(const unsigned char *p) { uint16_t v; memcpy(&v, p, 2); return v; }

// Consume as much of [p, end) as is complete; returns the first unconsumed byte, or nullptr on a
// malformed stream.
const unsigned char *bam_consume(const unsigned char *p, const unsigned char *end, BamState &st, FragFile &ff) {
    if (!st.header_done) {
        if (end - p < 12) return p;
        if (memcmp(p, "BAM\1", 4) != 0) return nullptr;
        const int64_t l_text = rd_i32(p + 4);
        if (l_text < 0) return nullptr;
        if (end - p < 12 + l_text) return p;                      // bound first, then form the pointer
        const unsigned char *q = p + 8 + l_text;
        const int64_t n_ref = rd_i32(q);
        if (n_ref < 0) return nullptr;
        q += 4;
        std::vector<std::string> names;
        std::vector<int64_t> lens;
        for (int64_t r = 0; r < n_ref; ++r) {
            if (end - q < 4) return p;
            const int64_t l_name = rd_i32(q);
            if (l_name < 1) return nullptr;
            if (end - q < 8 + l_name) return p;
            names.emplace_back(reinterpret_cast<const char *>(q + 4), (size_t)l_name - 1);
            lens.push_back(rd_i32(q + 4 + l_name));
            q += 8 + l_name;
        }
        st.ref_names.swap(names); st.ref_lens.swap(lens);
        st.ref_to_col.assign((size_t)n_ref, -1);
        st.header_done = true;
        p = q;
    }
    while (end - p >= 4) {
        const int64_t bs = rd_i32(p);
        if (bs < 32) return nullptr;
        if (end - p < 4 + bs) break;
        const unsigned char *r = p + 4;
        const int32_t ref_id = rd_i32(r), pos = rd_i32(r + 4);
        const unsigned l_read_name = r[8], mapq = r[9];
        const unsigned n_cigar = rd_u16(r + 12), flag = rd_u16(r + 14);
        const int32_t tlen = rd_i32(r + 28);
        p += 4 + bs;
        // paired, proper pair; not unmapped / mate unmapped / read2 / secondary / QC fail / duplicate / supplementary
        if ((flag & 0x3u) != 0x3u || (flag & (0x4u | 0x8u | 0x80u | 0x100u | 0x200u | 0x400u | 0x800u))) continue;
        if (tlen == 0 || ref_id < 0 || (size_t)ref_id >= st.ref_names.size()) continue;
        // reference span of the read itself: htslib's region test is `pos < stop && bam_endpos > start`,
        // bam_endpos = pos + (reference bases the CIGAR consumes, 1 when there are none)
        const bool cigar_ok = n_cigar != 0 && 32 + (int64_t)l_read_name + 4 * (int64_t)n_cigar <= bs;
        int64_t ref_end = pos;
        if (cigar_ok) {
            const unsigned char *c = r + 32 + l_read_name;
            for (unsigned k = 0; k < n_cigar; ++k) {
                const uint32_t op = rd_u32(c + 4 * k);
                const unsigned code = op & 15u;   // M, D, N, =, X consume the reference
                if (code == 0 || code == 2 || code == 3 || code == 7 || code == 8) ref_end += op >> 4;
            }
        }
        int64_t fs, fe;
        if (tlen > 0) {
            fs = pos; fe = (int64_t)pos + tlen;
        } else {
            if (!cigar_ok) continue;   // pysam: reference_end is None without an alignment
            fs = ref_end + tlen; fe = ref_end;
        }
        const int64_t r1e = std::min<int64_t>(ref_end > pos ? ref_end : (int64_t)pos + 1, INT32_MAX);
        if (fs < INT32_MIN || fs > INT32_MAX || fe < INT32_MIN || fe > INT32_MAX) continue;
        int &col = st.ref_to_col[(size_t)ref_id];
        if (col < 0) {
            col = (int)ff.contigs.size();
            ff.contigs.emplace_back();
            ff.contigs.back().name = st.ref_names[(size_t)ref_id];
        }
        Columns &dst = ff.contigs[(size_t)col];
        dst.start.push_back((int32_t)fs); dst.stop.push_back((int32_t)fe);
        dst.mapq.push_back((uint8_t)mapq); dst.strand.push_back((flag & 0x10u) ? 0 : 1);
        dst.r1_start.push_back(pos); dst.r1_end.push_back((int32_t)r1e);
    }
    return p;
}
}  // namespace

struct BamFile { FragFile frags; std::vector<std::string> ref_names; std::vector<int64_t> ref_lens; };

extern "C" void *ftk_bamfile_open(const char *path, int32_t n_threads, int32_t *err) {
    auto fail = [&](int code) -> void * { if (err) *err = code; return nullptr; };
    if (!path) return fail(FTK_E_INVALID);
    FILE *fh = fopen(path, "rb");
    if (!fh) return fail(FTK_E_IO);
    if (n_threads < 1) n_threads = (int32_t)std::max(1u, std::thread::hardware_concurrency());
    BamFile *bf = new BamFile();
    BamState st;
    std::vector<unsigned char> cbuf, ubuf;           // compressed carry + batch, inflated carry + batch
    size_t ufill = 0;
    const size_t kRead = (size_t)64 << 20;
    bool eof = false, ok = true;
    while (ok && !eof) {
        const size_t have = cbuf.size();
        cbuf.resize(have + kRead);
        const size_t got = fread(cbuf.data() + have, 1, kRead, fh);
        cbuf.resize(have + got);
        eof = got < kRead;
        // complete BGZF blocks in cbuf
        std::vector<Block> blocks;
        size_t p = 0, uoff = 0;
        while (p + 18 <= cbuf.size()) {
            size_t xend, end, usize;
            const int st_ = bgzf_member(cbuf.data(), cbuf.size(), p, xend, end, usize);
            if (st_ < 0) { ok = false; break; }
            if (st_ == 0) break;                                  // incomplete member: read more
            blocks.push_back({xend, end - 8 - xend, usize, uoff});
            uoff += usize;
            p = end;
        }
        if (!ok) break;
        if (eof && p != cbuf.size()) { ok = false; break; }      // trailing garbage / truncated block
        ubuf.resize(ufill + uoff);
        {
            const int T = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_threads, blocks.size()));
            std::vector<std::thread> th;
            std::vector<int> good((size_t)T, 1);
            for (int t = 0; t < T; ++t)
                th.emplace_back([&, t] {
                    for (size_t i = (size_t)t; i < blocks.size(); i += (size_t)T) {
                        const Block &b = blocks[i];
                        if (b.usize && !inflate_raw(cbuf.data() + b.off, b.csize, ubuf.data() + ufill + b.uoff, b.usize))
                            good[(size_t)t] = 0;
                    }
                });
            for (auto &x : th) x.join();
            for (int v : good) if (!v) ok = false;
        }
        if (!ok) break;
        cbuf.erase(cbuf.begin(), cbuf.begin() + (ptrdiff_t)p);
        const unsigned char *base = ubuf.data(), *endp = ubuf.data() + ubuf.size();
        const unsigned char *next = bam_consume(base, endp, st, bf->frags);
        if (!next) { ok = false; break; }
        ufill = (size_t)(endp - next);
        memmove(ubuf.data(), next, ufill);
        ubuf.resize(ufill);
    }
    fclose(fh);
    if (!ok || !st.header_done || ufill != 0) { delete bf; return fail(FTK_E_IO); }
    bf->ref_names = st.ref_names; bf->ref_lens = st.ref_lens;
    if (err) *err = FTK_OK;
    return bf;
}
extern "C" void *ftk_bamfile_fragments(void *h) { return h ? &static_cast<BamFile *>(h)->frags : nullptr; }
extern "C" int32_t ftk_bamfile_n_refs(void *h) { return h ? (int32_t)static_cast<BamFile *>(h)->ref_names.size() : 0; }
extern "C" const char *ftk_bamfile_ref_name(void *h, int32_t i) {
    BamFile *b = static_cast<BamFile *>(h);
    return (b && i >= 0 && (size_t)i < b->ref_names.size()) ? b->ref_names[(size_t)i].c_str() : "";
}
extern "C" int64_t ftk_bamfile_ref_length(void *h, int32_t i) {
    BamFile *b = static_cast<BamFile *>(h);
    return (b && i >= 0 && (size_t)i < b->ref_lens.size()) ? b->ref_lens[(size_t)i] : -1;
}
extern "C" void ftk_bamfile_close(void *h) { delete static_cast<BamFile *>(h); }

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
// read-1 reference spans of a BAM-derived contig (same row order as ftk_fragfile_copy); 1 = the handle has
// none (fragment files), 0 = copied
extern "C" int ftk_fragfile_copy_read1(void *h, int32_t i, int32_t *r1_start, int32_t *r1_end) {
    FragFile *ff = static_cast<FragFile *>(h);
    if (!ff || i < 0 || (size_t)i >= ff->contigs.size() || !r1_start || !r1_end) return FTK_E_INVALID;
    const Columns &c = ff->contigs[(size_t)i];
    const size_t n = c.start.size();
    if (c.r1_start.size() != n || c.r1_end.size() != n) return 1;
    if (n) { memcpy(r1_start, c.r1_start.data(), n * 4); memcpy(r1_end, c.r1_end.data(), n * 4); }
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
