// Packed fragment columns: the wire format between the host decoder and HBM.
//
// The reference streams text rows contig/start/stop/mapq/strand per interval
// (io/alignment.py:270-302, utils/_frag_generator.py:124-130).  Here a contig's start-sorted
// fragments travel host -> device ONCE, and PCIe is the end-to-end limit, so the
// decoder hands the columns over in 4 bytes per fragment instead of 10:
//
//   word[i] (uint32) =  dstart | length << 11 | strand << 23 | mapq << 24
//     dstart = start[i] - start[i-1]   (11 bits; 0 for the first fragment of a block)
//     length = stop[i] - start[i]      (12 bits)
//   anchor[b] (int32), one per block of FTK_PACK_BLOCK = 64 fragments:
//     >= 0 : absolute start of the block's first fragment
//     <  0 : "raw" block number -1 - anchor[b]: some fragment of the block does not fit the
//            fields (gap >= 2048 bp, length >= 4096 or negative, negative start, unsorted rows);
//            its 64 rows are stored verbatim in the raw side columns and the words are ignored.
// 4.0625 B per fragment when nothing escapes (chr1 at 30x: 80 M fragments -> 325 MB instead of
// 720-800 MB); escapes cost 10 B per fragment of the affected block only.  Lossless: unpacking
// reproduces start / stop / mapq / strand bit for bit (tests/test_gpu_pack.py).
//
// record_bytes = 3 is the narrow variant of the same scheme for deep, short-fragment data (cfDNA at
// >= ~5x): 24-bit records  dstart (6 bits) | length << 6 (9 bits) | strand << 15 | mapq << 16,
// 48 words per block = 3.0625 B per fragment (chr1 at 30x: 245 MB).  A block with a gap >= 64 bp or
// a fragment >= 512 bp escapes to the raw columns exactly like above; the packer reports how many
// do, so the caller can pick the width that puts fewer bytes on the wire (packed.py does).
//
// ftk_pack_fragments_host : host, multi-threaded (the decoder side)
// ftk_unpack_fragments    : device; one warp per block, two fragments per lane, shuffle prefix sum
//                           of the start deltas, coalesced 8-byte stores.  HBM-bound:
//                           4.06 B read + 9-10 B written per fragment.
#include <algorithm>
#include <thread>
#include <vector>

#include "ftk_common.cuh"

namespace ftk {

constexpr int kPackBlock = FTK_PACK_BLOCK;
static_assert(kPackBlock == 64, "one warp unpacks a block as 32 lanes x 2 fragments");
constexpr uint32_t kDeltaBits = 11, kLenBits = 12;
constexpr uint32_t kDeltaMax = (1u << kDeltaBits) - 1, kLenMax = (1u << kLenBits) - 1;
constexpr uint32_t kDeltaBits3 = 6, kLenBits3 = 9;
constexpr uint32_t kDeltaMax3 = (1u << kDeltaBits3) - 1, kLenMax3 = (1u << kLenBits3) - 1;
constexpr int kWordsPerBlock3 = kPackBlock * 3 / 4;      // 48

template <int RB>
__global__ void __launch_bounds__(256)
unpack_fragments_kernel(const uint2 *__restrict__ words, const int32_t *__restrict__ anchors,
                        const int32_t *__restrict__ raw_start, const int32_t *__restrict__ raw_stop,
                        const uint8_t *__restrict__ raw_mapq, const uint8_t *__restrict__ raw_strand,
                        int64_t n, int64_t n_blocks, int32_t *__restrict__ start, int32_t *__restrict__ stop,
                        uint8_t *__restrict__ mapq, uint8_t *__restrict__ strand) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b = warp; b < n_blocks; b += n_warps) {
        const int anchor = __ldg(anchors + b);
        const int64_t i0 = b * kPackBlock + 2 * lane;      // this lane's two fragments
        int s0, s1, e0, e1;
        unsigned q0, q1, d0, d1;
        if (anchor >= 0 && RB == 3) {
            // the lane's two 24-bit records are bytes [6 lane, 6 lane + 6) of the block: two words
            const uint32_t *__restrict__ w32 = reinterpret_cast<const uint32_t *>(words) + b * kWordsPerBlock3;
            const int wi = (6 * lane) >> 2;
            const unsigned long long both = ((unsigned long long)__ldcs(w32 + wi + 1) << 32) | __ldcs(w32 + wi);
            const unsigned long long v = both >> (((6 * lane) & 3) * 8);
            const uint32_t r0 = (uint32_t)v & 0xffffffu, r1 = (uint32_t)(v >> 24) & 0xffffffu;
            const int dx = (int)(r0 & kDeltaMax3), dy = (int)(r1 & kDeltaMax3);
            int t = dx + dy;                                 // inclusive scan of the pair sums
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, t, off);
                if (lane >= off) t += u;
            }
            s1 = anchor + t;
            s0 = s1 - dy;
            e0 = s0 + (int)((r0 >> kDeltaBits3) & kLenMax3);
            e1 = s1 + (int)((r1 >> kDeltaBits3) & kLenMax3);
            d0 = (r0 >> 15) & 1u; d1 = (r1 >> 15) & 1u;
            q0 = r0 >> 16; q1 = r1 >> 16;
        } else if (anchor >= 0) {
            const uint2 w = __ldcs(words + b * (kPackBlock / 2) + lane);
            const int dx = (int)(w.x & kDeltaMax), dy = (int)(w.y & kDeltaMax);
            int t = dx + dy;                                 // inclusive scan of the pair sums
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, t, off);
                if (lane >= off) t += v;
            }
            s1 = anchor + t;
            s0 = s1 - dy;
            e0 = s0 + (int)((w.x >> kDeltaBits) & kLenMax);
            e1 = s1 + (int)((w.y >> kDeltaBits) & kLenMax);
            d0 = (w.x >> 23) & 1u; d1 = (w.y >> 23) & 1u;
            q0 = w.x >> 24; q1 = w.y >> 24;
        } else {
            const int64_t r = ((int64_t)(-1 - anchor)) * kPackBlock + 2 * lane;
            s0 = raw_start[r]; s1 = raw_start[r + 1];
            e0 = raw_stop[r]; e1 = raw_stop[r + 1];
            q0 = raw_mapq[r]; q1 = raw_mapq[r + 1];
            d0 = raw_strand ? raw_strand[r] : 0u; d1 = raw_strand ? raw_strand[r + 1] : 0u;
        }
        if (i0 + 1 < n) {                                   // outputs are 8-byte aligned: i0 is even
            *reinterpret_cast<int2 *>(start + i0) = make_int2(s0, s1);
            *reinterpret_cast<int2 *>(stop + i0) = make_int2(e0, e1);
            if (mapq) *reinterpret_cast<uchar2 *>(mapq + i0) = make_uchar2((unsigned char)q0, (unsigned char)q1);
            if (strand) *reinterpret_cast<uchar2 *>(strand + i0) = make_uchar2((unsigned char)d0, (unsigned char)d1);
        } else if (i0 < n) {
            start[i0] = s0; stop[i0] = e0;
            if (mapq) mapq[i0] = (uint8_t)q0;
            if (strand) strand[i0] = (uint8_t)d0;
        }
    }
}

// true when the 64 rows [i0, i1) fit the packed fields
static bool block_fits(const int32_t *start, const int32_t *stop, int64_t i0, int64_t i1, int64_t len_max,
                       int64_t delta_max) {
    if (start[i0] < 0) return false;
    for (int64_t i = i0; i < i1; ++i) {
        const int64_t len = (int64_t)stop[i] - start[i];
        if (len < 0 || len > len_max) return false;
        if (i > i0) {
            const int64_t d = (int64_t)start[i] - start[i - 1];
            if (d < 0 || d > delta_max) return false;
        }
    }
    return true;
}

}  // namespace ftk

using namespace ftk;

extern "C" int64_t ftk_pack_fragments_host(const int32_t *start, const int32_t *stop, const uint8_t *mapq,
                                           const uint8_t *strand, int64_t n, int32_t threads,
                                           uint32_t *words, int32_t *anchors,
                                           int32_t *raw_start, int32_t *raw_stop, uint8_t *raw_mapq,
                                           uint8_t *raw_strand, int64_t raw_capacity_blocks, int32_t record_bytes) {
    if (n < 0 || (n > 0 && (!start || !stop))) return FTK_E_INVALID;
    if (record_bytes != 3 && record_bytes != 4) return FTK_E_INVALID;
    const bool narrow = record_bytes == 3;
    const int64_t len_max = narrow ? kLenMax3 : kLenMax, delta_max = narrow ? kDeltaMax3 : kDeltaMax;
    const int64_t n_blocks = (n + kPackBlock - 1) / kPackBlock;
    if (n_blocks == 0) return 0;
    int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    nt = (int)std::max<int64_t>(1, std::min<int64_t>(nt, (n_blocks + 4095) / 4096));
    // pass 1: which blocks escape (one flag per block), in parallel
    std::vector<uint8_t> is_raw((size_t)n_blocks);
    auto for_blocks = [&](auto &&fn) {
        std::vector<std::thread> pool;
        for (int t = 0; t < nt; ++t)
            pool.emplace_back([&, t] {
                const int64_t b0 = n_blocks * t / nt, b1 = n_blocks * (t + 1) / nt;
                fn(b0, b1);
            });
        for (auto &th : pool) th.join();
    };
    for_blocks([&](int64_t b0, int64_t b1) {
        for (int64_t b = b0; b < b1; ++b) {
            const int64_t i0 = b * kPackBlock, i1 = std::min<int64_t>(n, i0 + kPackBlock);
            is_raw[(size_t)b] = block_fits(start, stop, i0, i1, len_max, delta_max) ? 0 : 1;
        }
    });
    std::vector<int64_t> raw_index((size_t)n_blocks);
    int64_t n_raw = 0;
    for (int64_t b = 0; b < n_blocks; ++b) { raw_index[(size_t)b] = n_raw; n_raw += is_raw[(size_t)b]; }
    if (!words || !anchors) return n_raw;                       // count-only call
    if (n_raw > raw_capacity_blocks) return FTK_E_RANGE;
    if (n_raw > 0 && (!raw_start || !raw_stop || !raw_mapq)) return FTK_E_INVALID;
    if (n_raw > INT32_MAX - 1) return FTK_E_RANGE;
    // pass 2: fill
    for_blocks([&](int64_t b0, int64_t b1) {
        for (int64_t b = b0; b < b1; ++b) {
            const int64_t i0 = b * kPackBlock, i1 = std::min<int64_t>(n, i0 + kPackBlock);
            if (narrow) {
                uint8_t *w8 = reinterpret_cast<uint8_t *>(words) + i0 * 3;
                std::fill(w8, w8 + kPackBlock * 3, (uint8_t)0);
                if (!is_raw[(size_t)b]) {
                    anchors[b] = start[i0];
                    for (int64_t i = i0; i < i1; ++i) {
                        const uint32_t d = (i > i0) ? (uint32_t)(start[i] - start[i - 1]) : 0u;
                        const uint32_t len = (uint32_t)(stop[i] - start[i]);
                        const uint32_t q = mapq ? mapq[i] : 255u;
                        const uint32_t sd = strand ? (strand[i] & 1u) : 0u;
                        const uint32_t rec = d | (len << kDeltaBits3) | (sd << 15) | (q << 16);
                        uint8_t *r8 = w8 + (i - i0) * 3;
                        r8[0] = (uint8_t)rec; r8[1] = (uint8_t)(rec >> 8); r8[2] = (uint8_t)(rec >> 16);
                    }
                    continue;
                }
            }
            uint32_t *w = narrow ? nullptr : words + i0;
            if (!is_raw[(size_t)b]) {
                anchors[b] = start[i0];
                for (int64_t i = i0; i < i1; ++i) {
                    const uint32_t d = (i > i0) ? (uint32_t)(start[i] - start[i - 1]) : 0u;
                    const uint32_t len = (uint32_t)(stop[i] - start[i]);
                    const uint32_t q = mapq ? mapq[i] : 255u;
                    const uint32_t sd = strand ? (strand[i] & 1u) : 0u;
                    w[i - i0] = d | (len << kDeltaBits) | (sd << 23) | (q << 24);
                }
                for (int64_t i = i1; i < i0 + kPackBlock; ++i) w[i - i0] = 0u;
            } else {
                const int64_t r = raw_index[(size_t)b];
                anchors[b] = (int32_t)(-1 - r);
                for (int64_t i = i0; i < i0 + kPackBlock; ++i) {
                    const int64_t k = r * kPackBlock + (i - i0);
                    if (w) w[i - i0] = 0u;
                    const bool live = i < i1;
                    raw_start[k] = live ? start[i] : 0;
                    raw_stop[k] = live ? stop[i] : 0;
                    raw_mapq[k] = live ? (mapq ? mapq[i] : 255) : 0;
                    if (raw_strand) raw_strand[k] = live ? (strand ? strand[i] : 0) : 0;
                }
            }
        }
    });
    return n_raw;
}

extern "C" int ftk_unpack_fragments(const uint32_t *words, const int32_t *anchors,
                                    const int32_t *raw_start, const int32_t *raw_stop,
                                    const uint8_t *raw_mapq, const uint8_t *raw_strand, int64_t n_raw_blocks,
                                    int64_t n, int32_t *start, int32_t *stop, uint8_t *mapq, uint8_t *strand,
                                    int32_t record_bytes, ftk_stream_t stream_) {
    if (n == 0) return FTK_OK;
    if (record_bytes != 3 && record_bytes != 4) return FTK_E_INVALID;
    if (n < 0 || n_raw_blocks < 0 || !words || !anchors || !start || !stop) return FTK_E_INVALID;
    if (n_raw_blocks > 0 && (!raw_start || !raw_stop || !raw_mapq)) return FTK_E_INVALID;
    // two-fragment stores need 8-byte aligned int32 outputs and 2-byte aligned byte outputs
    if ((reinterpret_cast<uintptr_t>(start) & 7) || (reinterpret_cast<uintptr_t>(stop) & 7) ||
        (reinterpret_cast<uintptr_t>(words) & 7) || (reinterpret_cast<uintptr_t>(mapq) & 1) ||
        (reinterpret_cast<uintptr_t>(strand) & 1))
        return FTK_E_INVALID;
    const int64_t n_blocks = (n + kPackBlock - 1) / kPackBlock;
    const int64_t ctas = (n_blocks + 7) / 8;                   // 8 warps = 8 blocks per CTA
    const unsigned grid = (unsigned)std::min<int64_t>(ctas, (int64_t)kNumSMs * 32);
    if (record_bytes == 3)
        unpack_fragments_kernel<3><<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
            reinterpret_cast<const uint2 *>(words), anchors, raw_start, raw_stop, raw_mapq, raw_strand, n, n_blocks,
            start, stop, mapq, strand);
    else
        unpack_fragments_kernel<4><<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
            reinterpret_cast<const uint2 *>(words), anchors, raw_start, raw_stop, raw_mapq, raw_strand, n, n_blocks,
            start, stop, mapq, strand);
    FTK_CHECK_LAUNCH("unpack_fragments_kernel");
    return FTK_OK;
}
