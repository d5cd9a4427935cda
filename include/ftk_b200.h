/*
 * ftk_b200.h - C ABI of the B200-native FinaleToolkit hot path.
 *
 * FinaleToolkit (the reference, /root/reference/src/finaletoolkit) has no FFI:
 * its seam is the Python function layer.  This header is the array-level seam
 * a maintainer would bind (ctypes) in place of the
 * reference's per-interval Python/numba loops.  Each entry point cites the
 * reference code it replaces.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch types.
 *   - `*_dev` pointers are DEVICE pointers; `*_host` are host pointers.
 *   - All device entry points are asynchronous on `stream` (a cudaStream_t
 *     passed as void*), allocate nothing, and keep no global state.
 *   - Return 0 on success, a negative FTK_E_* code otherwise
 *     (ftk_error_string() describes it).  No exceptions cross the ABI.
 *   - Fragments of ONE contig, sorted ascending by start (the order of a
 *     tabix-indexed .frag.gz), as columns: start,stop int32; mapq,strand uint8.
 *     Coordinates are 0-based half-open and must fit int32.
 *   - FTK_NONE stands for Python None (unbounded) in length / region bounds.
 */
#ifndef FTK_B200_H
#define FTK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FTK_ABI_VERSION 1
#define FTK_NONE INT32_MIN

#define FTK_OK 0
#define FTK_E_INVALID -1   /* bad argument (null pointer, negative size, ...) */
#define FTK_E_CUDA -2      /* a CUDA runtime call or launch failed            */
#define FTK_E_RANGE -3     /* a size exceeds what the kernel supports         */

#define FTK_POLICY_MIDPOINT 0
#define FTK_POLICY_ANY 1

/* Positions per WPS tile (one CTA).  Intervals longer than this are split. */
#define FTK_WPS_TILE 5119 /* 5120 smem slots minus one guard slot for odd windows */

typedef void *ftk_stream_t; /* cudaStream_t */

int ftk_abi_version(void);
const char *ftk_error_string(int code);
/* last CUDA error text recorded by this library on the calling thread ("" if none) */
const char *ftk_last_cuda_error(void);

/* ------------------------------------------------------------------ WPS
 * Replaces the per-position loop of wps() - frag/_wps.py:176-188 calling the
 * numba kernel _single_nt_wps (frag/_wps.py:25-53) - and the per-interval
 * fragment selection frag_array(..., start=minimum, stop=maximum) of
 * frag/_wps.py:156-169 (mapq filter io/alignment.py:291; inclusive length
 * filter and midpoint policy utils/_frag_generator.py:117-123).
 *
 * Work unit = tile: positions [p0, p0+len) of one interval, len <= FTK_WPS_TILE.
 * A fragment takes part in a tile iff mapq >= min_mapq, min_len <= L <= max_len
 * and mid_lo <= (start+stop)/2 < mid_hi, where [mid_lo, mid_hi) is the
 * interval's padded fetch window [max(S-max_len,0), min(E+max_len,chrom_size)).
 * out[tile_out_off + (c - p0)] = WPS at position c, int32 (the reference's
 * int64 column; abs(WPS) <= local depth).
 */

/* Host helper: expand intervals into tiles.  Pass NULL outputs to count only.
 * Returns the number of tiles (>= 0) or a negative error code.
 * ivl_out_off[k] = offset of interval k's first position in `out`. */
int64_t ftk_wps_plan_tiles(const int64_t *ivl_start, const int64_t *ivl_stop,
                           const int64_t *ivl_out_off, int64_t n_ivl,
                           int64_t chrom_size, int32_t max_len,
                           int32_t *tile_p0, int32_t *tile_len,
                           int32_t *tile_mid_lo, int32_t *tile_mid_hi,
                           int64_t *tile_out_off);

/* scratch_dev: int64[2 * n_tiles] (per-tile fragment index range). */
int ftk_wps_tiles_i32(const int32_t *frag_start_dev, const int32_t *frag_stop_dev,
                      const uint8_t *frag_mapq_dev /* NULL = no mapq filter */, int64_t n_frag,
                      const int32_t *tile_p0_dev, const int32_t *tile_len_dev,
                      const int32_t *tile_mid_lo_dev, const int32_t *tile_mid_hi_dev,
                      const int64_t *tile_out_off_dev, int64_t n_tiles,
                      int32_t window_size, int32_t min_len, int32_t max_len, int32_t min_mapq,
                      int64_t *scratch_dev, int32_t *out_dev, ftk_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FTK_B200_H */
