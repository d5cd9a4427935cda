// Shared helpers for the sm_100a kernels of libftk_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ftk_b200.h"

namespace ftk {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// Record-and-translate a CUDA error; implemented in ftk_api.cu.
int cuda_fail(cudaError_t e, const char *what);

#define FTK_CUDA_TRY(expr)                                      \
    do {                                                        \
        cudaError_t _e = (expr);                                \
        if (_e != cudaSuccess) return ::ftk::cuda_fail(_e, #expr); \
    } while (0)

#define FTK_CHECK_LAUNCH(name)                                   \
    do {                                                         \
        cudaError_t _e = cudaGetLastError();                     \
        if (_e != cudaSuccess) return ::ftk::cuda_fail(_e, name); \
    } while (0)

// Streaming (read-once) global loads: keep them out of L1.
__device__ __forceinline__ int ld_stream(const int32_t *p) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int4 ld_stream4(const int4 *p) {
    int4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ unsigned ld_stream_u8(const uint8_t *p) {
    unsigned v;
    asm volatile("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream4(int4 *p, const int4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Shared-memory atomics through an explicit 32-bit shared-window address.  For atomicAdd(&smem[i], v) on
// a static __shared__ array nvcc (12.9, sm_100a) re-derives the array's window address next to every
// atomic in an unrolled loop (S2R SR_CgaCtaId + LEA per atomic); computing it once, hiding it behind an
// opaque move and issuing red.shared on it keeps the hot loops free of special-register reads.
__device__ __forceinline__ uint32_t smem_addr_once(const void *p) {
    const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
    uint32_t r;
    asm volatile("mov.b32 %0, %1;" : "=r"(r) : "r"(a));
    return r;
}
__device__ __forceinline__ void red_shared_add(uint32_t addr, int v) {
    asm volatile("red.shared.add.s32 [%0], %1;" :: "r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void red_shared_inc(uint32_t addr) {
    asm volatile("red.shared.add.u32 [%0], 1;" :: "r"(addr) : "memory");
}
__device__ __forceinline__ void red_shared_min(uint32_t addr, int v) {
    asm volatile("red.shared.min.s32 [%0], %1;" :: "r"(addr), "r"(v) : "memory");
}

// first index i in [0, n) with a[i] >= key (a ascending); n if none
__device__ __forceinline__ int64_t lower_bound(const int32_t *__restrict__ a, int64_t n, int64_t key) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = lo + ((hi - lo) >> 1);
        if ((int64_t)__ldg(a + mid) < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// The reference's fragment predicate (utils/_frag_generator.py:117-123,
// io/alignment.py:291): mapq >= q, min <= L <= max (FTK_NONE = unbounded).
__device__ __forceinline__ bool frag_len_ok(int len, int min_len, int max_len) {
    return (min_len == FTK_NONE || len >= min_len) && (max_len == FTK_NONE || len <= max_len);
}

}  // namespace ftk
