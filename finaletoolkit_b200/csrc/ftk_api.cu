// ABI bookkeeping: version, error strings, last CUDA error (thread-local).
#include <stdio.h>

#include "ftk_common.cuh"

namespace ftk {
static thread_local char g_last_cuda_error[512] = "";

int cuda_fail(cudaError_t e, const char *what) {
    snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "%s: %s (%s)", what,
             cudaGetErrorName(e), cudaGetErrorString(e));
    return FTK_E_CUDA;
}
}  // namespace ftk

extern "C" int ftk_abi_version(void) { return FTK_ABI_VERSION; }

extern "C" const char *ftk_last_cuda_error(void) { return ftk::g_last_cuda_error; }

extern "C" const char *ftk_error_string(int code) {
    switch (code) {
        case FTK_OK: return "ok";
        case FTK_E_INVALID: return "invalid argument";
        case FTK_E_CUDA: return "CUDA error (see ftk_last_cuda_error)";
        case FTK_E_RANGE: return "size or coordinate out of the supported range";
        case FTK_E_IO: return "file unreadable or not valid (b)gzip";
        default: return "unknown error";
    }
}
