// Shared helpers of the chromosight_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include "../../include/chromosight_b200.h"

namespace cs {

void set_error(const char *fmt, ...);
extern std::atomic<long long> g_launches;

#define CS_CUDA(call)                                                              \
    do {                                                                           \
        cudaError_t _e = (call);                                                   \
        if (_e != cudaSuccess) {                                                   \
            cs::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,             \
                          cudaGetErrorString(_e));                                 \
            return CS_ERR_CUDA;                                                    \
        }                                                                          \
    } while (0)

#define CS_REQUIRE(cond, ...)                                                      \
    do {                                                                           \
        if (!(cond)) {                                                             \
            cs::set_error(__VA_ARGS__);                                            \
            return CS_ERR_INVALID;                                                 \
        }                                                                          \
    } while (0)

#define CS_LAUNCHED() (cs::g_launches.fetch_add(1, std::memory_order_relaxed))

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Pixel (Y, X) -> element offset in an image with layout L.
__host__ __device__ __forceinline__ long long img_index(int pitch, int dlo, int Y, int X) {
    return (long long)Y * pitch + (X - dlo);
}

__device__ __forceinline__ float quiet_nan_f() { return __int_as_float(0x7fc00000); }

}  // namespace cs
