// Shared helpers for the s2d_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/s2d_b200.h"

namespace s2d {

void set_error(const char* fmt, ...);

#define S2D_CHECK_ARG(cond, ...)            \
    do {                                    \
        if (!(cond)) {                      \
            s2d::set_error(__VA_ARGS__);    \
            return -1;                      \
        }                                   \
    } while (0)

#define S2D_CHECK_LAUNCH(name)                                                       \
    do {                                                                             \
        cudaError_t e_ = cudaGetLastError();                                         \
        if (e_ != cudaSuccess) {                                                     \
            s2d::set_error("%s: launch failed: %s", name, cudaGetErrorString(e_));   \
            return -2;                                                               \
        }                                                                            \
    } while (0)

// streaming 128-bit load that does not pollute L1 (read-once data: tracks, flags, labels)
__device__ __forceinline__ int4 ld_stream(const int4* p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ int warp_sum(int v) { return __reduce_add_sync(0xffffffffu, v); }

__device__ __forceinline__ int cdiv(int a, int b) { return (a + b - 1) / b; }

}  // namespace s2d
