// Shared helpers for libscrib200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/scrib200.h"

namespace scrib200 {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define SCRIB200_REQUIRE(cond, ...)          \
    do {                                     \
        if (!(cond)) {                       \
            scrib200::set_error(__VA_ARGS__); \
            return SCRIB200_EINVAL;          \
        }                                    \
    } while (0)

#define SCRIB200_CHECK_LAUNCH(name)                                                     \
    do {                                                                                \
        cudaError_t e__ = cudaGetLastError();                                           \
        if (e__ != cudaSuccess) {                                                       \
            scrib200::set_error("%s: CUDA launch failed: %s", name, cudaGetErrorString(e__)); \
            return SCRIB200_ECUDA;                                                      \
        }                                                                               \
        scrib200::count_launch();                                                       \
    } while (0)

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cscale(double s, double2 a) { return make_double2(s * a.x, s * a.y); }
__device__ __forceinline__ void cfma(double2& acc, double2 a, double2 b) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace scrib200
