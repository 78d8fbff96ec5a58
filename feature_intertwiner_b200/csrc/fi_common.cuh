// Shared helpers for libfi_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/fi_b200.h"

#define FI_API extern "C" __attribute__((visibility("default")))

namespace fi {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

void set_error(int status, const char *fmt, ...);
int ok();
void count_launch();   // bumps the process-wide kernel-launch counter behind fi_kernel_launches()
int option(int key);   // fi_get_option without the range check (FI_OPT_*)

// Checks the launch that was just enqueued (no sync).
inline int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error(FI_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
        return FI_ERR_CUDA;
    }
    count_launch();
    return ok();
}

#define FI_REQUIRE(cond, ...)                    \
    do {                                         \
        if (!(cond)) {                           \
            fi::set_error(FI_ERR_INVALID, __VA_ARGS__); \
            return FI_ERR_INVALID;               \
        }                                        \
    } while (0)

// ------------------------------------------------------------------------------------------------
// crop_and_resize sampling geometry, one axis.
//
// Every operation is an explicitly rounded fp32 intrinsic so that nvcc cannot contract mul+add into an
// FMA: the reference's CPU build (-std=c99, x86-64) evaluates these un-fused, and the tap INDICES
// floor/ceil(pos) must come out bit-identical (SURVEY.md section 7 "Index parity").
// Semantics: lib/roi_align/src/crop_and_resize.c:44-74 == cuda/crop_and_resize_kernel.cu:40-70.
// ------------------------------------------------------------------------------------------------
struct AxisTap {
    int lo, hi;   // floor / ceil pixel
    float frac;   // weight of `hi`
    bool inside;  // false -> extrapolation value
};

__device__ __forceinline__ float axis_step(float c1, float c2, int extent, int crop) {
    if (crop > 1) return __fdiv_rn(__fmul_rn(__fsub_rn(c2, c1), (float)(extent - 1)), (float)(crop - 1));
    return 0.f;
}

__device__ __forceinline__ AxisTap axis_sample(float c1, float c2, float step, int k, int extent, int crop) {
    float pos;
    if (crop > 1) {
        pos = __fadd_rn(__fmul_rn(c1, (float)(extent - 1)), __fmul_rn((float)k, step));
    } else {
        // `0.5 * (c1 + c2) * (extent - 1)` is double arithmetic in C (crop_and_resize.c:56)
        pos = (float)(0.5 * (double)__fadd_rn(c1, c2) * (double)(extent - 1));
    }
    AxisTap t;
    t.inside = !(pos < 0.f || pos > (float)(extent - 1));
    t.lo = (int)floorf(pos);
    t.hi = (int)ceilf(pos);
    t.frac = __fsub_rn(pos, (float)t.lo);
    return t;
}

// top + (bottom - top) * w without contraction (crop_and_resize.c:102-106)
__device__ __forceinline__ float lerp_rn(float a, float b, float w) { return __fadd_rn(a, __fmul_rn(__fsub_rn(b, a), w)); }

__device__ __forceinline__ float4 lerp_rn(float4 a, float4 b, float w) {
    return make_float4(lerp_rn(a.x, b.x, w), lerp_rn(a.y, b.y, w), lerp_rn(a.z, b.z, w), lerp_rn(a.w, b.w, w));
}

// Read-only 128-bit load that does not allocate in L1 is NOT what we want for the feature map: taps are
// re-read by neighbouring samples, so the default (L1-allocating, read-only) path is used.
__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

// Streaming 128-bit store: crops / grads are written once and never re-read by this kernel; evict-first
// keeps them from displacing the feature map in L2.
__device__ __forceinline__ void st_stream4(float *p, float4 v) { __stcs(reinterpret_cast<float4 *>(p), v); }

__host__ __device__ constexpr int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace fi
