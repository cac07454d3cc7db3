// common.cuh -- shared helpers for libarseg_sm100a.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/arseg.h"

namespace arseg {

void set_error(const char* fmt, ...);

#define ARSEG_REQUIRE(cond, ...)                                  \
    do {                                                          \
        if (!(cond)) {                                            \
            ::arseg::set_error(__VA_ARGS__);                      \
            return ARSEG_E_BADARG;                                \
        }                                                         \
    } while (0)

#define ARSEG_UNSUPPORTED(...)                                    \
    do {                                                          \
        ::arseg::set_error(__VA_ARGS__);                          \
        return ARSEG_E_UNSUPPORTED;                               \
    } while (0)

// Launch check: catches configuration errors without synchronising.
#define ARSEG_CHECK_LAUNCH(name)                                                         \
    do {                                                                                 \
        cudaError_t e__ = cudaGetLastError();                                            \
        if (e__ != cudaSuccess) {                                                        \
            ::arseg::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));  \
            return ARSEG_E_CUDA;                                                         \
        }                                                                                \
    } while (0)

#define ARSEG_CUDA(call)                                                                 \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            ::arseg::set_error("%s failed: %s", #call, cudaGetErrorString(e__));         \
            return ARSEG_E_CUDA;                                                         \
        }                                                                                \
    } while (0)

static inline cudaStream_t as_stream(arseg_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

int sm_count();

// ---- element type helpers -------------------------------------------------------------------
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
// fp16 storage (11-bit significand, the same as TF32): round to nearest, saturate to +-65504 instead of overflowing to inf
__device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) {
    unsigned short r;
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(v));
    return __ushort_as_half(r);
}

// ---- ATen-compatible resampling index math (aten/src/ATen/native/UpSample.h semantics) --------
// mode: ARSEG_RESIZE_BILINEAR (align_corners=False), ARSEG_RESIZE_BILINEAR_AC (True)
__host__ __device__ __forceinline__ float resize_scale(int in, int out, int mode) {
    if (mode == ARSEG_RESIZE_BILINEAR_AC) return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
    return (float)in / (float)out;
}
__device__ __forceinline__ void bilinear_src(float scale, int dst, int in, int mode, int& i0, int& i1, float& l0, float& l1) {
    float r;
    if (mode == ARSEG_RESIZE_BILINEAR_AC) {
        r = scale * (float)dst;
    } else {
        r = scale * ((float)dst + 0.5f) - 0.5f;
        r = r < 0.f ? 0.f : r;
    }
    i0 = (int)r;
    if (i0 > in - 1) i0 = in - 1;
    i1 = i0 + ((i0 < in - 1) ? 1 : 0);
    l1 = r - (float)i0;
    l0 = 1.f - l1;
}
__device__ __forceinline__ int nearest_src(float scale, int dst, int in) {
    int s = (int)floorf((float)dst * scale);
    return s < in - 1 ? s : in - 1;
}

// Sample position (input pixels) of warpFeature for destination (x,y) with flow (u,v) [f64 pixels].
__device__ __forceinline__ void warp_source_pos(int x, int y, double u, double v, int W, int H, float& ix, float& iy) {
    // evaluation.py:76-83: grid(float32)+flow in the flow's dtype, 2*g/max(W-1,1)-1, cast to fp32
    const double vx = (double)(float)x + u, vy = (double)(float)y + v;
    const float gx = (float)(2.0 * vx / (double)max(W - 1, 1) - 1.0);
    const float gy = (float)(2.0 * vy / (double)max(H - 1, 1) - 1.0);
    // F.grid_sample defaults (evaluation.py:85): align_corners=False un-normalisation
    ix = ((gx + 1.f) * (float)W - 1.f) / 2.f;
    iy = ((gy + 1.f) * (float)H - 1.f) / 2.f;
}

}  // namespace arseg
