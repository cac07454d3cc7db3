// mma_bench.cu -- micro-benchmark: warp-level mma.sync throughput on sm_100a (tf32 m16n8k8, bf16 m16n8k16)
// and FFMA throughput for comparison.  Used to choose the arithmetic engine of the CReFF window attention
// (see DESIGN.md).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench mma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

template <int ACC>
__global__ void __launch_bounds__(256) k_tf32(float* out, int iters) {
    float c[ACC][4];
#pragma unroll
    for (int i = 0; i < ACC; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 9, b0 = threadIdx.x ^ 5, b1 = 11;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456f) out[0] = s;
}

template <int ACC>
__global__ void __launch_bounds__(256) k_bf16(float* out, int iters) {
    float c[ACC][4];
#pragma unroll
    for (int i = 0; i < ACC; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 9, b0 = threadIdx.x ^ 5, b1 = 11;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456f) out[0] = s;
}

template <int ACC>
__global__ void __launch_bounds__(256) k_ffma(float* out, int iters, float x, float y) {
    float c[ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i) c[i] = (float)i + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i) c[i] = fmaf(c[i], x, y);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ACC; ++i) s += c[i];
    if (s == 123.456f) out[0] = s;
}

template <typename F>
static float time_ms(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    float* out;
    cudaMalloc(&out, 4);
    const int iters = 4096;
    for (int warps_per_sm : {4, 8, 16, 32}) {
        const int blocks = 148 * warps_per_sm / 8;
        {
            float ms = time_ms([&] { k_tf32<8><<<blocks, 256>>>(out, iters); });
            double fl = (double)blocks * 8 * iters * 8 * 2.0 * 16 * 8 * 8;
            printf("tf32 m16n8k8  warps/SM=%2d: %.3f ms  %.1f TFLOP/s\n", warps_per_sm, ms, fl / ms * 1e-9);
        }
        {
            float ms = time_ms([&] { k_bf16<8><<<blocks, 256>>>(out, iters); });
            double fl = (double)blocks * 8 * iters * 8 * 2.0 * 16 * 8 * 16;
            printf("bf16 m16n8k16 warps/SM=%2d: %.3f ms  %.1f TFLOP/s\n", warps_per_sm, ms, fl / ms * 1e-9);
        }
        {
            float ms = time_ms([&] { k_ffma<8><<<blocks, 256>>>(out, iters * 8, 1.0001f, 0.5f); });
            double fl = (double)blocks * 256 * iters * 8 * 8 * 2.0;
            printf("ffma          warps/SM=%2d: %.3f ms  %.1f TFLOP/s\n", warps_per_sm, ms, fl / ms * 1e-9);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
