// stem_mma.cu -- the 7x7 stride-2 stem convolution (Cin = 3, Cout = 64) + folded BN + ReLU on tensor cores (sm_100a).
//
// Reference: model/extractors.py:112-114,148-150 (conv1 7x7 s2 p3 no bias -> bn1 -> relu), model/bisenet.py:72-74,85-87,
// SpatialPath.conv1 (:329).  Used by the 16-bit plans (fp16 / bf16 activations): the frame is rounded to the plan's
// 16-bit type on its way into shared memory, exactly like every later activation of such a plan; accumulation is fp32.
//
// Implicit GEMM without an im2col buffer: K = (ky, kx, ci) with ci padded 3 -> 4 and the 49 taps padded to 52, so one
// m16n8k16 k-step covers four taps.  The input tile lives in shared memory as [row][col][4 x 16-bit]; the A fragment of a
// 16-pixel row segment is built directly in registers -- every 32-bit A register is ONE aligned LDS.32 (channel pair
// (0,1) or (2,pad) of one tap of one pixel).  Weights [64][52*4] sit in shared memory once per (persistent) CTA and are
// read with ldmatrix; a warp owns two 16-pixel m-tiles x all 64 output channels, so a B fragment feeds two MMAs.
// Epilogue: scale/shift + ReLU in registers, the warp's 16 x 64 tile goes through a padded shared-memory buffer so the
// NHWC stores are 16 bytes per lane.
//
// Why mma.sync and not tcgen05: the whole layer is 0.12 GMAC per frame -- 0.15 % of the network; the work is building
// the A operand from a 3-channel image (a gather with 4-byte granularity that TMA cannot express), not the MMAs.
#include "common.cuh"
#include <cuda_fp16.h>
#include <cuda_bf16.h>

namespace arseg {

constexpr int SM_TH = 8, SM_TW = 32;                         // output tile (pixels)
constexpr int SM_IH = 2 * SM_TH + 5, SM_IW = 2 * SM_TW + 5;  // input tile 21 x 69
constexpr int SM_KPOS = 52, SM_K = SM_KPOS * 4;              // 49 taps padded to 52; 4 channels per tap -> K = 208 = 13 k-steps
constexpr int SM_WLD = 216;                                  // weight row stride (16-bit elements): 432 B, ldmatrix conflict-free
constexpr int SM_OLD = 72;                                   // epilogue staging row stride (16-bit elements): 144 B
constexpr int SM_THREADS = 256;
constexpr size_t SM_IN_BYTES = (size_t)SM_IH * SM_IW * 8;                 // 11592
constexpr size_t SM_W_BYTES = (size_t)64 * SM_WLD * 2;                    // 27648
constexpr size_t SM_O_BYTES = (size_t)(SM_THREADS / 32) * 16 * SM_OLD * 2;  // 18432
constexpr size_t SM_TAB_BYTES = SM_KPOS * 4;
constexpr size_t SM_SMEM = ((SM_IN_BYTES + 15) / 16 * 16) + SM_W_BYTES + SM_O_BYTES + SM_TAB_BYTES + 2 * 64 * 4;

template <typename T> struct StemMma;
template <> struct StemMma<__half> {
    static __device__ __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
    static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
        uint32_t r;
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
        return r;
    }
};
template <> struct StemMma<__nv_bfloat16> {
    static __device__ __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
    static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
        uint32_t r;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
        return r;
    }
};

__device__ __forceinline__ uint32_t stem_s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t stem_lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void stem_ldsm4(uint32_t (&r)[4], uint32_t a) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}

template <typename T>
__global__ void __launch_bounds__(SM_THREADS, 2) stem_mma_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                                 const float* __restrict__ scale, const float* __restrict__ shift,
                                                                 T* __restrict__ out, int N, int H, int W, int Ho, int Wo, int tiles_x, int tiles_y) {
    extern __shared__ __align__(16) uint8_t ssm[];
    uint8_t* s_in = ssm;                                                   // [SM_IH][SM_IW][4] 16-bit
    T* s_w = reinterpret_cast<T*>(ssm + (SM_IN_BYTES + 15) / 16 * 16);      // [64][SM_WLD]
    T* s_o = s_w + 64 * SM_WLD;                                            // [8 warps][16][SM_OLD]
    int* s_tab = reinterpret_cast<int*>(s_o + (SM_THREADS / 32) * 16 * SM_OLD);   // byte offset of tap pz inside the input tile
    float* s_sc = reinterpret_cast<float*>(s_tab + SM_KPOS);
    float* s_sh = s_sc + 64;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;

    // weights: [64][7][7][3] fp32 -> s_w[oc][(ky*7+kx)*4 + ci], zero for ci = 3 and the padding taps (once per CTA)
    for (int i = tid; i < 64 * SM_K; i += SM_THREADS) {
        const int oc = i / SM_K, k = i - oc * SM_K, pz = k >> 2, ci = k & 3;
        s_w[oc * SM_WLD + k] = from_f32<T>((pz < 49 && ci < 3) ? __ldg(w + (size_t)oc * 147 + pz * 3 + ci) : 0.f);
    }
    if (tid < SM_KPOS) { const int pz = tid < 49 ? tid : 48; s_tab[tid] = ((pz / 7) * SM_IW + pz % 7) * 8; }
    if (tid < 64) { s_sc[tid] = __ldg(scale + tid); s_sh[tid] = __ldg(shift + tid); }

    const uint32_t in_a = stem_s32(s_in), w_a = stem_s32(s_w), tab_a = stem_s32(s_tab);
    // ldmatrix lane address inside the weight tile: matrix m = lane >> 3: n-tile (m >> 1), k half (m & 1); row = lane & 7
    const uint32_t b_lane = w_a + (uint32_t)((((lane >> 4) & 1) * 8 + (lane & 7)) * SM_WLD * 2 + ((lane >> 3) & 1) * 16);
    // the warp's two m-tiles: output row `warp` of the tile, columns [0,16) and [16,32); fragment rows g and g+8 = pixels
    const uint32_t a_pix = in_a + (uint32_t)(((2 * warp) * SM_IW + 2 * g) * 8 + (t & 1) * 4);
    T* const so = s_o + warp * 16 * SM_OLD;

    const int tiles_per_img = tiles_x * tiles_y, total = N * tiles_per_img;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int n = tile / tiles_per_img, r = tile - n * tiles_per_img, ty = r / tiles_x, tx = r - ty * tiles_x;
        const int oy0 = ty * SM_TH, ox0 = tx * SM_TW, iy0 = 2 * oy0 - 3, ix0 = 2 * ox0 - 3;
        __syncthreads();                                   // previous tile's A reads are done (first pass: weights / tables written)
        const float* src = in + (size_t)n * 3 * H * W;
        for (int i = tid; i < SM_IH * SM_IW; i += SM_THREADS) {
            const int rr = i / SM_IW, cc = i - rr * SM_IW, y = iy0 + rr, x = ix0 + cc;
            float v0 = 0.f, v1 = 0.f, v2 = 0.f;
            if (y >= 0 && y < H && x >= 0 && x < W) {
                const size_t o = (size_t)y * W + x;
                v0 = __ldg(src + o); v1 = __ldg(src + (size_t)H * W + o); v2 = __ldg(src + 2 * (size_t)H * W + o);
            }
            *reinterpret_cast<uint2*>(s_in + (size_t)i * 8) = make_uint2(StemMma<T>::pack(v0, v1), StemMma<T>::pack(v2, 0.f));
        }
        __syncthreads();

        float acc[2][8][4];
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) acc[m][nt][0] = acc[m][nt][1] = acc[m][nt][2] = acc[m][nt][3] = 0.f;
#pragma unroll 1
        for (int s = 0; s < SM_KPOS / 4; ++s) {
            // taps of this k-step seen by this thread: k = 16s + 2t (+1) -> tap 4s + (t >> 1); k + 8 -> two taps further
            const uint32_t o0 = stem_lds32(tab_a + (uint32_t)((4 * s + (t >> 1)) * 4)), o1 = stem_lds32(tab_a + (uint32_t)((4 * s + 2 + (t >> 1)) * 4));
            uint32_t a[2][4];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const uint32_t p0 = a_pix + (uint32_t)(m * 32 * 8);      // m-tile 1 starts 16 output columns = 32 input columns further
                a[m][0] = stem_lds32(p0 + o0);
                a[m][1] = stem_lds32(p0 + o0 + 16 * 8);                  // fragment row g + 8 = 8 output pixels = 16 input columns
                a[m][2] = stem_lds32(p0 + o1);
                a[m][3] = stem_lds32(p0 + o1 + 16 * 8);
            }
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                uint32_t b[4];
                stem_ldsm4(b, b_lane + (uint32_t)(np * 16 * SM_WLD * 2 + s * 32));
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    StemMma<T>::mma(acc[m][2 * np], a[m], b[0], b[1]);
                    StemMma<T>::mma(acc[m][2 * np + 1], a[m], b[2], b[3]);
                }
            }
        }
        // epilogue: folded BN + ReLU, 16 x 64 tile through shared memory, 16-byte NHWC stores
        const int oy = oy0 + warp;
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            __syncwarp();
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const int c = 8 * nt + 2 * t;
                const float s0 = s_sc[c], s1 = s_sc[c + 1], h0 = s_sh[c], h1 = s_sh[c + 1];
                *reinterpret_cast<uint32_t*>(so + g * SM_OLD + c) =
                    StemMma<T>::pack(fmaxf(fmaf(acc[m][nt][0], s0, h0), 0.f), fmaxf(fmaf(acc[m][nt][1], s1, h1), 0.f));
                *reinterpret_cast<uint32_t*>(so + (g + 8) * SM_OLD + c) =
                    StemMma<T>::pack(fmaxf(fmaf(acc[m][nt][2], s0, h0), 0.f), fmaxf(fmaf(acc[m][nt][3], s1, h1), 0.f));
            }
            __syncwarp();
            if (oy < Ho) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int px = (lane >> 3) + 4 * i, ch = lane & 7, ox = ox0 + 16 * m + px;
                    if (ox < Wo)
                        *reinterpret_cast<uint4*>(out + (((size_t)n * Ho + oy) * Wo + ox) * 64 + ch * 8) =
                            *reinterpret_cast<const uint4*>(so + px * SM_OLD + ch * 8);
                }
            }
        }
    }
}

template <typename T>
static int stem_mma_launch_t(const float* in, const float* w, const float* scale, const float* shift, T* out, int N, int H, int W,
                             int Ho, int Wo, cudaStream_t st) {
    static bool configured[64] = {false};
    int dev = 0;
    ARSEG_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        ARSEG_CUDA(cudaFuncSetAttribute(stem_mma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_SMEM));
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    const int tiles_x = ceil_div(Wo, SM_TW), tiles_y = ceil_div(Ho, SM_TH);
    const long long total = (long long)N * tiles_x * tiles_y;
    ARSEG_REQUIRE(total > 0 && total < 2147483647LL, "stem: too many tiles");
    const int sms = sm_count() > 0 ? sm_count() : 148;
    const int grid = (int)(total < 2LL * sms ? total : 2LL * sms);       // persistent: two CTAs per SM walk the tile list
    stem_mma_kernel<T><<<grid, SM_THREADS, SM_SMEM, st>>>(in, w, scale, shift, out, N, H, W, Ho, Wo, tiles_x, tiles_y);
    ARSEG_CHECK_LAUNCH("stem_mma");
    return ARSEG_OK;
}

int stem_mma_launch(const float* in, const float* w, const float* scale, const float* shift, void* out, int out_dtype, int N, int H,
                    int W, int Ho, int Wo, cudaStream_t st) {
    if (out_dtype == ARSEG_F16) return stem_mma_launch_t<__half>(in, w, scale, shift, (__half*)out, N, H, W, Ho, Wo, st);
    if (out_dtype == ARSEG_BF16) return stem_mma_launch_t<__nv_bfloat16>(in, w, scale, shift, (__nv_bfloat16*)out, N, H, W, Ho, Wo, st);
    ARSEG_UNSUPPORTED("stem_mma: dtype %d", out_dtype);
}

}  // namespace arseg
