// pyramid.cu -- the pyramid-pooling module as three launches instead of 4 x (pool, 1x1 conv, up-sample) + concat.
//
// Reference: PSPModule (model/pspnet.py:14-31: AdaptiveAvgPool2d(s) -> 1x1 conv (no bias) -> bilinear up-sample,
// align_corners=False -> cat(stages..., feats)) and PPM (model/pspnet_semseg.py:12-30: AdaptiveAvgPool2d(bin) ->
// 1x1 conv (no bias) -> BN -> ReLU -> bilinear, align_corners=True -> cat(x, branches...)).
// The pooled maps are tiny ([N, sum s^2, C]: 50 positions for bins 1,2,3,6) and stay fp32; the 1x1 convolutions
// run in plain fp32 FMA with fp32 weights in every precision mode.
#include "common.cuh"

namespace arseg {

constexpr int PYR_MAX_LEV = 8;
struct PyrLevels {
    int nlev;
    int bins[PYR_MAX_LEV];
    int off[PYR_MAX_LEV + 1];      // prefix sums of bins^2
};

template <typename T, int VEC> struct alignas(sizeof(T) * VEC) PVec { T v[VEC]; };

// ---- 1. all pyramid levels of adaptive average pooling in one launch ---------------------------------------
// grid = (N * total_bins, C / 64); block = 256 = (64 / VEC channel lanes) x pixel lanes
template <typename T, int VEC>
__global__ void __launch_bounds__(256) pyramid_pool_kernel(const T* __restrict__ in, float* __restrict__ out, int N, int H, int W, int C,
                                                           PyrLevels lv) {
    constexpr int CL = 64 / VEC, PL = 256 / CL;
    __shared__ float red[PL][65];
    const int total = lv.off[lv.nlev];
    const int n = blockIdx.x / total, b = blockIdx.x - n * total;
    int l = 0;
    while (l + 1 < lv.nlev && b >= lv.off[l + 1]) ++l;
    const int s = lv.bins[l], bi = b - lv.off[l], oy = bi / s, ox = bi - oy * s;
    // nn.AdaptiveAvgPool2d window: [floor(i*H/s), ceil((i+1)*H/s))
    const int ys = (oy * H) / s, ye = ((oy + 1) * H + s - 1) / s, xs = (ox * W) / s, xe = ((ox + 1) * W + s - 1) / s;
    const int ww = xe - xs, cnt = (ye - ys) * ww;
    const int cl = threadIdx.x % CL, pl = threadIdx.x / CL;
    const int c0 = blockIdx.y * 64 + cl * VEC;
    float acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
    if (c0 < C) {
        const T* base = in + (size_t)n * H * W * C + c0;
        for (int i = pl; i < cnt; i += PL) {
            const int y = ys + i / ww, x = xs + i % ww;
            const PVec<T, VEC> v = *reinterpret_cast<const PVec<T, VEC>*>(base + ((size_t)y * W + x) * C);
#pragma unroll
            for (int k = 0; k < VEC; ++k) acc[k] += to_f32(v.v[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) red[pl][cl * VEC + k] = acc[k];
    __syncthreads();
    if (threadIdx.x < 64) {
        float sum = 0.f;
#pragma unroll 8
        for (int i = 0; i < PL; ++i) sum += red[i][threadIdx.x];
        const int c = blockIdx.y * 64 + threadIdx.x;
        if (c < C) out[((size_t)n * total + b) * C + c] = sum / (float)cnt;
    }
}

// ---- 2. the per-level 1x1 convolutions (+ folded BN, + ReLU) as one grouped fp32 GEMM -----------------------
// rows of level l: N * s_l^2 pooled positions; CTA = 16 rows x 64 output channels; grid.x walks the 16-row blocks
// of all levels, grid.y the 64-channel slabs.
struct PyrConvParams {
    const float* pooled; const float* w; const float* scale; const float* shift; float* out;
    int N, C, Cout, relu;
    PyrLevels lv;
    int blk_off[PYR_MAX_LEV + 1];   // prefix sums of ceil(N * s^2 / 16)
};
__global__ void __launch_bounds__(256) pyramid_conv_kernel(PyrConvParams p) {
    extern __shared__ __align__(16) float s_x[];          // [16][C]
    int l = 0;
    while (l + 1 < p.lv.nlev && (int)blockIdx.x >= p.blk_off[l + 1]) ++l;
    const int s2 = p.lv.bins[l] * p.lv.bins[l], rows = p.N * s2;
    const int r0 = ((int)blockIdx.x - p.blk_off[l]) * 16;
    const int total = p.lv.off[p.lv.nlev];
    for (int i = threadIdx.x; i < 16 * (p.C / 4); i += 256) {
        const int r = i / (p.C / 4), k4 = i - r * (p.C / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r0 + r < rows) {
            const int n = (r0 + r) / s2, b = (r0 + r) - n * s2;
            v = *reinterpret_cast<const float4*>(p.pooled + ((size_t)n * total + p.lv.off[l] + b) * p.C + 4 * k4);
        }
        *reinterpret_cast<float4*>(s_x + r * p.C + 4 * k4) = v;
    }
    __syncthreads();
    const int col = blockIdx.y * 64 + (threadIdx.x & 63), rg = threadIdx.x >> 6;
    if (col >= p.Cout) return;
    const float* wrow = p.w + ((size_t)l * p.Cout + col) * p.C;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < p.C; k += 4) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(wrow + k));
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float4 xv = *reinterpret_cast<const float4*>(s_x + (rg * 4 + r) * p.C + k);
            acc[r] = fmaf(xv.x, wv.x, acc[r]); acc[r] = fmaf(xv.y, wv.y, acc[r]);
            acc[r] = fmaf(xv.z, wv.z, acc[r]); acc[r] = fmaf(xv.w, wv.w, acc[r]);
        }
    }
    const float sc = p.scale ? __ldg(p.scale + l * p.Cout + col) : 1.f, sh = p.shift ? __ldg(p.shift + l * p.Cout + col) : 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int row = r0 + rg * 4 + r;
        if (row < rows) {
            const int n = row / s2, b = row - n * s2;
            float v = fmaf(acc[r], sc, sh);
            if (p.relu) v = fmaxf(v, 0.f);
            p.out[((size_t)n * total + p.lv.off[l] + b) * p.Cout + col] = v;
        }
    }
}

// ---- 3. bilinear up-sampling of every level + concat with the feature map ---------------------------------
// out[n,y,x, stage_coff + l*Cout + c] = bilinear(stage_l)[y,x,c];  out[n,y,x, feats_coff + c] = feats[n,y,x,c]
template <typename T>
__global__ void __launch_bounds__(256) pyramid_upcat_kernel(const float* __restrict__ stage, const T* __restrict__ feats, T* __restrict__ out,
                                                            int N, int H, int W, int Cout, int Cf, int stage_coff, int feats_coff, int mode,
                                                            PyrLevels lv) {
    constexpr int VEC = 16 / (int)sizeof(T);               // channels per 16-byte store
    const int Ct = lv.nlev * Cout + Cf, total = lv.off[lv.nlev];
    const int sv = lv.nlev * Cout / VEC, fv = Cf / VEC;
    // horizontal taps / weights per (level, x): computed once per CTA instead of once per 16-byte store
    extern __shared__ __align__(16) float4 s_xt[];          // [nlev][W] {x0, x1 (as int bits), lx0, lx1}
    for (int i = threadIdx.x; i < lv.nlev * W; i += blockDim.x) {
        const int l = i / W, x = i - l * W, s = lv.bins[l];
        int x0, x1;
        float lx0, lx1;
        bilinear_src(resize_scale(s, W, mode), x, s, mode, x0, x1, lx0, lx1);
        s_xt[i] = make_float4(__int_as_float(x0), __int_as_float(x1), lx0, lx1);
    }
    __syncthreads();
    for (int row = blockIdx.x; row < N * H; row += gridDim.x) {
        const int n = row / H, y = row - n * H;
        T* orow = out + (size_t)row * W * Ct;
        const T* frow = feats + (size_t)row * W * Cf;
        // a thread owns one 16-byte channel pack and walks the row: level, channel and the vertical taps are loop invariant
        for (int e = threadIdx.x; e < sv + fv; e += blockDim.x) {
            if (e >= sv) {
                const int c = (e - sv) * VEC;
                for (int x = 0; x < W; ++x)
                    *reinterpret_cast<PVec<T, VEC>*>(orow + (size_t)x * Ct + feats_coff + c) =
                        *reinterpret_cast<const PVec<T, VEC>*>(frow + (size_t)x * Cf + c);
                continue;
            }
            const int ch = e * VEC, l = ch / Cout, c = ch - l * Cout, s = lv.bins[l];
            int y0, y1;
            float ly0, ly1;
            bilinear_src(resize_scale(s, H, mode), y, s, mode, y0, y1, ly0, ly1);
            const float* sb = stage + ((size_t)n * total + lv.off[l]) * Cout + c;
            const float* r0 = sb + (size_t)(y0 * s) * Cout;
            const float* r1 = sb + (size_t)(y1 * s) * Cout;
            const float4* xt = s_xt + l * W;
            for (int x = 0; x < W; ++x) {
                const float4 t = xt[x];
                const int x0 = __float_as_int(t.x), x1 = __float_as_int(t.y);
                const float lx0 = t.z, lx1 = t.w;
                const float4* a = reinterpret_cast<const float4*>(r0 + (size_t)x0 * Cout);
                const float4* b = reinterpret_cast<const float4*>(r0 + (size_t)x1 * Cout);
                const float4* cc = reinterpret_cast<const float4*>(r1 + (size_t)x0 * Cout);
                const float4* d = reinterpret_cast<const float4*>(r1 + (size_t)x1 * Cout);
                PVec<T, VEC> o;
#pragma unroll
                for (int q = 0; q < VEC / 4; ++q) {
                    const float4 va = __ldg(a + q), vb = __ldg(b + q), vc = __ldg(cc + q), vd = __ldg(d + q);
                    // same association as ATen: ly0 * (lx0 * a + lx1 * b) + ly1 * (lx0 * c + lx1 * d)
                    o.v[4 * q + 0] = from_f32<T>(ly0 * (lx0 * va.x + lx1 * vb.x) + ly1 * (lx0 * vc.x + lx1 * vd.x));
                    o.v[4 * q + 1] = from_f32<T>(ly0 * (lx0 * va.y + lx1 * vb.y) + ly1 * (lx0 * vc.y + lx1 * vd.y));
                    o.v[4 * q + 2] = from_f32<T>(ly0 * (lx0 * va.z + lx1 * vb.z) + ly1 * (lx0 * vc.z + lx1 * vd.z));
                    o.v[4 * q + 3] = from_f32<T>(ly0 * (lx0 * va.w + lx1 * vb.w) + ly1 * (lx0 * vc.w + lx1 * vd.w));
                }
                *reinterpret_cast<PVec<T, VEC>*>(orow + (size_t)x * Ct + stage_coff + ch) = o;
            }
        }
    }
}

static int make_levels(const int* bins, int nlev, PyrLevels& lv) {
    ARSEG_REQUIRE(bins && nlev > 0 && nlev <= PYR_MAX_LEV, "pyramid: 1..%d levels", PYR_MAX_LEV);
    lv.nlev = nlev;
    lv.off[0] = 0;
    for (int i = 0; i < nlev; ++i) {
        ARSEG_REQUIRE(bins[i] > 0 && bins[i] <= 64, "pyramid: bad bin count %d", bins[i]);
        lv.bins[i] = bins[i];
        lv.off[i + 1] = lv.off[i] + bins[i] * bins[i];
    }
    return ARSEG_OK;
}

}  // namespace arseg

using namespace arseg;

extern "C" {

int arseg_pyramid_pool_nhwc(const void* in, float* out, int dtype, int N, int H, int W, int C, const int* bins, int nlev,
                            arseg_stream_t stream) {
    ARSEG_REQUIRE(in && out && N > 0 && H > 0 && W > 0 && C > 0, "pyramid_pool: bad args");
    PyrLevels lv;
    if (int rc = make_levels(bins, nlev, lv)) return rc;
    dim3 grid(N * lv.off[nlev], ceil_div(C, 64));
    ARSEG_REQUIRE(grid.y <= 65535, "pyramid_pool: C too large");
    cudaStream_t st = as_stream(stream);
    if (dtype == ARSEG_F32) {
        ARSEG_REQUIRE(C % 4 == 0 && (uintptr_t)in % 16 == 0, "pyramid_pool: C %% 4 and 16-byte alignment");
        pyramid_pool_kernel<float, 4><<<grid, 256, 0, st>>>((const float*)in, out, N, H, W, C, lv);
    } else if (dtype == ARSEG_F16 || dtype == ARSEG_BF16) {
        ARSEG_REQUIRE(C % 8 == 0 && (uintptr_t)in % 16 == 0, "pyramid_pool: C %% 8 and 16-byte alignment");
        if (dtype == ARSEG_F16) pyramid_pool_kernel<__half, 8><<<grid, 256, 0, st>>>((const __half*)in, out, N, H, W, C, lv);
        else pyramid_pool_kernel<__nv_bfloat16, 8><<<grid, 256, 0, st>>>((const __nv_bfloat16*)in, out, N, H, W, C, lv);
    } else ARSEG_UNSUPPORTED("pyramid_pool: dtype %d", dtype);
    ARSEG_CHECK_LAUNCH("pyramid_pool");
    return ARSEG_OK;
}

int arseg_pyramid_conv1x1(const float* pooled, const float* w, const float* scale, const float* shift, int relu, float* out,
                          int N, int C, int Cout, const int* bins, int nlev, arseg_stream_t stream) {
    ARSEG_REQUIRE(pooled && w && out && N > 0 && C > 0 && Cout > 0, "pyramid_conv: bad args");
    ARSEG_REQUIRE(C % 4 == 0 && C <= 896, "pyramid_conv: C=%d must be a multiple of 4, <= 896", C);
    PyrConvParams p;
    if (int rc = make_levels(bins, nlev, p.lv)) return rc;
    p.pooled = pooled; p.w = w; p.scale = scale; p.shift = shift; p.out = out; p.N = N; p.C = C; p.Cout = Cout; p.relu = relu;
    p.blk_off[0] = 0;
    for (int i = 0; i < nlev; ++i) p.blk_off[i + 1] = p.blk_off[i] + ceil_div(N * bins[i] * bins[i], 16);
    const size_t smem = (size_t)16 * C * 4;
    static bool configured[64] = {false};
    int dev = 0;
    ARSEG_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        ARSEG_CUDA(cudaFuncSetAttribute(pyramid_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 896 * 4));
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    dim3 grid(p.blk_off[nlev], ceil_div(Cout, 64));
    pyramid_conv_kernel<<<grid, 256, smem, as_stream(stream)>>>(p);
    ARSEG_CHECK_LAUNCH("pyramid_conv");
    return ARSEG_OK;
}

int arseg_pyramid_upsample_concat(const float* stage, const void* feats, void* out, int dtype, int N, int H, int W, int Cout,
                                  int Cf, int stage_coff, int feats_coff, int mode, const int* bins, int nlev,
                                  arseg_stream_t stream) {
    ARSEG_REQUIRE(stage && feats && out && N > 0 && H > 0 && W > 0 && Cout > 0 && Cf > 0, "pyramid_upcat: bad args");
    ARSEG_REQUIRE(mode == ARSEG_RESIZE_BILINEAR || mode == ARSEG_RESIZE_BILINEAR_AC, "pyramid_upcat: bilinear modes only");
    PyrLevels lv;
    if (int rc = make_levels(bins, nlev, lv)) return rc;
    const int Ct = nlev * Cout + Cf;
    const int vec = dtype == ARSEG_F32 ? 4 : 8;
    ARSEG_REQUIRE((uintptr_t)stage % 16 == 0 && Cout % 4 == 0, "pyramid_upcat: stage alignment");
    ARSEG_REQUIRE(Cout % vec == 0 && Cf % vec == 0 && stage_coff % vec == 0 && feats_coff % vec == 0, "pyramid_upcat: channel counts %% %d", vec);
    ARSEG_REQUIRE((stage_coff == 0 && feats_coff == nlev * Cout) || (feats_coff == 0 && stage_coff == Cf), "pyramid_upcat: slices must tile the output");
    ARSEG_REQUIRE((uintptr_t)feats % 16 == 0 && (uintptr_t)out % 16 == 0, "pyramid_upcat: 16-byte alignment");
    (void)Ct;
    const long long rows = (long long)N * H, cap = (long long)sm_count() * 16;
    const int grid = (int)(rows < cap ? rows : cap);
    cudaStream_t st = as_stream(stream);
    const size_t xt_bytes = (size_t)lv.nlev * W * sizeof(float4);
    ARSEG_REQUIRE(xt_bytes <= 48 * 1024, "pyramid_upcat: W too large");
    if (dtype == ARSEG_F32)
        pyramid_upcat_kernel<float><<<grid, 256, xt_bytes, st>>>(stage, (const float*)feats, (float*)out, N, H, W, Cout, Cf, stage_coff, feats_coff, mode, lv);
    else if (dtype == ARSEG_F16)
        pyramid_upcat_kernel<__half><<<grid, 256, xt_bytes, st>>>(stage, (const __half*)feats, (__half*)out, N, H, W, Cout, Cf, stage_coff, feats_coff, mode, lv);
    else if (dtype == ARSEG_BF16)
        pyramid_upcat_kernel<__nv_bfloat16><<<grid, 256, xt_bytes, st>>>(stage, (const __nv_bfloat16*)feats, (__nv_bfloat16*)out, N, H, W, Cout, Cf, stage_coff, feats_coff, mode, lv);
    else ARSEG_UNSUPPORTED("pyramid_upcat: dtype %d", dtype);
    ARSEG_CHECK_LAUNCH("pyramid_upcat");
    return ARSEG_OK;
}

}  // extern "C"
