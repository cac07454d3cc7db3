// conv_simt.cu -- implicit-GEMM convolution on CUDA cores with fp32 accumulation in plain FMA order.
// This is the exact-arithmetic engine (ARSEG_CONV_SIMT_F32): it reproduces the reference's fp32 conv
// numerics up to summation order, is used for the parity gate and for layers the tcgen05 engine does
// not take (stride 2, tiny channel counts).  The throughput engine is conv_tc.cu.
//
// GEMM view: M = N*Ho*Wo output pixels, Ncol = Cout, K = KH*KW*Cin; A is gathered on the fly from the
// NHWC input (zero fill outside the image), B = weights [Cout][KH][KW][Cin] (K contiguous).
#include "common.cuh"

namespace arseg {

constexpr int BM = 128, BN = 64, BK = 16, LDA = BM + 4, LDB = BN + 4;

template <typename T> struct Vec4 { T v[4]; };
template <> struct alignas(16) Vec4<float> { float v[4]; };
template <> struct alignas(8) Vec4<__nv_bfloat16> { __nv_bfloat16 v[4]; };
template <> struct alignas(8) Vec4<__half> { __half v[4]; };

struct ConvParams {
    const void* in; const void* w; const float* scale; const float* shift; const void* res; void* out;
    int N, Hi, Wi, Cin, Cout, KH, KW, stride, pad, dil, Ho, Wo, ocs, oco, act;
    float slope;
    long long M; int K;
};

template <typename T>
__global__ void __launch_bounds__(256) conv_simt_kernel(ConvParams p) {
    __shared__ __align__(16) float As[BK][LDA];
    __shared__ __align__(16) float Bs[BK][LDB];
    const T* __restrict__ in = reinterpret_cast<const T*>(p.in);
    const T* __restrict__ w = reinterpret_cast<const T*>(p.w);
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    // A-gather bookkeeping: this thread fetches rows a_row[0..1], channels kq*4..+3 of each K chunk
    const int kq = tid & 3;
    int a_iy[2], a_ix[2];
    long long a_base[2];
    bool a_ok[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int row = (tid >> 2) + i * 64;
        const long long m = m0 + row;
        a_ok[i] = m < p.M;
        long long mm = a_ok[i] ? m : 0;
        const int ox = (int)(mm % p.Wo); mm /= p.Wo;
        const int oy = (int)(mm % p.Ho);
        const int n = (int)(mm / p.Ho);
        a_iy[i] = oy * p.stride - p.pad;
        a_ix[i] = ox * p.stride - p.pad;
        a_base[i] = (long long)n * p.Hi * p.Wi;
    }
    const int b_co = n0 + (tid >> 2);
    const bool b_ok = b_co < p.Cout;

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    Vec4<T> ra[2], rb;
    auto fetch = [&](int k0) {
        const int tap = k0 / p.Cin, c0 = k0 - tap * p.Cin;
        const int ky = tap / p.KW, kx = tap - ky * p.KW;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int iy = a_iy[i] + ky * p.dil, ix = a_ix[i] + kx * p.dil;
            if (a_ok[i] && iy >= 0 && iy < p.Hi && ix >= 0 && ix < p.Wi)
                ra[i] = *reinterpret_cast<const Vec4<T>*>(in + (a_base[i] + (long long)iy * p.Wi + ix) * p.Cin + c0 + kq * 4);
            else
#pragma unroll
                for (int j = 0; j < 4; ++j) ra[i].v[j] = from_f32<T>(0.f);
        }
        if (b_ok) rb = *reinterpret_cast<const Vec4<T>*>(w + (long long)b_co * p.K + k0 + kq * 4);
        else
#pragma unroll
            for (int j = 0; j < 4; ++j) rb.v[j] = from_f32<T>(0.f);
    };
    auto stash = [&]() {
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) As[kq * 4 + j][(tid >> 2) + i * 64] = to_f32(ra[i].v[j]);
#pragma unroll
        for (int j = 0; j < 4; ++j) Bs[kq * 4 + j][tid >> 2] = to_f32(rb.v[j]);
    };

    fetch(0);
    stash();
    __syncthreads();
    for (int k0 = 0; k0 < p.K; k0 += BK) {
        const bool more = k0 + BK < p.K;
        if (more) fetch(k0 + BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
        if (more) {
            stash();
            __syncthreads();
        }
    }

    // epilogue: scale/shift (folded BN + bias), residual, activation
    T* __restrict__ out = reinterpret_cast<T*>(p.out);
    const T* __restrict__ res = reinterpret_cast<const T*>(p.res);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long m = m0 + ty * 8 + i;
        if (m >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int co = n0 + tx * 4 + j;
            if (co >= p.Cout) continue;
            float v = acc[i][j];
            v = v * (p.scale ? p.scale[co] : 1.f) + (p.shift ? p.shift[co] : 0.f);
            if (res) v += to_f32(res[m * p.Cout + co]);
            if (p.act == ARSEG_ACT_RELU) v = fmaxf(v, 0.f);
            else if (p.act == ARSEG_ACT_PRELU) v = v > 0.f ? v : v * p.slope;
            out[m * p.ocs + p.oco + co] = from_f32<T>(v);
        }
    }
}

int conv_simt_launch(const arseg_conv_desc* d, cudaStream_t st) {
    ConvParams p;
    p.in = d->in; p.w = d->w; p.scale = d->scale; p.shift = d->shift; p.res = d->residual; p.out = d->out;
    p.N = d->N; p.Hi = d->Hi; p.Wi = d->Wi; p.Cin = d->Cin; p.Cout = d->Cout; p.KH = d->KH; p.KW = d->KW;
    p.stride = d->stride; p.pad = d->pad; p.dil = d->dil;
    p.Ho = (d->Hi + 2 * d->pad - d->dil * (d->KH - 1) - 1) / d->stride + 1;
    p.Wo = (d->Wi + 2 * d->pad - d->dil * (d->KW - 1) - 1) / d->stride + 1;
    p.ocs = d->out_cstride; p.oco = d->out_coff; p.act = d->act; p.slope = d->prelu_slope;
    p.M = (long long)d->N * p.Ho * p.Wo;
    p.K = d->KH * d->KW * d->Cin;
    ARSEG_REQUIRE(d->Cin % 16 == 0, "conv_simt: Cin=%d must be a multiple of 16", d->Cin);
    ARSEG_REQUIRE(p.Ho > 0 && p.Wo > 0, "conv_simt: empty output");
    dim3 grid((unsigned)ceil_div_ll(p.M, BM), ceil_div(d->Cout, BN));
    ARSEG_REQUIRE(grid.y <= 65535, "conv_simt: Cout too large");
    if (d->dtype == ARSEG_F32) conv_simt_kernel<float><<<grid, 256, 0, st>>>(p);
    else if (d->dtype == ARSEG_BF16) conv_simt_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(p);
    else if (d->dtype == ARSEG_F16) conv_simt_kernel<__half><<<grid, 256, 0, st>>>(p);
    else ARSEG_UNSUPPORTED("conv_simt: dtype %d", d->dtype);
    ARSEG_CHECK_LAUNCH("conv_simt");
    return ARSEG_OK;
}

int conv_tc_launch(const arseg_conv_desc* d, cudaStream_t st);  // conv_tc.cu
bool conv_tc_supported(const arseg_conv_desc* d);

}  // namespace arseg

using namespace arseg;

extern "C" int arseg_conv2d_nhwc(const arseg_conv_desc* d, arseg_stream_t stream) {
    ARSEG_REQUIRE(d && d->in && d->w && d->out, "conv2d: null pointer");
    ARSEG_REQUIRE(d->N > 0 && d->Hi > 0 && d->Wi > 0 && d->Cin > 0 && d->Cout > 0 && d->KH > 0 && d->KW > 0 &&
                      d->stride > 0 && d->dil > 0 && d->pad >= 0,
                  "conv2d: bad shape");
    ARSEG_REQUIRE(d->out_cstride >= d->Cout && d->out_coff >= 0 && d->out_coff + d->Cout <= d->out_cstride,
                  "conv2d: bad output slice (cstride=%d coff=%d Cout=%d)", d->out_cstride, d->out_coff, d->Cout);
    ARSEG_REQUIRE(d->act >= ARSEG_ACT_NONE && d->act <= ARSEG_ACT_PRELU, "conv2d: bad act %d", d->act);
    ARSEG_REQUIRE(!d->out_f32 || d->dtype == ARSEG_F32 || d->engine != ARSEG_CONV_SIMT_F32, "conv2d: out_f32 needs a tcgen05 engine");
    switch (d->engine) {
        case ARSEG_CONV_SIMT_F32:
            return conv_simt_launch(d, as_stream(stream));
        case ARSEG_CONV_TC_TF32:
        case ARSEG_CONV_TC_BF16:
        case ARSEG_CONV_TC_F16:
            if (!conv_tc_supported(d))
                ARSEG_UNSUPPORTED("conv2d: tcgen05 engine does not take this shape (Cin=%d Cout=%d stride=%d k=%dx%d)",
                                  d->Cin, d->Cout, d->stride, d->KH, d->KW);
            return conv_tc_launch(d, as_stream(stream));
        default:
            ARSEG_UNSUPPORTED("conv2d: unknown engine %d", d->engine);
    }
}
