// post.cu -- evaluation.py:197-209 post-processing: logits resize -> (softmax, monotone) -> argmax -> histogram.
#include "common.cuh"

namespace arseg {

constexpr int POST_MAX_CLS = 32;

__global__ void resize_argmax_kernel(const float* __restrict__ logits, float* __restrict__ out_logits,
                                     uint8_t* __restrict__ out_argmax, int ncls, int Hi, int Wi, int Ho, int Wo,
                                     int mode, float sh, float sw) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, n = blockIdx.z;
    if (x >= Wo) return;
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    if (mode == ARSEG_RESIZE_NEAREST) {
        y0 = y1 = nearest_src(sh, y, Hi); x0 = x1 = nearest_src(sw, x, Wi);
        ly0 = lx0 = 1.f; ly1 = lx1 = 0.f;
    } else {
        bilinear_src(sh, y, Hi, mode, y0, y1, ly0, ly1);
        bilinear_src(sw, x, Wi, mode, x0, x1, lx0, lx1);
    }
    float best = -INFINITY;
    int arg = 0;
    for (int c = 0; c < ncls; ++c) {
        const float* s = logits + ((size_t)n * ncls + c) * Hi * Wi;
        float v;
        if (mode == ARSEG_RESIZE_NEAREST) v = s[(size_t)y0 * Wi + x0];
        else v = __fmaf_rn(ly1, __fmaf_rn(lx1, s[(size_t)y1 * Wi + x1], __fmul_rn(lx0, s[(size_t)y1 * Wi + x0])),
                           __fmul_rn(ly0, __fmaf_rn(lx1, s[(size_t)y0 * Wi + x1], __fmul_rn(lx0, s[(size_t)y0 * Wi + x0]))));
        if (out_logits) out_logits[(((size_t)n * ncls + c) * Ho + y) * Wo + x] = v;
        if (v > best) { best = v; arg = c; }
    }
    if (out_argmax) out_argmax[((size_t)n * Ho + y) * Wo + x] = (uint8_t)arg;
}

// Class maps only (evaluation.py:201-204, what the evaluation loop consumes): one thread per output column walks RSEG
// output rows.  ATen's bilinear formula is horizontal-first -- h0 * (w0 p00 + w1 p01) + h1 * (w0 p10 + w1 p11) -- so the
// two horizontally interpolated source rows of all classes stay in registers and are re-loaded only when the output row
// crosses into the next source row (every ~Ho/Hi rows): ~4 loads per class per source-row change instead of 4 per class
// per pixel.  Same expression tree as resize_argmax_kernel, spelled with explicit roundings in both.
constexpr int RA_RSEG = 32;
template <int NCMAX>
__global__ void __launch_bounds__(128) resize_argmax_rows_kernel(const float* __restrict__ logits, uint8_t* __restrict__ out_argmax, int ncls,
                                                                 int Hi, int Wi, int Ho, int Wo, int mode, float sh, float sw) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int ya = blockIdx.y * RA_RSEG, n = blockIdx.z;
    if (x >= Wo) return;
    int x0, x1;
    float lx0, lx1;
    bilinear_src(sw, x, Wi, mode, x0, x1, lx0, lx1);
    const float* src = logits + (size_t)n * ncls * Hi * Wi;
    float h0[NCMAX], h1[NCMAX];
    int cy0 = -1, cy1 = -1;
    const int yb = min(ya + RA_RSEG, Ho);
    for (int y = ya; y < yb; ++y) {
        int y0, y1;
        float ly0, ly1;
        bilinear_src(sh, y, Hi, mode, y0, y1, ly0, ly1);
        if (y0 != cy0) {
            if (y0 == cy1) {
#pragma unroll
                for (int c = 0; c < NCMAX; ++c) h0[c] = h1[c];
            } else {
#pragma unroll
                for (int c = 0; c < NCMAX; ++c)
                    if (c < ncls) {
                        const float* s = src + ((size_t)c * Hi + y0) * Wi;
                        h0[c] = __fmaf_rn(lx1, __ldg(s + x1), __fmul_rn(lx0, __ldg(s + x0)));
                    }
            }
            cy0 = y0;
        }
        if (y1 != cy1) {
            if (y1 == y0) {
#pragma unroll
                for (int c = 0; c < NCMAX; ++c) h1[c] = h0[c];
            } else {
#pragma unroll
                for (int c = 0; c < NCMAX; ++c)
                    if (c < ncls) {
                        const float* s = src + ((size_t)c * Hi + y1) * Wi;
                        h1[c] = __fmaf_rn(lx1, __ldg(s + x1), __fmul_rn(lx0, __ldg(s + x0)));
                    }
            }
            cy1 = y1;
        }
        float best = -INFINITY;
        int arg = 0;
#pragma unroll
        for (int c = 0; c < NCMAX; ++c)
            if (c < ncls) {
                const float v = __fmaf_rn(ly1, h1[c], __fmul_rn(ly0, h0[c]));
                if (v > best) { best = v; arg = c; }
            }
        out_argmax[((size_t)n * Ho + y) * Wo + x] = (uint8_t)arg;
    }
}

// nn.LogSoftmax over dim 1 of an NCHW tensor (model/pspnet.py:122,229), in place allowed
__global__ void log_softmax_nchw_kernel(const float* __restrict__ in, float* __restrict__ out, int ncls, long long plane) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int n = blockIdx.y;
    if (i >= plane) return;
    const float* s = in + (size_t)n * ncls * plane + i;
    float* d = out + (size_t)n * ncls * plane + i;
    float mx = -INFINITY;
    for (int c = 0; c < ncls; ++c) mx = fmaxf(mx, s[(size_t)c * plane]);
    float sum = 0.f;
    for (int c = 0; c < ncls; ++c) sum += expf(s[(size_t)c * plane] - mx);
    const float lse = logf(sum) + mx;
    for (int c = 0; c < ncls; ++c) d[(size_t)c * plane] = s[(size_t)c * plane] - lse;
}

__global__ void confusion_hist_kernel(const uint8_t* __restrict__ pred, const int64_t* __restrict__ label,
                                      unsigned long long* __restrict__ hist, long long npix, int ncls, int ignore) {
    __shared__ unsigned int s_h[POST_MAX_CLS * POST_MAX_CLS];
    const int bins = ncls * ncls;
    for (int i = threadIdx.x; i < bins; i += blockDim.x) s_h[i] = 0;
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
        const long long l = label[i];
        if (l != ignore && l >= 0 && l < ncls) atomicAdd(&s_h[(int)l * ncls + pred[i]], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < bins; i += blockDim.x)
        if (s_h[i]) atomicAdd(&hist[i], (unsigned long long)s_h[i]);
}

}  // namespace arseg

using namespace arseg;

extern "C" {

int arseg_resize_argmax_nchw(const float* logits, float* out_logits, uint8_t* out_argmax, int N, int ncls, int Hi,
                             int Wi, int Ho, int Wo, int mode, arseg_stream_t stream) {
    ARSEG_REQUIRE(logits && (out_logits || out_argmax), "resize_argmax: null pointer");
    ARSEG_REQUIRE(N > 0 && ncls > 0 && ncls <= 255 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0 && Ho <= 65535 && N <= 65535,
                  "resize_argmax: bad shape");
    ARSEG_REQUIRE(mode >= 0 && mode <= 2, "resize_argmax: bad mode");
    const float sh = resize_scale(Hi, Ho, mode), sw = resize_scale(Wi, Wo, mode);
    if (!out_logits && mode != ARSEG_RESIZE_NEAREST && ncls <= 32 && ceil_div(Ho, RA_RSEG) <= 65535) {
        dim3 grid(ceil_div(Wo, 128), ceil_div(Ho, RA_RSEG), N);
        if (ncls <= 20) resize_argmax_rows_kernel<20><<<grid, 128, 0, as_stream(stream)>>>(logits, out_argmax, ncls, Hi, Wi, Ho, Wo, mode, sh, sw);
        else resize_argmax_rows_kernel<32><<<grid, 128, 0, as_stream(stream)>>>(logits, out_argmax, ncls, Hi, Wi, Ho, Wo, mode, sh, sw);
        ARSEG_CHECK_LAUNCH("resize_argmax_rows");
        return ARSEG_OK;
    }
    dim3 grid(ceil_div(Wo, 128), Ho, N);
    resize_argmax_kernel<<<grid, 128, 0, as_stream(stream)>>>(logits, out_logits, out_argmax, ncls, Hi, Wi, Ho, Wo, mode, sh, sw);
    ARSEG_CHECK_LAUNCH("resize_argmax");
    return ARSEG_OK;
}

int arseg_log_softmax_nchw(const float* in, float* out, int N, int ncls, int H, int W, arseg_stream_t stream) {
    ARSEG_REQUIRE(in && out && N > 0 && ncls > 0 && H > 0 && W > 0 && N <= 65535, "log_softmax: bad args");
    const long long plane = (long long)H * W;
    dim3 grid((unsigned)ceil_div_ll(plane, 256), N);
    log_softmax_nchw_kernel<<<grid, 256, 0, as_stream(stream)>>>(in, out, ncls, plane);
    ARSEG_CHECK_LAUNCH("log_softmax");
    return ARSEG_OK;
}

int arseg_confusion_hist(const uint8_t* pred, const int64_t* label, long long* hist, long long npix, int ncls,
                         int ignore_label, arseg_stream_t stream) {
    ARSEG_REQUIRE(pred && label && hist && npix > 0, "confusion_hist: bad args");
    ARSEG_REQUIRE(ncls > 0 && ncls <= POST_MAX_CLS, "confusion_hist: ncls=%d unsupported", ncls);
    long long g = ceil_div_ll(npix, 256 * 8);
    if (g > sm_count() * 8) g = sm_count() * 8;
    confusion_hist_kernel<<<(int)g, 256, 0, as_stream(stream)>>>(pred, label, (unsigned long long*)hist, npix, ncls, ignore_label);
    ARSEG_CHECK_LAUNCH("confusion_hist");
    return ARSEG_OK;
}

}  // extern "C"
