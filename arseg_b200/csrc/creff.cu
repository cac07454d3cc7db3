// creff.cu -- fused MV-warp + CReFF + classifier kernel (exact fp32 SIMT version).
//
// One launch replaces, per non-keyframe (reference file:line):
//   evaluation.py:177-180  flow * Hf/H, bilinear (align_corners=True) resize to feature size   [float64]
//   evaluation.py:61-87    warpFeature: grid normalise (align_corners=True formula) + grid_sample (False)
//   model/attention.py:191 lr_up = bilinear(lr, (H,W), align_corners=True)
//   model/attention.py:194-197 V,K = dw3x3(warped hr)+bias ; Q = dw3x3(lr_up)+bias   (zero padding 1)
//   model/attention.py:199 S = similar_forward(Q,K,k,k)      (out-of-image taps: logit exactly 0)
//   model/attention.py:203 A = softmax over the k*k taps (out-of-image taps DO take mass)
//   model/attention.py:207 O = weighting_forward(V,A,k,k)    (out-of-image taps: value 0)
//   model/attention.py:210 fused = lr_up + O
//   model/pspnet.py:226-229 final_conv 1x1 (+bias) [+ identity-size interpolate] + LogSoftmax(dim=1)
//   evaluation.py:204      argmax over classes
//
// HBM traffic per launch (the algorithmic minimum): read hr once (+halo re-reads that hit
// L2), read lr, read the MV field, write fused p and the logits.  No intermediate (warped hr, lr_up, Q, K,
// V, S, A) ever goes to HBM.
//
// Work decomposition: CTA = TH x TW output pixels, one thread per pixel.  Channels are streamed in chunks
// of CC: pass A accumulates the k*k logits in registers, softmax in registers, pass B re-gathers the hr
// chunk (L2 hit) to form V, applies the attention, adds lr_up, writes fused p and accumulates the
// classifier.
#include "common.cuh"

namespace arseg {

constexpr int TH = 8, TW = 32, CC = 8, MAX_CLS = 32;

struct CreffParams {
    const float* hr; int hr_shared;
    const void* flow; int flow_dtype, Hm, Wm;
    const void* lr; int h, w;
    const float *wq, *bq, *wk, *bk, *wv, *bv, *wcls, *bcls;
    int ncls, log_softmax;
    float* out_p; float* out_logits; uint8_t* out_argmax;
    int N, C, H, W;
};

__device__ __forceinline__ double flow_raw(const void* flow, int dtype, size_t idx) {
    if (dtype == ARSEG_I16) return (double)reinterpret_cast<const int16_t*>(flow)[idx] / 4.0;  // camvid.py:625
    if (dtype == ARSEG_F64) return reinterpret_cast<const double*>(flow)[idx];
    return (double)reinterpret_cast<const float*>(flow)[idx];
}

// evaluation.py:177-180 at one feature pixel: (flow * Hf / Hm) resized bilinear/align_corners=True, in f64
__device__ __forceinline__ void flow_at(const CreffParams& p, int n, int fy, int fx, double& u, double& v) {
    const size_t base = (size_t)n * p.Hm * p.Wm;
    if (p.Hm == p.H && p.Wm == p.W) {
        const size_t i = (base + (size_t)fy * p.Wm + fx) * 2;
        u = flow_raw(p.flow, p.flow_dtype, i) * (double)p.H / (double)p.Hm;
        v = flow_raw(p.flow, p.flow_dtype, i + 1) * (double)p.H / (double)p.Hm;
        return;
    }
    const double sh = p.H > 1 ? (double)(p.Hm - 1) / (double)(p.H - 1) : 0.0;
    const double sw = p.W > 1 ? (double)(p.Wm - 1) / (double)(p.W - 1) : 0.0;
    const double ry = sh * fy, rx = sw * fx;
    int y0 = (int)ry, x0 = (int)rx;
    y0 = min(y0, p.Hm - 1); x0 = min(x0, p.Wm - 1);
    const int y1 = y0 + (y0 < p.Hm - 1 ? 1 : 0), x1 = x0 + (x0 < p.Wm - 1 ? 1 : 0);
    const double ly1 = ry - y0, ly0 = 1.0 - ly1, lx1 = rx - x0, lx0 = 1.0 - lx1;
    double r[2];
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
        const double s = (double)p.H / (double)p.Hm;
        const double a = flow_raw(p.flow, p.flow_dtype, (base + (size_t)y0 * p.Wm + x0) * 2 + ch) * (double)p.H / (double)p.Hm;
        const double b = flow_raw(p.flow, p.flow_dtype, (base + (size_t)y0 * p.Wm + x1) * 2 + ch) * (double)p.H / (double)p.Hm;
        const double c = flow_raw(p.flow, p.flow_dtype, (base + (size_t)y1 * p.Wm + x0) * 2 + ch) * (double)p.H / (double)p.Hm;
        const double d = flow_raw(p.flow, p.flow_dtype, (base + (size_t)y1 * p.Wm + x1) * 2 + ch) * (double)p.H / (double)p.Hm;
        (void)s;
        r[ch] = ly0 * (lx0 * a + lx1 * b) + ly1 * (lx0 * c + lx1 * d);
    }
    u = r[0]; v = r[1];
}

template <int K, int LR_LAYOUT, typename TLR>
__global__ void __launch_bounds__(TH * TW) creff_kernel(CreffParams p) {
    constexpr int R = K / 2, T = K * K;
    constexpr int HRH = TH + 2 * R + 2, HRW = TW + 2 * R + 2;      // warped-hr tile (K/V halo + dw halo)
    constexpr int KH_ = TH + 2 * R, KW_ = TW + 2 * R, KLD = KW_ + 1;  // K / V tile
    constexpr int LH = TH + 2, LW = TW + 2, LLD = LW + 1;           // lr_up tile (dw halo)
    extern __shared__ __align__(16) float smem[];
    float* s_hr = smem;                               // [CC][HRH][HRW]
    float* s_kv = s_hr + CC * HRH * HRW;              // [CC][KH_][KLD]
    float* s_lr = s_kv + CC * KH_ * KLD;              // [CC][LH][LLD]
    float2* s_pos = reinterpret_cast<float2*>(s_lr + CC * LH * LLD);  // [HRH*HRW] source position (ix,iy)
    float* s_wd = reinterpret_cast<float*>(s_pos + HRH * HRW);        // [3][CC][10] dw weights+bias of the chunk
    float* s_wc = s_wd + 3 * CC * 10;                 // [MAX_CLS][CC] classifier slice

    const int tid = threadIdx.x, tx = tid % TW, ty = tid / TW;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, n = blockIdx.z;
    const int H = p.H, W = p.W, C = p.C;
    const size_t plane = (size_t)H * W;
    const float* hr = p.hr + (p.hr_shared ? 0 : (size_t)n * C * plane);
    const bool has_flow = p.flow != nullptr;

    // ---- sampling positions of the warped-hr tile (channel independent) ----
    for (int i = tid; i < HRH * HRW; i += TH * TW) {
        const int fy = y0 - R - 1 + i / HRW, fx = x0 - R - 1 + i % HRW;
        float ix = -1e30f, iy = -1e30f;  // outside the image: dw-conv zero padding
        if (fy >= 0 && fy < H && fx >= 0 && fx < W) {
            if (has_flow) {
                double u, v;
                flow_at(p, n, fy, fx, u, v);
                warp_source_pos(fx, fy, u, v, W, H, ix, iy);
            } else {
                ix = (float)fx; iy = (float)fy;
            }
        }
        s_pos[i] = make_float2(ix, iy);
    }
    __syncthreads();

    // lr_up bilinear coefficients are recomputed per item (cheap fp32 math)
    const float lsh = resize_scale(p.h, H, ARSEG_RESIZE_BILINEAR_AC), lsw = resize_scale(p.w, W, ARSEG_RESIZE_BILINEAR_AC);
    const TLR* lr = reinterpret_cast<const TLR*>(p.lr);

    auto load_chunk = [&](int c0, const float* wkv, const float* bkv) {
        // (1) dw weights of this chunk: [0]=q, [1]=k or v (selected by caller), plus classifier slice
        for (int i = tid; i < CC * 10; i += TH * TW) {
            const int c = i / 10, t = i % 10;
            s_wd[i] = t < 9 ? p.wq[(size_t)(c0 + c) * 9 + t] : p.bq[c0 + c];
            s_wd[CC * 10 + i] = t < 9 ? wkv[(size_t)(c0 + c) * 9 + t] : bkv[c0 + c];
        }
        // (2) warped hr chunk: bilinear gather, zeros outside (grid_sample zeros padding)
        for (int i = tid; i < HRH * HRW; i += TH * TW) {
            const float2 ps = s_pos[i];
            if (ps.x < -1e29f) {
#pragma unroll
                for (int c = 0; c < CC; ++c) s_hr[c * HRH * HRW + i] = 0.f;
                continue;
            }
            if (!has_flow) {
                const size_t o = (size_t)(int)ps.y * W + (int)ps.x;
#pragma unroll
                for (int c = 0; c < CC; ++c) s_hr[c * HRH * HRW + i] = hr[(size_t)(c0 + c) * plane + o];
                continue;
            }
            const float fx = floorf(ps.x), fy = floorf(ps.y);
            const int xa = (int)fx, ya = (int)fy, xb = xa + 1, yb = ya + 1;
            const float wnw = ((fx + 1.f) - ps.x) * ((fy + 1.f) - ps.y), wne = (ps.x - fx) * ((fy + 1.f) - ps.y);
            const float wsw = ((fx + 1.f) - ps.x) * (ps.y - fy), wse = (ps.x - fx) * (ps.y - fy);
            const bool vxa = xa >= 0 && xa < W, vxb = xb >= 0 && xb < W, vya = ya >= 0 && ya < H, vyb = yb >= 0 && yb < H;
#pragma unroll
            for (int c = 0; c < CC; ++c) {
                const float* pl = hr + (size_t)(c0 + c) * plane;
                float acc = 0.f;
                if (vya && vxa) acc += pl[(size_t)ya * W + xa] * wnw;
                if (vya && vxb) acc += pl[(size_t)ya * W + xb] * wne;
                if (vyb && vxa) acc += pl[(size_t)yb * W + xa] * wsw;
                if (vyb && vxb) acc += pl[(size_t)yb * W + xb] * wse;
                s_hr[c * HRH * HRW + i] = acc;
            }
        }
        // (3) lr_up chunk (model/attention.py:191), zero outside the image (dw-conv padding)
        for (int i = tid; i < CC * LH * LW; i += TH * TW) {
            int c, pos;
            if (LR_LAYOUT == ARSEG_NHWC) { c = i % CC; pos = i / CC; } else { pos = i % (LH * LW); c = i / (LH * LW); }
            const int fy = y0 - 1 + pos / LW, fx = x0 - 1 + pos % LW;
            float v = 0.f;
            if (fy >= 0 && fy < H && fx >= 0 && fx < W) {
                int ya, yb, xa, xb; float lya, lyb, lxa, lxb;
                bilinear_src(lsh, fy, p.h, ARSEG_RESIZE_BILINEAR_AC, ya, yb, lya, lyb);
                bilinear_src(lsw, fx, p.w, ARSEG_RESIZE_BILINEAR_AC, xa, xb, lxa, lxb);
                float a, b, cc_, d;
                if (LR_LAYOUT == ARSEG_NHWC) {
                    const TLR* q = lr + (size_t)n * p.h * p.w * C + c0 + c;
                    a = to_f32(q[((size_t)ya * p.w + xa) * C]); b = to_f32(q[((size_t)ya * p.w + xb) * C]);
                    cc_ = to_f32(q[((size_t)yb * p.w + xa) * C]); d = to_f32(q[((size_t)yb * p.w + xb) * C]);
                } else {
                    const TLR* q = lr + ((size_t)n * C + c0 + c) * p.h * p.w;
                    a = to_f32(q[(size_t)ya * p.w + xa]); b = to_f32(q[(size_t)ya * p.w + xb]);
                    cc_ = to_f32(q[(size_t)yb * p.w + xa]); d = to_f32(q[(size_t)yb * p.w + xb]);
                }
                v = lya * (lxa * a + lxb * b) + lyb * (lxa * cc_ + lxb * d);
            }
            s_lr[c * LH * LLD + (pos / LW) * LLD + pos % LW] = v;
        }
        __syncthreads();
        // (4) K (or V) tile = dw3x3(warped hr)+bias inside the image, exactly 0 outside (attention zero padding)
        for (int i = tid; i < CC * KH_ * KW_; i += TH * TW) {
            const int c = i / (KH_ * KW_), pos = i % (KH_ * KW_);
            const int r = pos / KW_, q = pos % KW_;
            const int fy = y0 - R + r, fx = x0 - R + q;
            float v = 0.f;
            if (fy >= 0 && fy < H && fx >= 0 && fx < W) {
                const float* wd = s_wd + CC * 10 + c * 10;
                const float* s = s_hr + c * HRH * HRW + r * HRW + q;  // top-left of the 3x3 (tile origin is -R-1)
                v = wd[9];
#pragma unroll
                for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) v = fmaf(wd[dy * 3 + dx], s[dy * HRW + dx], v);
            }
            s_kv[c * KH_ * KLD + r * KLD + q] = v;
        }
        __syncthreads();
    };

    const int x = x0 + tx, y = y0 + ty;
    const bool live = x < W && y < H;

    // ---------------- pass A: logits ----------------
    float S[T];
#pragma unroll
    for (int t = 0; t < T; ++t) S[t] = 0.f;
    for (int c0 = 0; c0 < C; c0 += CC) {
        load_chunk(c0, p.wk, p.bk);
#pragma unroll 1
        for (int c = 0; c < CC; ++c) {
            const float* wd = s_wd + c * 10;
            const float* sl = s_lr + c * LH * LLD + ty * LLD + tx;
            float q = wd[9];
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) q = fmaf(wd[dy * 3 + dx], sl[dy * LLD + dx], q);
            const float* sk = s_kv + c * KH_ * KLD + ty * KLD + tx;
#pragma unroll
            for (int i = 0; i < K; ++i)
#pragma unroll
                for (int j = 0; j < K; ++j) S[i * K + j] = fmaf(q, sk[i * KLD + j], S[i * K + j]);
        }
        __syncthreads();
    }
    // ---------------- softmax over the k*k taps (model/attention.py:203) ----------------
    float mx = S[0];
#pragma unroll
    for (int t = 1; t < T; ++t) mx = fmaxf(mx, S[t]);
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < T; ++t) { S[t] = expf(S[t] - mx); sum += S[t]; }
    const float inv = 1.f / sum;
#pragma unroll
    for (int t = 0; t < T; ++t) S[t] *= inv;

    // ---------------- pass B: weighting + residual + classifier ----------------
    float logit[MAX_CLS];
#pragma unroll
    for (int j = 0; j < MAX_CLS; ++j) logit[j] = 0.f;
    const bool do_cls = p.wcls != nullptr;
    for (int c0 = 0; c0 < C; c0 += CC) {
        if (do_cls)
            for (int i = tid; i < p.ncls * CC; i += TH * TW) s_wc[i] = p.wcls[(size_t)(i / CC) * C + c0 + i % CC];
        load_chunk(c0, p.wv, p.bv);
#pragma unroll 1
        for (int c = 0; c < CC; ++c) {
            const float* sv = s_kv + c * KH_ * KLD + ty * KLD + tx;
            float o = 0.f;
#pragma unroll
            for (int i = 0; i < K; ++i)
#pragma unroll
                for (int j = 0; j < K; ++j) o = fmaf(S[i * K + j], sv[i * KLD + j], o);
            const float f = s_lr[c * LH * LLD + (ty + 1) * LLD + tx + 1] + o;   // model/attention.py:210
            if (live && p.out_p) p.out_p[((size_t)n * C + c0 + c) * plane + (size_t)y * W + x] = f;
            if (do_cls) {
#pragma unroll
                for (int j = 0; j < MAX_CLS; ++j)
                    if (j < p.ncls) logit[j] = fmaf(s_wc[j * CC + c], f, logit[j]);
            }
        }
        __syncthreads();
    }
    if (!do_cls || !live) return;
    float lmax = -INFINITY;
    int amax = 0;
#pragma unroll
    for (int j = 0; j < MAX_CLS; ++j)
        if (j < p.ncls) {
            logit[j] += p.bcls ? p.bcls[j] : 0.f;
            if (logit[j] > lmax) { lmax = logit[j]; amax = j; }   // first maximum, like torch.argmax
        }
    if (p.out_argmax) p.out_argmax[(size_t)n * plane + (size_t)y * W + x] = (uint8_t)amax;
    if (p.out_logits) {
        float lse = 0.f;
        if (p.log_softmax) {
#pragma unroll
            for (int j = 0; j < MAX_CLS; ++j)
                if (j < p.ncls) lse += expf(logit[j] - lmax);
            lse = logf(lse) + lmax;
        }
#pragma unroll
        for (int j = 0; j < MAX_CLS; ++j)
            if (j < p.ncls) p.out_logits[((size_t)n * p.ncls + j) * plane + (size_t)y * W + x] = logit[j] - lse;
    }
}

template <int K>
static size_t creff_smem_bytes() {
    constexpr int R = K / 2;
    constexpr int HRH = TH + 2 * R + 2, HRW = TW + 2 * R + 2, KH_ = TH + 2 * R, KLD = TW + 2 * R + 1, LH = TH + 2, LLD = TW + 3;
    return sizeof(float) * (CC * HRH * HRW + CC * KH_ * KLD + CC * LH * LLD + 3 * CC * 10 + MAX_CLS * CC) +
           sizeof(float2) * HRH * HRW;
}

template <int K, int LAYOUT, typename TLR>
static int creff_launch_t(const CreffParams& p, cudaStream_t st) {
    const size_t smem = creff_smem_bytes<K>();
    auto kern = creff_kernel<K, LAYOUT, TLR>;
    static bool configured[64] = {false};  // per device; benign race (idempotent attribute)
    int dev = 0;
    ARSEG_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        ARSEG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    dim3 grid(ceil_div(p.W, TW), ceil_div(p.H, TH), p.N);
    kern<<<grid, TH * TW, smem, st>>>(p);
    ARSEG_CHECK_LAUNCH("creff_fused");
    return ARSEG_OK;
}

template <int K>
static int creff_launch_k(const CreffParams& p, int layout, int dtype, cudaStream_t st) {
    if (layout == ARSEG_NCHW) return creff_launch_t<K, ARSEG_NCHW, float>(p, st);
    if (dtype == ARSEG_F32) return creff_launch_t<K, ARSEG_NHWC, float>(p, st);
    if (dtype == ARSEG_F16) return creff_launch_t<K, ARSEG_NHWC, __half>(p, st);
    return creff_launch_t<K, ARSEG_NHWC, __nv_bfloat16>(p, st);
}

bool creff_mma_supported(const arseg_creff_args* a);
// the tcgen05 engine: asked for by name, or (ABI 5 behaviour) ARSEG_CREFF_MMA_F16 with an f16 keyframe feature
static inline bool creff_uses_tc(const arseg_creff_args* a) {
    return a->engine == ARSEG_CREFF_TCGEN05 || (a->engine == ARSEG_CREFF_MMA_F16 && a->C == 64 && a->hr_dtype == ARSEG_F16);
}
int creff_mma_launch(const arseg_creff_args* a, cudaStream_t st);
bool creff_wide_supported(const arseg_creff_args* a);                                   // creff_wide.cu
int creff_wide_launch(const arseg_creff_args* a, void* ws, size_t ws_bytes, cudaStream_t st);
size_t creff_wide_workspace_bytes(int N, int C, int H, int W);
size_t creff_tc_workspace_bytes(int N, int H, int W, int k);                                   // creff_tc.cu

}  // namespace arseg

using namespace arseg;

extern "C" int arseg_creff_fused_fwd(const arseg_creff_args* a, arseg_stream_t stream) {
    ARSEG_REQUIRE(a && a->hr && a->lr && a->wq && a->bq && a->wk && a->bk && a->wv && a->bv, "creff: null pointer");
    ARSEG_REQUIRE(a->N > 0 && a->C > 0 && a->H > 0 && a->W > 0 && a->h > 0 && a->w > 0, "creff: bad shape");
    ARSEG_REQUIRE(a->C % CC == 0, "creff: C=%d must be a multiple of %d", a->C, CC);
    ARSEG_REQUIRE(a->H <= 65535 * TH && a->N <= 65535, "creff: H or N too large");
    ARSEG_REQUIRE(a->out_p || a->out_logits || a->out_argmax, "creff: no output requested");
    if (a->wcls) ARSEG_REQUIRE(a->ncls > 0 && a->ncls <= MAX_CLS, "creff: ncls=%d unsupported (1..%d)", a->ncls, MAX_CLS);
    else ARSEG_REQUIRE(!a->out_logits && !a->out_argmax, "creff: logits/argmax requested without classifier weights");
    if (a->flow) {
        ARSEG_REQUIRE(a->Hm > 0 && a->Wm > 0, "creff: bad MV field size");
        ARSEG_REQUIRE(a->flow_dtype == ARSEG_I16 || a->flow_dtype == ARSEG_F32 || a->flow_dtype == ARSEG_F64,
                      "creff: flow dtype %d", a->flow_dtype);
    }
    ARSEG_REQUIRE(a->lr_layout == ARSEG_NCHW || a->lr_layout == ARSEG_NHWC, "creff: lr layout %d", a->lr_layout);
    if (a->lr_layout == ARSEG_NCHW) ARSEG_REQUIRE(a->lr_dtype == ARSEG_F32, "creff: NCHW lr must be fp32");
    else ARSEG_REQUIRE(a->lr_dtype == ARSEG_F32 || a->lr_dtype == ARSEG_BF16 || a->lr_dtype == ARSEG_F16, "creff: lr dtype %d", a->lr_dtype);
    ARSEG_REQUIRE(a->hr_dtype == ARSEG_F32 || a->hr_dtype == ARSEG_F16, "creff: hr dtype %d", a->hr_dtype);
    if (a->hr_dtype == ARSEG_F16)
        ARSEG_REQUIRE((a->engine == ARSEG_CREFF_MMA_F16 || a->engine == ARSEG_CREFF_TCGEN05) && a->hr_layout == ARSEG_NHWC && a->C == 64,
                      "creff: an f16 keyframe feature needs the tcgen05 engine, NHWC and C = 64");
    ARSEG_REQUIRE(a->phase == ARSEG_CREFF_PHASE_ALL || a->phase == ARSEG_CREFF_PHASE_PREPASS || a->phase == ARSEG_CREFF_PHASE_MAIN, "creff: phase %d", a->phase);
    cudaStream_t st = as_stream(stream);
    if (creff_uses_tc(a)) {
        if (!creff_mma_supported(a) || a->lr_dtype != ARSEG_F16 || a->k > 7)
            ARSEG_UNSUPPORTED("creff: the tcgen05 engine needs C = 64 (got %d), NHWC hr, NHWC f16 lr (dtype %d), k in {3,5,7} (got %d), ncls <= 32", a->C, a->lr_dtype, a->k);
        return creff_mma_launch(a, st);
    }
    if (a->phase == ARSEG_CREFF_PHASE_PREPASS) return ARSEG_OK;      // only the tcgen05 engine has a pre-pass
    if (a->engine == ARSEG_CREFF_MMA_F16) {
        if (creff_wide_supported(a)) return creff_wide_launch(a, a->workspace, a->workspace_bytes, st);
        if (!creff_mma_supported(a))
            ARSEG_UNSUPPORTED("creff: the MMA engine needs C a multiple of 64 up to 1024 (got %d), NHWC hr and lr, k in {3,5,7,9}, ncls <= 32", a->C);
        return creff_mma_launch(a, st);
    }
    ARSEG_REQUIRE(a->engine == ARSEG_CREFF_EXACT_F32, "creff: unknown engine %d", a->engine);
    if (a->hr_layout != ARSEG_NCHW) ARSEG_UNSUPPORTED("creff: the exact fp32 engine takes hr in NCHW");
    CreffParams p;
    p.hr = reinterpret_cast<const float*>(a->hr); p.hr_shared = a->hr_shared; p.flow = a->flow; p.flow_dtype = a->flow_dtype; p.Hm = a->Hm; p.Wm = a->Wm;
    p.lr = a->lr; p.h = a->h; p.w = a->w;
    p.wq = a->wq; p.bq = a->bq; p.wk = a->wk; p.bk = a->bk; p.wv = a->wv; p.bv = a->bv; p.wcls = a->wcls; p.bcls = a->bcls;
    p.ncls = a->ncls; p.log_softmax = a->log_softmax; p.out_p = a->out_p; p.out_logits = a->out_logits;
    p.out_argmax = a->out_argmax; p.N = a->N; p.C = a->C; p.H = a->H; p.W = a->W;
    switch (a->k) {
        case 3: return creff_launch_k<3>(p, a->lr_layout, a->lr_dtype, st);
        case 5: return creff_launch_k<5>(p, a->lr_layout, a->lr_dtype, st);
        case 7: return creff_launch_k<7>(p, a->lr_layout, a->lr_dtype, st);
        case 9: return creff_launch_k<9>(p, a->lr_layout, a->lr_dtype, st);
        default: ARSEG_UNSUPPORTED("creff: window k=%d (supported 3,5,7,9)", a->k);
    }
}

extern "C" size_t arseg_creff_workspace_bytes(const arseg_creff_args* a) {
    if (!a) return 0;
    if (creff_uses_tc(a)) return creff_tc_workspace_bytes(a->N, a->H, a->W, a->k);   // lr_up gather records + MV-warped keyframe rows
    if (a->engine != ARSEG_CREFF_MMA_F16) return 0;
    if (!creff_wide_supported(a)) return 0;
    return creff_wide_workspace_bytes(a->N, a->C, a->H, a->W);
}
