// creff_wide.cu -- MV-warp + CReFF + classifier on tensor cores for C = 64 m channels (m >= 2), sm_100a.
//
// Same contract and arithmetic as creff_march.cu (reference: evaluation.py:177-183 MV rescale + warpFeature,
// model/attention.py:184-213 MyAttention.forward, model/pspnet.py:226-229 / model/bisenet.py:565-575 /
// model/pspnet_semseg.py:237-250 final_conv, evaluation.py:204 argmax), for the two architectures whose feature p has
// more than 64 channels on a 1/8-resolution map: BiSeNet-18 (C = 256 @90x120) and Cityscapes PSPNet-18 (C = 512 @128x256).
//
// The softmax couples all channels of a pixel (S = sum over channels), so the one-pass column march of the C = 64 engine
// would need m times its shared memory.  These maps are small (a few MB per operand), so the work is split in two
// launches (plus a tiny one that computes the per-pixel gather records once) with Q, K, V in fp16 and the residual in
// fp32 going through HBM / L2 once:
//   1. creff_wide_prep_kernel: CTA = 8x16 pixels x 64 channels.  Per position of the 10x18 halo tile one thread does the
//      f64 MV arithmetic and leaves a gather record (the same 2x2-block records as the march engine); half-warps gather
//      the MV-warped hr tile and the lr_up tile into shared memory (fp32); the three depthwise 3x3 convolutions run from
//      there: K, V, Q -> fp16 NHWC, lr_up centre -> fp32 NHWC (the residual).
//   2. creff_wide_attn_kernel: CTA = 4 warps = a 4x16-pixel strip, one 4x4 block per warp (the C role of the march
//      engine: same fragment layouts, masks, softmax, ones-column row sums).  Channel chunks of 64 stream through
//      double-buffered shared-memory tiles filled with cp.async (zero fill outside the image = the attention zero
//      padding): phase 1 accumulates S = sum_c Q_c K_c^T in registers, phase 2 is the softmax, phase 3 forms
//      O_c = P V_c per chunk, adds the residual, writes the fused p chunk and accumulates the classifier MMA.
#include "creff_mma_common.cuh"
#include <cstdlib>

namespace arseg {

struct WideParams {
    CreffMmaParams b;
    __half *Q, *K, *V;        // [N,H,W,C] fp16
    float* R;                 // [N,H,W,C] fp32 residual (lr_up)
    float4* rec_w;            // [N,H,W][2: hr, lr] block weights of the gather records
    uint4* rec_a;             // [N,H,W][2] {block byte offset lo, hi (channel 0 of the block's top-left pixel), row stride, valid}
    int lr_dtype;
};

// ---------------------------------------------------------------------------------------------
// launch 1: gather + depthwise convolutions
// ---------------------------------------------------------------------------------------------
constexpr int WP_TH = 8, WP_TW = 16, WP_HH = WP_TH + 2, WP_HW = WP_TW + 2, WP_NP = WP_HH * WP_HW;   // 10 x 18 = 180 positions
constexpr int WP_THREADS = 256;
constexpr int WP_GJ = 3;                 // gather positions in flight per half-warp
constexpr size_t WP_SMEM = (size_t)2 * WP_NP * 256 + (size_t)2 * WP_NP * 32 + 3 * 10 * 64 * 4;

__device__ __forceinline__ void w_block_of(const PosRec& r, int Wimg, int Himg, float4& w, int& bx, int& by) {
    w = r.w; bx = r.cx; by = r.cy;
    if (!((r.info >> 1) & 1)) {
        const float nn = w.x + w.y, ss = w.z + w.w;
        if (bx > 0 && bx == Wimg - 1) { bx -= 1; w.x = 0.f; w.y = nn; w.z = 0.f; w.w = ss; }
        else { w.x = nn; w.y = 0.f; w.z = ss; w.w = 0.f; }
    }
    if (!(r.info & 1)) {
        const float ww = w.x + w.z, ee = w.y + w.w;
        if (by > 0 && by == Himg - 1) { by -= 1; w.x = 0.f; w.y = 0.f; w.z = ww; w.w = ee; }
        else { w.x = ww; w.y = ee; w.z = 0.f; w.w = 0.f; }
    }
}

// launch 0: the gather records, once per pixel (the prep CTAs of all C/64 channel groups and of neighbouring tiles share
// them; the hr record carries the f64 MV arithmetic of evaluation.py:177-183, incl. the bilinear resize of the MV field)
template <typename TLR>
__global__ void __launch_bounds__(256) creff_wide_rec_kernel(WideParams q) {
    const CreffMmaParams& p = q.b;
    constexpr int LR_ES = (int)sizeof(TLR);
    const long long total = (long long)p.N * p.H * p.W * 2;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int src = (int)(i & 1);
    long long r = i >> 1;
    const int fx = (int)(r % p.W); r /= p.W;
    const int fy = (int)(r % p.H);
    const int n = (int)(r / p.H);
    const float lsh = resize_scale(p.h, p.H, ARSEG_RESIZE_BILINEAR_AC), lsw = resize_scale(p.w, p.W, ARSEG_RESIZE_BILINEAR_AC);
    const PosRec rec = src == 0 ? pos_hr(p, n, fy, fx) : pos_lr(p, lsh, lsw, fy, fx);
    float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
    unsigned long long off = 0;
    uint32_t rs = src == 0 ? (uint32_t)p.W * p.C * 4 : (uint32_t)p.w * p.C * LR_ES;
    if (rec.info >= 0) {
        int bx, by;
        if (src == 0) { w_block_of(rec, p.W, p.H, w, bx, by); off = ((unsigned long long)by * p.W + bx) * p.C * 4; }
        else { w_block_of(rec, p.w, p.h, w, bx, by); off = ((unsigned long long)by * p.w + bx) * p.C * LR_ES; }
    }
    q.rec_w[i] = w;
    q.rec_a[i] = make_uint4((uint32_t)off, (uint32_t)(off >> 32), rs, (uint32_t)src);
}

template <typename TLR>
__global__ void __launch_bounds__(WP_THREADS) creff_wide_prep_kernel(WideParams q) {
    const CreffMmaParams& p = q.b;
    extern __shared__ __align__(16) uint8_t wsm[];
    float* s_hr = reinterpret_cast<float*>(wsm);                         // [180][64]
    float* s_lr = s_hr + WP_NP * 64;                                     // [180][64]
    float4* s_w = reinterpret_cast<float4*>(s_lr + WP_NP * 64);          // [2][180] block weights (hr, lr)
    uint4* s_a = reinterpret_cast<uint4*>(s_w + 2 * WP_NP);              // [2][180] {address lo, hi, row stride, valid}
    float* s_dw = reinterpret_cast<float*>(s_a + 2 * WP_NP);             // [3: k, v, q][10][64]
    constexpr int LR_ES = (int)sizeof(TLR);
    const int tid = threadIdx.x, C = p.C;
    const int tiles_x = (p.W + WP_TW - 1) / WP_TW;
    const int x0 = (blockIdx.x % tiles_x) * WP_TW, y0 = (blockIdx.x / tiles_x) * WP_TH;
    const int c0 = blockIdx.y * 64, n = blockIdx.z;
    const char* const hrb = reinterpret_cast<const char*>(p.hr + (p.hr_shared ? 0 : (size_t)n * p.H * p.W * C) + c0);
    const char* const lrb = reinterpret_cast<const char*>(reinterpret_cast<const TLR*>(p.lr) + (size_t)n * p.h * p.w * C + c0);
    const float lsh = resize_scale(p.h, p.H, ARSEG_RESIZE_BILINEAR_AC), lsw = resize_scale(p.w, p.W, ARSEG_RESIZE_BILINEAR_AC);

    for (int i = tid; i < 3 * 10 * 64; i += WP_THREADS) {
        const int cv = i / 640, tp = (i / 64) % 10, ch = i % 64;
        const float* w = cv == 0 ? p.wk : (cv == 1 ? p.wv : p.wq);
        const float* b = cv == 0 ? p.bk : (cv == 1 ? p.bv : p.bq);
        s_dw[i] = tp < 9 ? __ldg(w + (size_t)(c0 + ch) * 9 + tp) : __ldg(b + c0 + ch);
    }
    // gather records of the halo tile (computed once per pixel by creff_wide_rec_kernel); positions outside the image are
    // zero-weight records (depthwise zero padding); the block offset becomes an address of this frame / channel group
    for (int i = tid; i < 2 * WP_NP; i += WP_THREADS) {
        const int src = i / WP_NP, pz = i - src * WP_NP;
        const int fy = y0 - 1 + pz / WP_HW, fx = x0 - 1 + pz % WP_HW;
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        uint4 a = make_uint4(0u, 0u, src == 0 ? (uint32_t)p.W * C * 4 : (uint32_t)p.w * C * LR_ES, (uint32_t)src);
        if (fy >= 0 && fy < p.H && fx >= 0 && fx < p.W) {
            const size_t ri = ((((size_t)n * p.H + fy) * p.W + fx) << 1) + src;
            w = __ldg(q.rec_w + ri);
            a = __ldg(q.rec_a + ri);
        }
        const unsigned long long au = reinterpret_cast<unsigned long long>(src == 0 ? hrb : lrb) + (((unsigned long long)a.y << 32) | a.x);
        s_w[i] = w;
        s_a[i] = make_uint4((uint32_t)au, (uint32_t)(au >> 32), a.z, (uint32_t)src);
    }
    __syncthreads();
    // gather: half-warp per position, 4 channels per lane, WP_GJ positions' loads in flight per half-warp
    {
        const int hw = tid >> 4, cl = tid & 15;
        constexpr int NHW = WP_THREADS / 16;
        for (int i0 = hw; i0 < 2 * WP_NP; i0 += NHW * WP_GJ) {
            float4 tp[WP_GJ][4];
#pragma unroll
            for (int u = 0; u < WP_GJ; ++u) {
                const int i = i0 + u * NHW;
                if (i < 2 * WP_NP) {
                    const uint4 id = s_a[i];
                    const char* a0 = reinterpret_cast<const char*>(((unsigned long long)id.y << 32) | id.x);
                    if (LR_ES == 4 || id.w == 0) {
                        a0 += 16 * cl;
                        const char* a1 = a0 + id.z;
                        tp[u][0] = ld4(reinterpret_cast<const float*>(a0)); tp[u][1] = ld4(reinterpret_cast<const float*>(a0 + (size_t)C * 4));
                        tp[u][2] = ld4(reinterpret_cast<const float*>(a1)); tp[u][3] = ld4(reinterpret_cast<const float*>(a1 + (size_t)C * 4));
                    } else {
                        a0 += LR_ES * 4 * cl;
                        const char* a1 = a0 + id.z;
                        tp[u][0] = ld4(reinterpret_cast<const TLR*>(a0)); tp[u][1] = ld4(reinterpret_cast<const TLR*>(a0 + (size_t)C * LR_ES));
                        tp[u][2] = ld4(reinterpret_cast<const TLR*>(a1)); tp[u][3] = ld4(reinterpret_cast<const TLR*>(a1 + (size_t)C * LR_ES));
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < WP_GJ; ++u) {
                const int i = i0 + u * NHW;
                if (i < 2 * WP_NP) {
                    const float4 w = s_w[i];
                    float4 v;
                    v.x = tp[u][0].x * w.x + tp[u][1].x * w.y + tp[u][2].x * w.z + tp[u][3].x * w.w;
                    v.y = tp[u][0].y * w.x + tp[u][1].y * w.y + tp[u][2].y * w.z + tp[u][3].y * w.w;
                    v.z = tp[u][0].z * w.x + tp[u][1].z * w.y + tp[u][2].z * w.z + tp[u][3].z * w.w;
                    v.w = tp[u][0].w * w.x + tp[u][1].w * w.y + tp[u][2].w * w.z + tp[u][3].w * w.w;
                    float* dst = (i < WP_NP ? s_hr + i * 64 : s_lr + (i - WP_NP) * 64) + 4 * cl;
                    *reinterpret_cast<float4*>(dst) = v;
                }
            }
        }
    }
    __syncthreads();
    // depthwise 3x3 convolutions (model/attention.py:194-197): warp = one output row of the tile, lane = channel pair,
    // marching along x with a 3-column register window per source (3 new LDS.64 per source and column); weights in registers
    {
        const int lane = tid & 31, py = tid >> 5, y = y0 + py;
        if (y < p.H) {
            float2 wk[10], wv[10], wq[10];
#pragma unroll
            for (int i = 0; i < 10; ++i) {
                wk[i] = *reinterpret_cast<const float2*>(s_dw + i * 64 + 2 * lane);
                wv[i] = *reinterpret_cast<const float2*>(s_dw + (10 + i) * 64 + 2 * lane);
                wq[i] = *reinterpret_cast<const float2*>(s_dw + (20 + i) * 64 + 2 * lane);
            }
            const float* hrow = s_hr + (py * WP_HW) * 64 + 2 * lane;       // halo rows py .. py+2, column 0
            const float* lrow = s_lr + (py * WP_HW) * 64 + 2 * lane;
            float2 hwin[3][3], lwin[3][3];                                   // [row][slot], slot (x + d) % 3 holds column x + d
#pragma unroll
            for (int cc = 0; cc < 2; ++cc)
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    hwin[r][cc] = *reinterpret_cast<const float2*>(hrow + (r * WP_HW + cc) * 64);
                    lwin[r][cc] = *reinterpret_cast<const float2*>(lrow + (r * WP_HW + cc) * 64);
                }
            auto dw9 = [](const float2 (&w)[10], const float2 (&win)[3][3], int sa, int sb, int sc) {
                float2 a0 = __ffma2_rn(w[0], win[0][sa], w[9]), a1 = __fmul2_rn(w[3], win[1][sa]), a2 = __fmul2_rn(w[6], win[2][sa]);
                a0 = __ffma2_rn(w[1], win[0][sb], a0); a1 = __ffma2_rn(w[4], win[1][sb], a1); a2 = __ffma2_rn(w[7], win[2][sb], a2);
                a0 = __ffma2_rn(w[2], win[0][sc], a0); a1 = __ffma2_rn(w[5], win[1][sc], a1); a2 = __ffma2_rn(w[8], win[2][sc], a2);
                return __fadd2_rn(__fadd2_rn(a0, a1), a2);
            };
            const size_t orow = (((size_t)n * p.H + y) * p.W + x0) * C + c0 + 2 * lane;
#pragma unroll
            for (int x = 0; x < WP_TW; ++x) {
                const int sa = x % 3, sb = (x + 1) % 3, sc = (x + 2) % 3;
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    hwin[r][sc] = *reinterpret_cast<const float2*>(hrow + (r * WP_HW + x + 2) * 64);
                    lwin[r][sc] = *reinterpret_cast<const float2*>(lrow + (r * WP_HW + x + 2) * 64);
                }
                if (x0 + x < p.W) {
                    const float2 kk = dw9(wk, hwin, sa, sb, sc), vv = dw9(wv, hwin, sa, sb, sc), qq = dw9(wq, lwin, sa, sb, sc);
                    const size_t o = orow + (size_t)x * C;
                    *reinterpret_cast<uint32_t*>(q.K + o) = pack_h2_sat(kk.x, kk.y);
                    *reinterpret_cast<uint32_t*>(q.V + o) = pack_h2_sat(vv.x, vv.y);
                    *reinterpret_cast<uint32_t*>(q.Q + o) = pack_h2_sat(qq.x, qq.y);
                    *reinterpret_cast<float2*>(q.R + o) = lwin[1][sb];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// launch 2: window attention + classifier
// ---------------------------------------------------------------------------------------------
constexpr int WA_SW = 16, WA_THREADS = 128;
constexpr int WA_CLS_PAD = 8;            // f16 pad per classifier-weight row (bank spread)

template <int K> struct WACfg {
    static constexpr int R = K / 2;
    static constexpr int KVC = WA_SW + K - 1, KVR = 4 + K - 1;            // key tile: columns x rows
    static constexpr int WN = K + 3, NK = WN * WN;
    static constexpr int NT16 = (NK + 15) / 16;
    static constexpr int NT8Q = (NK + 7) / 8;
    static constexpr uint32_t ROWB = KVC * 128;
    static constexpr size_t KV_BYTES = (size_t)KVR * KVC * 128;
    static constexpr size_t Q_BYTES = 64 * 128;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ float wa_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float wa_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int K, int NCT>
__global__ void __launch_bounds__(WA_THREADS, 2) creff_wide_attn_kernel(WideParams q) {
    using Cf = WACfg<K>;
    const CreffMmaParams& p = q.b;
    extern __shared__ __align__(1024) uint8_t asm_[];
    uint8_t* s_kv = asm_;                                       // [2][KV_BYTES]: K tiles in phase 1, V tiles in phase 3
    uint8_t* s_q = s_kv + 2 * Cf::KV_BYTES;                     // [2][Q_BYTES]
    __half* s_wc = reinterpret_cast<__half*>(s_q + 2 * Cf::Q_BYTES);   // [8 * NCT][C + pad] classifier weights
    const int C = p.C, H = p.H, W = p.W, nch = C / 64, cld = C + WA_CLS_PAD;
    const int tid = threadIdx.x, lane = tid & 31, cw = tid >> 5;
    const int g = lane >> 2, t = lane & 3, mi = lane >> 3;
    const int tiles_x = (W + WA_SW - 1) / WA_SW;
    const int x0 = (blockIdx.x % tiles_x) * WA_SW, y0 = (blockIdx.x / tiles_x) * 4, n = blockIdx.y;
    const size_t plane = (size_t)H * W, img = (size_t)n * plane;

    // cp.async tile fills (16 bytes per request).  Tile position (ky, kx) <-> pixel (y0 - R + ky, x0 - R + kx); positions
    // outside the image are zero-filled: K / V are exactly 0 there (attention zero padding, model/attention.py:199,207).
    auto fill_kv = [&](const __half* src, int buf, int c) {
        const uint32_t base = s_u32(s_kv + (size_t)buf * Cf::KV_BYTES);
        for (int i = tid; i < Cf::KVR * Cf::KVC * 8; i += WA_THREADS) {
            const int pos = i >> 3, ch = i & 7, ky = pos / Cf::KVC, kx = pos - ky * Cf::KVC;
            const int y = y0 - Cf::R + ky, x = x0 - Cf::R + kx;
            const bool ok = y >= 0 && y < H && x >= 0 && x < W;
            const __half* s = src + (img + (ok ? (size_t)y * W + x : 0)) * C + c * 64 + ch * 8;
            cp_async16(base + (uint32_t)(pos * 128 + (((ch ^ kx) & 7) << 4)), s, ok);
        }
    };
    auto fill_q = [&](int buf, int c) {
        const uint32_t base = s_u32(s_q + (size_t)buf * Cf::Q_BYTES);
        for (int i = tid; i < 64 * 8; i += WA_THREADS) {
            const int pos = i >> 3, ch = i & 7, row = pos >> 4, col = pos & 15;
            const int y = y0 + row, x = x0 + col;
            const bool ok = y < H && x < W;
            const __half* s = q.Q + (img + (ok ? (size_t)y * W + x : 0)) * C + c * 64 + ch * 8;
            cp_async16(base + (uint32_t)(pos * 128 + (((ch ^ ((col & 3) | ((row & 1) << 2))) & 7) << 4)), s, ok);
        }
    };
    fill_kv(q.K, 0, 0);
    fill_q(0, 0);
    cp_async_commit();
    if (NCT > 0) {
        for (int i = tid; i < 8 * NCT * C; i += WA_THREADS) {
            const int j = i / C, c = i - j * C;
            s_wc[j * cld + c] = __float2half_rn(j < p.ncls ? clamp_h(__ldg(p.wcls + (size_t)j * C + c)) : 0.f);
        }
    }

    // validity masks of this thread's logits (identical to the march engine's C role)
    constexpr int NT8Q = Cf::NT8Q, MW = (2 * NT8Q + 31) / 32;
    uint32_t mA[MW], mB[MW];
#pragma unroll
    for (int w = 0; w < MW; ++w) mA[w] = mB[w] = 0u;
    {
        const int qy0 = g >> 2, qx0 = g & 3, qy1 = qy0 + 2;
#pragma unroll
        for (int j = 0; j < NT8Q; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int nk = 8 * j + 2 * t + e;
                const int ky = nk / Cf::WN, kx = nk % Cf::WN;
                const bool okx = nk < Cf::NK && (unsigned)(kx - qx0) < (unsigned)K;
                if (okx && (unsigned)(ky - qy0) < (unsigned)K) mA[(2 * j + e) >> 5] |= 1u << ((2 * j + e) & 31);
                if (okx && (unsigned)(ky - qy1) < (unsigned)K) mB[(2 * j + e) >> 5] |= 1u << ((2 * j + e) & 31);
            }
    }
    uint32_t aq[NT8Q], av[Cf::NT16];
#pragma unroll
    for (int j = 0; j < NT8Q; ++j) {
        int nk = 8 * j + (lane & 7);
        nk = nk < Cf::NK ? nk : Cf::NK - 1;
        const int ky = nk / Cf::WN, kx = nk - ky * Cf::WN, col = 4 * cw + kx;
        aq[j] = (uint32_t)ky * Cf::ROWB + (uint32_t)(col * 128 + (((mi ^ col) & 7) << 4));
    }
#pragma unroll
    for (int i = 0; i < Cf::NT16; ++i) {
        int nk = 16 * i + ((mi & 1) << 3) + (lane & 7);
        nk = nk < Cf::NK ? nk : Cf::NK - 1;
        const int ky = nk / Cf::WN, kx = nk - ky * Cf::WN, col = 4 * cw + kx;
        av[i] = (uint32_t)ky * Cf::ROWB + (uint32_t)(col * 128 + ((((mi >> 1) ^ col) & 7) << 4));
    }
    uint32_t qoff[4];
    {
        const int r = ((mi & 1) << 3) + (lane & 7), row = r >> 2, col = 4 * cw + (r & 3);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
            qoff[ks] = (uint32_t)((row * WA_SW + col) * 128 + ((((2 * ks + (mi >> 1)) ^ ((col & 3) | ((row & 1) << 2))) & 7) << 4));
    }

    // ---------------- phase 1: S = sum over channel chunks of Q_c K_c^T (model/attention.py:199) ----------------
    float Sx[NT8Q][4];
#pragma unroll
    for (int j = 0; j < NT8Q; ++j) {
        const int b0 = 2 * j, b1 = 2 * j + 1;
        Sx[j][0] = (mA[b0 >> 5] >> (b0 & 31)) & 1u ? 0.f : -INFINITY;
        Sx[j][1] = (mA[b1 >> 5] >> (b1 & 31)) & 1u ? 0.f : -INFINITY;
        Sx[j][2] = (mB[b0 >> 5] >> (b0 & 31)) & 1u ? 0.f : -INFINITY;
        Sx[j][3] = (mB[b1 >> 5] >> (b1 & 31)) & 1u ? 0.f : -INFINITY;
    }
#pragma unroll 1
    for (int c = 0; c < nch; ++c) {
        const int buf = c & 1;
        if (c + 1 < nch) { fill_kv(q.K, buf ^ 1, c + 1); fill_q(buf ^ 1, c + 1); }
        else fill_kv(q.V, buf ^ 1, 0);                                   // first V tile rides behind the last K tile
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const uint32_t kb = s_u32(s_kv + (size_t)buf * Cf::KV_BYTES), qb = s_u32(s_q + (size_t)buf * Cf::Q_BYTES);
        uint32_t qa[4][4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) ldsm_x4(qa[ks], qb + qoff[ks]);
#pragma unroll
        for (int j = 0; j < NT8Q; ++j) {
            uint32_t b0[4], b1[4];
            ldsm_x4(b0, kb + aq[j]);
            ldsm_x4(b1, (kb + aq[j]) ^ 64u);
            mma16816(Sx[j], qa[0], b0[0], b0[1]);
            mma16816(Sx[j], qa[1], b0[2], b0[3]);
            mma16816(Sx[j], qa[2], b1[0], b1[1]);
            mma16816(Sx[j], qa[3], b1[2], b1[3]);
        }
        __syncthreads();                                                 // everyone is done with buffer `buf` before it is refilled
    }
    // ---------------- phase 2: softmax over the k*k window (model/attention.py:203) ----------------
    float mx0, mx1;
    {
        float a[NT8Q], b[NT8Q];
#pragma unroll
        for (int j = 0; j < NT8Q; ++j) { a[j] = fmaxf(Sx[j][0], Sx[j][1]); b[j] = fmaxf(Sx[j][2], Sx[j][3]); }
#pragma unroll
        for (int w = 1; w < NT8Q; w <<= 1)
#pragma unroll
            for (int j = 0; j + w < NT8Q; j += 2 * w) { a[j] = fmaxf(a[j], a[j + w]); b[j] = fmaxf(b[j], b[j + w]); }
        mx0 = a[0]; mx1 = b[0];
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    constexpr float LOG2E = 1.4426950408889634f;
    const float o0 = mx0 * LOG2E, o1 = mx1 * LOG2E;
    uint32_t pa[Cf::NT16][4];
#pragma unroll
    for (int i = 0; i < Cf::NT16; ++i)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int j = 2 * i + h;
            if (j < NT8Q) {
                pa[i][2 * h] = pack_h2(wa_ex2(fmaf(Sx[j][0], LOG2E, -o0)), wa_ex2(fmaf(Sx[j][1], LOG2E, -o0)));
                pa[i][2 * h + 1] = pack_h2(wa_ex2(fmaf(Sx[j][2], LOG2E, -o1)), wa_ex2(fmaf(Sx[j][3], LOG2E, -o1)));
            } else {
                pa[i][2 * h] = 0u; pa[i][2 * h + 1] = 0u;
            }
        }
    // row sums of the f16-rounded P (ones-column MMA), once
    float inv0, inv1;
    {
        float Ssum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < Cf::NT16; ++i) mma16816(Ssum, pa[i], 0x3C003C00u, 0x3C003C00u);
        inv0 = wa_rcp(Ssum[0]); inv1 = wa_rcp(Ssum[2]);
    }
    // ---------------- phase 3: per chunk O_c = P V_c (model/attention.py:207), residual, fused p, classifier ----------------
    const int pxA = x0 + 4 * cw + (g & 3), pyA = y0 + (g >> 2);
    const bool okA = pyA < H && pxA < W, okB = pyA + 2 < H && pxA < W;
    const size_t offA = (size_t)pyA * W + pxA, offB = offA + 2 * (size_t)W;
    constexpr int NCTA = NCT > 0 ? NCT : 1;
    float Lg[NCTA][4];
#pragma unroll
    for (int nt = 0; nt < NCTA; ++nt) Lg[nt][0] = Lg[nt][1] = Lg[nt][2] = Lg[nt][3] = 0.f;
#pragma unroll 1
    for (int c = 0; c < nch; ++c) {
        const int buf = (nch + c) & 1;                                   // V tile c lives in the buffer after the last K tile
        if (c + 1 < nch) fill_kv(q.V, buf ^ 1, c + 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const uint32_t vb = s_u32(s_kv + (size_t)buf * Cf::KV_BYTES);
        float O[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) O[j][0] = O[j][1] = O[j][2] = O[j][3] = 0.f;
        {
            uint32_t v[2][4][4];
#pragma unroll
            for (int cp = 0; cp < 4; ++cp) ldsm_x4_t(v[0][cp], (vb + av[0]) ^ (uint32_t)(cp << 5));
#pragma unroll
            for (int i = 0; i < Cf::NT16; ++i) {
                if (i + 1 < Cf::NT16) {
#pragma unroll
                    for (int cp = 0; cp < 4; ++cp) ldsm_x4_t(v[(i + 1) & 1][cp], (vb + av[i + 1]) ^ (uint32_t)(cp << 5));
                }
#pragma unroll
                for (int cp = 0; cp < 4; ++cp) {
                    mma16816(O[2 * cp], pa[i], v[i & 1][cp][0], v[i & 1][cp][1]);
                    mma16816(O[2 * cp + 1], pa[i], v[i & 1][cp][2], v[i & 1][cp][3]);
                }
            }
        }
        __syncthreads();                                                 // buffer `buf` may be refilled by the next iteration
        // fused = lr_up + O / sum (model/attention.py:210): thread (g,t) holds channels 64c + 8j + 2t, +1 of pixels A and B
        const float* ra = q.R + (img + offA) * C + c * 64 + 2 * t;
        const float* rb = q.R + (img + offB) * C + c * 64 + 2 * t;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float2 r0 = okA ? __ldg(reinterpret_cast<const float2*>(ra + 8 * j)) : make_float2(0.f, 0.f);
            const float2 r1 = okB ? __ldg(reinterpret_cast<const float2*>(rb + 8 * j)) : make_float2(0.f, 0.f);
            O[j][0] = fmaf(O[j][0], inv0, r0.x); O[j][1] = fmaf(O[j][1], inv0, r0.y);
            O[j][2] = fmaf(O[j][2], inv1, r1.x); O[j][3] = fmaf(O[j][3], inv1, r1.y);
        }
        if (p.out_p) {
            float* op = p.out_p + ((size_t)n * C + c * 64 + 2 * t) * plane;
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    float* o = op + (size_t)(8 * j + e) * plane;
                    if (okA) o[offA] = O[j][e];
                    if (okB) o[offB] = O[j][2 + e];
                }
        }
        if (NCT > 0) {
            uint32_t fa[4][4];
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                fa[ks][0] = pack_h2_sat(O[2 * ks][0], O[2 * ks][1]);
                fa[ks][1] = pack_h2_sat(O[2 * ks][2], O[2 * ks][3]);
                fa[ks][2] = pack_h2_sat(O[2 * ks + 1][0], O[2 * ks + 1][1]);
                fa[ks][3] = pack_h2_sat(O[2 * ks + 1][2], O[2 * ks + 1][3]);
            }
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
#pragma unroll
                for (int nt = 0; nt < NCT; ++nt) {
                    const __half* wp = s_wc + (8 * nt + g) * cld + c * 64 + 16 * ks + 2 * t;
                    mma16816(Lg[nt], fa[ks], *reinterpret_cast<const uint32_t*>(wp), *reinterpret_cast<const uint32_t*>(wp + 8));
                }
        }
    }
    if (NCT == 0) return;
    // ---------------- classifier bias, argmax, log-softmax (model/pspnet.py:226-229, evaluation.py:204) ----------------
#pragma unroll
    for (int nt = 0; nt < NCT; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int cls = 8 * nt + 2 * t + e;
            const float b = cls < p.ncls ? (p.bcls ? __ldg(p.bcls + cls) : 0.f) : -INFINITY;      // padding classes: -inf
            Lg[nt][e] += b; Lg[nt][2 + e] += b;
        }
    float lmax0 = -INFINITY, lmax1 = -INFINITY;
    int am0 = 0, am1 = 0;
#pragma unroll
    for (int nt = 0; nt < NCT; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int cls = 8 * nt + 2 * t + e;
            if (Lg[nt][e] > lmax0) { lmax0 = Lg[nt][e]; am0 = cls; }
            if (Lg[nt][2 + e] > lmax1) { lmax1 = Lg[nt][2 + e]; am1 = cls; }
        }
#pragma unroll
    for (int d = 1; d <= 2; d <<= 1) {
        const float v0 = __shfl_xor_sync(0xffffffffu, lmax0, d), v1 = __shfl_xor_sync(0xffffffffu, lmax1, d);
        const int i0 = __shfl_xor_sync(0xffffffffu, am0, d), i1 = __shfl_xor_sync(0xffffffffu, am1, d);
        if (v0 > lmax0 || (v0 == lmax0 && i0 < am0)) { lmax0 = v0; am0 = i0; }
        if (v1 > lmax1 || (v1 == lmax1 && i1 < am1)) { lmax1 = v1; am1 = i1; }
    }
    float lse0 = 0.f, lse1 = 0.f;
    if (p.log_softmax) {
        const float q0 = lmax0 * LOG2E, q1 = lmax1 * LOG2E;
#pragma unroll
        for (int nt = 0; nt < NCT; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                lse0 += wa_ex2(fmaf(Lg[nt][e], LOG2E, -q0));
                lse1 += wa_ex2(fmaf(Lg[nt][2 + e], LOG2E, -q1));
            }
        lse0 += __shfl_xor_sync(0xffffffffu, lse0, 1); lse0 += __shfl_xor_sync(0xffffffffu, lse0, 2);
        lse1 += __shfl_xor_sync(0xffffffffu, lse1, 1); lse1 += __shfl_xor_sync(0xffffffffu, lse1, 2);
        lse0 = __logf(lse0) + lmax0; lse1 = __logf(lse1) + lmax1;
    }
    if (p.out_logits) {
        float* ol = p.out_logits + ((size_t)n * p.ncls + 2 * t) * plane;
#pragma unroll
        for (int nt = 0; nt < NCT; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e)
                if (8 * nt + 2 * t + e < p.ncls) {
                    float* o = ol + (size_t)(8 * nt + e) * plane;
                    if (okA) o[offA] = Lg[nt][e] - lse0;
                    if (okB) o[offB] = Lg[nt][2 + e] - lse1;
                }
    }
    if (p.out_argmax && t == 0) {
        uint8_t* oa = p.out_argmax + img;
        if (okA) oa[offA] = (uint8_t)am0;
        if (okB) oa[offB] = (uint8_t)am1;
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
size_t creff_wide_workspace_bytes(int N, int C, int H, int W) {
    const size_t e = (size_t)N * H * W * C, px = (size_t)N * H * W;
    return e * (3 * sizeof(__half) + sizeof(float)) + px * 2 * (sizeof(float4) + sizeof(uint4)) + 1024;
}

bool creff_wide_supported(const arseg_creff_args* a) {
    return a->C > MC && a->C % MC == 0 && a->C <= 1024 && a->hr_layout == ARSEG_NHWC && a->lr_layout == ARSEG_NHWC &&
           (a->k == 3 || a->k == 5 || a->k == 7 || a->k == 9) && (!a->wcls || a->ncls <= 32) && a->H >= 2 && a->W >= 2 && a->h >= 2 && a->w >= 2 &&
           ((size_t)a->H * a->W * a->C < (1u << 30)) && ((size_t)a->h * a->w * a->C < (1u << 30));
}

template <int K, int NCT>
static int wide_attn_launch(const WideParams& q, cudaStream_t st) {
    using Cf = WACfg<K>;
    const size_t smem = 2 * Cf::KV_BYTES + 2 * Cf::Q_BYTES + (size_t)8 * NCT * (q.b.C + WA_CLS_PAD) * 2;
    auto kern = creff_wide_attn_kernel<K, NCT>;
    ARSEG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles = ceil_div(q.b.W, WA_SW) * ceil_div(q.b.H, 4);
    ARSEG_REQUIRE(q.b.N <= 65535, "creff_wide: N too large");
    kern<<<dim3((unsigned)tiles, (unsigned)q.b.N), WA_THREADS, smem, st>>>(q);
    ARSEG_CHECK_LAUNCH("creff_wide_attn");
    return ARSEG_OK;
}

template <int K>
static int wide_attn_launch_k(const WideParams& q, cudaStream_t st) {
    if (!q.b.wcls) return wide_attn_launch<K, 0>(q, st);
    if (q.b.ncls <= 16) return wide_attn_launch<K, 2>(q, st);
    return wide_attn_launch<K, 4>(q, st);
}

int creff_wide_launch(const arseg_creff_args* a, void* ws, size_t ws_bytes, cudaStream_t st) {
    ARSEG_REQUIRE(ws && ws_bytes >= creff_wide_workspace_bytes(a->N, a->C, a->H, a->W),
                  "creff_wide: workspace of %zu bytes needed (arseg_creff_workspace_bytes)", creff_wide_workspace_bytes(a->N, a->C, a->H, a->W));
    WideParams q;
    CreffMmaParams& p = q.b;
    p.hr = reinterpret_cast<const float*>(a->hr); p.hr_shared = a->hr_shared; p.flow = a->flow; p.flow_dtype = a->flow_dtype; p.Hm = a->Hm; p.Wm = a->Wm;
    p.lr = a->lr; p.h = a->h; p.w = a->w;
    p.wq = a->wq; p.bq = a->bq; p.wk = a->wk; p.bk = a->bk; p.wv = a->wv; p.bv = a->bv; p.wcls = a->wcls; p.bcls = a->bcls;
    p.ncls = a->ncls; p.log_softmax = a->log_softmax; p.out_p = a->out_p; p.out_logits = a->out_logits;
    p.out_argmax = a->out_argmax; p.N = a->N; p.C = a->C; p.H = a->H; p.W = a->W;
    p.tiles_x = p.tiles_y = p.seg_rows = p.ncols = p.nseg = p.dbg = 0;
    const size_t e = (size_t)a->N * a->H * a->W * a->C;
    uint8_t* w8 = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
    q.R = reinterpret_cast<float*>(w8);
    q.Q = reinterpret_cast<__half*>(w8 + e * 4);
    q.K = q.Q + e;
    q.V = q.K + e;
    q.rec_w = reinterpret_cast<float4*>(q.V + e);                 // e * 10 bytes so far: 16-byte aligned (e is a multiple of 64)
    q.rec_a = reinterpret_cast<uint4*>(q.rec_w + (size_t)a->N * a->H * a->W * 2);
    q.lr_dtype = a->lr_dtype;
    ARSEG_REQUIRE(a->N <= 65535 && a->C / 64 <= 65535, "creff_wide: N / C too large");
    const dim3 grid((unsigned)(ceil_div(a->W, WP_TW) * ceil_div(a->H, WP_TH)), (unsigned)(a->C / 64), (unsigned)a->N);
    {
        const long long nrec = (long long)a->N * a->H * a->W * 2;
        const unsigned rb = (unsigned)((nrec + 255) / 256);
        if (a->lr_dtype == ARSEG_F32) creff_wide_rec_kernel<float><<<rb, 256, 0, st>>>(q);
        else creff_wide_rec_kernel<__half><<<rb, 256, 0, st>>>(q);        // fp16 and bf16 LR features: same element size
        ARSEG_CHECK_LAUNCH("creff_wide_rec");
    }
    if (a->lr_dtype == ARSEG_F32) {
        ARSEG_CUDA(cudaFuncSetAttribute(creff_wide_prep_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WP_SMEM));
        creff_wide_prep_kernel<float><<<grid, WP_THREADS, WP_SMEM, st>>>(q);
    } else if (a->lr_dtype == ARSEG_F16) {
        ARSEG_CUDA(cudaFuncSetAttribute(creff_wide_prep_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WP_SMEM));
        creff_wide_prep_kernel<__half><<<grid, WP_THREADS, WP_SMEM, st>>>(q);
    } else {
        ARSEG_CUDA(cudaFuncSetAttribute(creff_wide_prep_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WP_SMEM));
        creff_wide_prep_kernel<__nv_bfloat16><<<grid, WP_THREADS, WP_SMEM, st>>>(q);
    }
    ARSEG_CHECK_LAUNCH("creff_wide_prep");
    switch (a->k) {
        case 3: return wide_attn_launch_k<3>(q, st);
        case 5: return wide_attn_launch_k<5>(q, st);
        case 7: return wide_attn_launch_k<7>(q, st);
        case 9: return wide_attn_launch_k<9>(q, st);
        default: ARSEG_UNSUPPORTED("creff_wide: window k=%d", a->k);
    }
}

}  // namespace arseg
