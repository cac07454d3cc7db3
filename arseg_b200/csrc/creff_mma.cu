// creff_mma.cu -- fused MV-warp + CReFF + classifier, tensor-core engine (ARSEG_CREFF_MMA_F16), sm_100a.
//
// Same contract as creff.cu (the exact fp32 SIMT engine; see the reference citations there), for C = 64
// (CamVid PSPNet-18, SURVEY.md 8a rows 2,3,10,11,12,13) with NHWC operands.  What changes is the arithmetic
// engine of the two window contractions (model/attention.py:199 f_similar and :207 f_weighting):
//
//   * a CTA owns a 16x16 pixel tile; each of its 16 warps owns one 4x4-pixel block = one M=16 MMA tile.
//     The k x k windows of a 4x4 block cover a (k+3)^2 key patch, so Q.K^T of the block is a dense
//     [16 x 64] x [64 x (k+3)^2] product of which (k/(k+3))^2 (49 % at k=7) is used -- a genuine dense
//     contraction with 2.04x redundancy, issued as warp-level mma.sync.m16n8k16 (f16 operands, fp32
//     accumulate).  tcgen05's minimum M=64/128 tile would need a (k+7)x(k+15) key patch per tile (6.3x
//     redundancy) and a per-lane band extraction from TMEM; the M=16 register-fragment path keeps the
//     logits, the softmax and P in registers.
//   * f16 (not bf16) operands: 11-bit significands = TF32's, so the error is TF32-class (~5e-4 relative);
//     values are clamped to +-60000 on conversion.  Everything else -- MV arithmetic (f64), bilinear warp,
//     bilinear up-sampling, the three depthwise 3x3 convolutions, softmax, residual -- is fp32 SIMT.
//   * production pipeline: the warped-hr tile (and the lr_up tile) is produced in 4-row strips into a
//     10-row shared-memory ring (fp32); tap loads of strip s+1 are in flight in registers while the
//     depthwise convolutions of strip s run.  K, V (and Q) live in shared memory as f16 rows of 64
//     channels (128 B) with a 16-byte-chunk XOR swizzle (chunk ^= row & 7) so ldmatrix is conflict free.
//
// HBM traffic per launch: hr (NHWC fp32, read through L2; frames of a GOP are scheduled tile-major so
// the shared keyframe tile stays L2 resident), lr, MV field, logits (+ argmax, + fused p if asked).
#include "creff_mma_common.cuh"
#include <cstdlib>

namespace arseg {

constexpr int MT = 16;            // tile edge (pixels)
constexpr int MTHREADS = 512;     // 16 warps
constexpr int MSR = 4;            // strip rows
constexpr int MRING = 10;         // ring rows
constexpr int MCLS_LD = 72;       // f16 row stride of the classifier weights in smem (bank-conflict pad)
constexpr int MOUT_LD = 260;      // fp32 row stride of the staged logits planes

template <int K> struct MCfg {
    static constexpr int R = K / 2;
    static constexpr int KR = MT + K - 1, KC = MT + K - 1;     // K / V tile (positions)
    static constexpr int WR = KR + 2, WC = KC + 2;             // warped-hr tile (+ depthwise halo)
    static constexpr int LRW = MT + 2;                         // lr_up tile edge
    static constexpr int WN = K + 3, NK = WN * WN;             // key patch of a 4x4 block
    static constexpr int NT16 = (NK + 15) / 16, NT8 = 2 * NT16;
    static constexpr int PPW = (MSR * WC + 15) / 16;           // gather positions per warp per strip
    static constexpr int RC_KV = (KC + 7) / 8;                 // depthwise column run per warp
    static constexpr size_t KV_BYTES = (size_t)KR * KC * 128;
    static constexpr size_t RING_BYTES = (size_t)MRING * WC * MC * 4;
    static constexpr size_t POS_N = (size_t)MSR * WC;          // per buffer
    static constexpr size_t SMEM = 2 * KV_BYTES + RING_BYTES + 2 * POS_N * 20 + 32 * MCLS_LD * 2 + 32 * 4 + 256;
};

// ---------------------------------------------------------------------------------------------
// Producer (warp specialised).  A (OR+2)x(OC+2) fp32 tile is gathered bilinearly from an NHWC source in
// 4-row strips into the smem ring by warps 0-7, and consumed by warps 8-15, which apply NOUT depthwise 3x3
// convolutions (+bias) and write swizzled f16 rows.  The two groups hand strips over through named barriers
// (FULL[s&1]: strip s is in the ring; EMPTY[s&1]: strip s has been consumed, its ring rows may be reused), so
// the gather of strip s+1/s+2 overlaps the convolutions of strip s.
//   MODE 0: source = lr (lr_up tile), NOUT = 1 (Q);   MODE 1: source = hr (warped tile), NOUT = 2 (K, V)
// ---------------------------------------------------------------------------------------------
template <int K, int MODE> struct PCfg {
    using Cf = MCfg<K>;
    static constexpr int OR = MODE ? Cf::KR : MT, OC = MODE ? Cf::KC : MT;   // output tile
    static constexpr int IR = OR + 2, IC = OC + 2;                           // input tile
    static constexpr int ORG = MODE ? Cf::R + 1 : 1;                         // input tile origin = (y0-ORG, x0-ORG)
    static constexpr int NSTEP = (OR + MSR - 1) / MSR;
    static constexpr int PPH = (MSR * IC + 15) / 16;                         // positions per half-warp per step
    static constexpr int JA = (PPH + 1) / 2, JB = PPH - JA;
    static constexpr int RSTRIDE = Cf::WC * MC;                              // ring row stride (floats)
    static constexpr int CHALF = (OC + 1) / 2;
};

template <int K, int MODE, typename TSRC>
__device__ __forceinline__ void gather_role(const CreffMmaParams& p, const TSRC* __restrict__ src, int srcW, int n, int y0, int x0,
                                            float* ring, float4* posw, int* posi) {
    using Cf = MCfg<K>;
    using Pc = PCfg<K, MODE>;
    constexpr int IR = Pc::IR, IC = Pc::IC, NSTEP = Pc::NSTEP;
    const int gt = threadIdx.x, lane = gt & 31, hw = gt >> 4, cl = lane & 15;   // gt in [0,256)
    const float lsh = resize_scale(p.h, p.H, ARSEG_RESIZE_BILINEAR_AC), lsw = resize_scale(p.w, p.W, ARSEG_RESIZE_BILINEAR_AC);
    // gather step g (g = -1: rows 0,1; g >= 0: rows 4g+2 .. 4g+5); position records of step g in buffer g&1
    auto step_rows = [&](int g, int& r0, int& nr) {
        if (g < 0) { r0 = 0; nr = 2; } else { r0 = MSR * g + 2; nr = min(MSR, IR - r0); }
    };
    auto compute_pos = [&](int g) {
        int r0, nr; step_rows(g, r0, nr);
        if (gt < nr * IC) {
            const int rr = r0 + gt / IC, cc = gt % IC;
            const int fy = y0 - Pc::ORG + rr, fx = x0 - Pc::ORG + cc;
            PosRec r = MODE ? pos_hr(p, n, fy, fx) : pos_lr(p, lsh, lsw, fy, fx);
            posw[(g & 1) * Cf::POS_N + gt] = r.w;
            posi[(g & 1) * Cf::POS_N + gt] = r.info;
        }
    };
    float4 tap[Pc::JA][4];
    auto issue = [&](int buf, int npos, int j0, int nj) {
#pragma unroll
        for (int j = 0; j < Pc::JA; ++j) {
            if (j < nj) {
                const int i = hw + 16 * (j0 + j);
                int info = -1;
                if (i < npos) info = posi[buf * Cf::POS_N + i];
                if (info >= 0) {
                    const int dx = (info >> 1) & 1, dy = info & 1;
                    const TSRC* s = src + (size_t)(info >> 2) * MC + 4 * cl;
                    const TSRC* s2 = s + (size_t)dy * srcW * MC;
                    tap[j][0] = ld4(s);
                    tap[j][1] = ld4(s + dx * MC);
                    tap[j][2] = ld4(s2);
                    tap[j][3] = ld4(s2 + dx * MC);
                } else {
                    tap[j][0] = tap[j][1] = tap[j][2] = tap[j][3] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
    };
    auto commit = [&](int buf, int r0, int npos, int j0, int nj) {
#pragma unroll
        for (int j = 0; j < Pc::JA; ++j) {
            if (j < nj) {
                const int i = hw + 16 * (j0 + j);
                if (i < npos) {
                    const float4 w = posw[buf * Cf::POS_N + i];
                    float4 v;
                    v.x = tap[j][0].x * w.x + tap[j][1].x * w.y + tap[j][2].x * w.z + tap[j][3].x * w.w;
                    v.y = tap[j][0].y * w.x + tap[j][1].y * w.y + tap[j][2].y * w.z + tap[j][3].y * w.w;
                    v.z = tap[j][0].z * w.x + tap[j][1].z * w.y + tap[j][2].z * w.z + tap[j][3].z * w.w;
                    v.w = tap[j][0].w * w.x + tap[j][1].w * w.y + tap[j][2].w * w.z + tap[j][3].w * w.w;
                    const int rr = r0 + i / IC, cc = i % IC;
                    *reinterpret_cast<float4*>(ring + (rr % MRING) * Pc::RSTRIDE + cc * MC + 4 * cl) = v;
                }
            }
        }
    };
    compute_pos(-1);
    nbar_sync(BAR_GATHER, 256);
#pragma unroll 1
    for (int g = -1; g < NSTEP; ++g) {
        const int buf = g & 1;
        if (g >= 2) nbar_sync(BAR_EMPTY + buf, MTHREADS);     // strip g-2 consumed: its ring rows are free
        int r0, nr; step_rows(g, r0, nr);
        const int npos = nr * IC;
        issue(buf, npos, 0, Pc::JA);
        if (g + 1 < NSTEP) compute_pos(g + 1);                // f64 MV arithmetic overlaps the loads in flight
        commit(buf, r0, npos, 0, Pc::JA);
        if (Pc::JB > 0) {
            issue(buf, npos, Pc::JA, Pc::JB);
            commit(buf, r0, npos, Pc::JA, Pc::JB);
        }
        nbar_sync(BAR_GATHER, 256);                           // position records of step g+1 visible to the group
        if (g >= 0) nbar_arrive(BAR_FULL + buf, MTHREADS);    // strip g (ring rows <= 4g+5) is complete
    }
}

template <int K, int MODE>
__device__ __forceinline__ void dw_role(const CreffMmaParams& p, int y0, int x0, const float* ring, uint8_t* out1, uint8_t* out2,
                                        const float* __restrict__ w1g, const float* __restrict__ b1g,
                                        const float* __restrict__ w2g, const float* __restrict__ b2g) {
    using Cf = MCfg<K>;
    using Pc = PCfg<K, MODE>;
    constexpr int OR = Pc::OR, OC = Pc::OC, NSTEP = Pc::NSTEP, NOUT = MODE ? 2 : 1;
    const int lane = threadIdx.x & 31, dwarp = (threadIdx.x >> 5) - 8;
    // this lane's channel pair (2*lane, 2*lane+1)
    float2 w1[9], w2[9], b1, b2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        w1[t] = make_float2(__ldg(w1g + (2 * lane) * 9 + t), __ldg(w1g + (2 * lane + 1) * 9 + t));
        w2[t] = NOUT == 2 ? make_float2(__ldg(w2g + (2 * lane) * 9 + t), __ldg(w2g + (2 * lane + 1) * 9 + t)) : make_float2(0.f, 0.f);
    }
    b1 = make_float2(__ldg(b1g + 2 * lane), __ldg(b1g + 2 * lane + 1));
    if (NOUT == 2) b2 = make_float2(__ldg(b2g + 2 * lane), __ldg(b2g + 2 * lane + 1));
    const int rsub = dwarp >> 1;
    const int c_lo = (dwarp & 1) * Pc::CHALF, c_hi = min(OC, c_lo + Pc::CHALF);
    const uint32_t lane_off = (uint32_t)((lane & 3) * 4);
    const int lane_chunk = lane >> 2;
#pragma unroll 1
    for (int s = 0; s < NSTEP; ++s) {
        nbar_sync(BAR_FULL + (s & 1), MTHREADS);
        const int orow = MSR * s + rsub;
        if (orow < OR) {
            const float* rp0 = ring + (orow % MRING) * Pc::RSTRIDE + 2 * lane;
            const float* rp1 = ring + ((orow + 1) % MRING) * Pc::RSTRIDE + 2 * lane;
            const float* rp2 = ring + ((orow + 2) % MRING) * Pc::RSTRIDE + 2 * lane;
            bool row_ok = true;
            if (MODE) { const int fy = y0 - Cf::R + orow; row_ok = fy >= 0 && fy < p.H; }
            float2 win[3][3];   // [input row][slot]; slot (x - c_lo + d) % 3 holds input column x + d
            win[0][0] = *reinterpret_cast<const float2*>(rp0 + c_lo * MC);
            win[1][0] = *reinterpret_cast<const float2*>(rp1 + c_lo * MC);
            win[2][0] = *reinterpret_cast<const float2*>(rp2 + c_lo * MC);
            win[0][1] = *reinterpret_cast<const float2*>(rp0 + (c_lo + 1) * MC);
            win[1][1] = *reinterpret_cast<const float2*>(rp1 + (c_lo + 1) * MC);
            win[2][1] = *reinterpret_cast<const float2*>(rp2 + (c_lo + 1) * MC);
            for (int xb = c_lo; xb < c_hi; xb += 3) {
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    const int x = xb + u;
                    if (x < c_hi) {
                        const int sa = u, sb = (u + 1) % 3, sc = (u + 2) % 3;     // slots of columns x, x+1, x+2
                        win[0][sc] = *reinterpret_cast<const float2*>(rp0 + (x + 2) * MC);
                        win[1][sc] = *reinterpret_cast<const float2*>(rp1 + (x + 2) * MC);
                        win[2][sc] = *reinterpret_cast<const float2*>(rp2 + (x + 2) * MC);
                        float2 a1 = b1;
                        a1 = __ffma2_rn(w1[0], win[0][sa], a1); a1 = __ffma2_rn(w1[1], win[0][sb], a1); a1 = __ffma2_rn(w1[2], win[0][sc], a1);
                        a1 = __ffma2_rn(w1[3], win[1][sa], a1); a1 = __ffma2_rn(w1[4], win[1][sb], a1); a1 = __ffma2_rn(w1[5], win[1][sc], a1);
                        a1 = __ffma2_rn(w1[6], win[2][sa], a1); a1 = __ffma2_rn(w1[7], win[2][sb], a1); a1 = __ffma2_rn(w1[8], win[2][sc], a1);
                        float2 a2 = b2;
                        if (NOUT == 2) {
                            a2 = __ffma2_rn(w2[0], win[0][sa], a2); a2 = __ffma2_rn(w2[1], win[0][sb], a2); a2 = __ffma2_rn(w2[2], win[0][sc], a2);
                            a2 = __ffma2_rn(w2[3], win[1][sa], a2); a2 = __ffma2_rn(w2[4], win[1][sb], a2); a2 = __ffma2_rn(w2[5], win[1][sc], a2);
                            a2 = __ffma2_rn(w2[6], win[2][sa], a2); a2 = __ffma2_rn(w2[7], win[2][sb], a2); a2 = __ffma2_rn(w2[8], win[2][sc], a2);
                        }
                        if (MODE) {   // K / V are exactly 0 outside the image (attention zero padding)
                            const int fx = x0 - Cf::R + x;
                            if (!(row_ok && fx >= 0 && fx < p.W)) { a1 = make_float2(0.f, 0.f); a2 = make_float2(0.f, 0.f); }
                        }
                        const int pos = orow * OC + x;
                        const uint32_t off = (uint32_t)(pos * 128 + (((lane_chunk ^ pos) & 7) << 4)) + lane_off;
                        *reinterpret_cast<uint32_t*>(out1 + off) = pack_h2_sat(a1.x, a1.y);
                        if (NOUT == 2) *reinterpret_cast<uint32_t*>(out2 + off) = pack_h2_sat(a2.x, a2.y);
                    }
                }
            }
        }
        if (s + 2 < NSTEP) nbar_arrive(BAR_EMPTY + (s & 1), MTHREADS);
    }
}

template <int K, int MODE, typename TSRC>
__device__ __forceinline__ void produce(const CreffMmaParams& p, const TSRC* __restrict__ src, int srcW, int n, int y0, int x0,
                                        float* ring, float4* posw, int* posi, uint8_t* out1, uint8_t* out2,
                                        const float* __restrict__ w1g, const float* __restrict__ b1g,
                                        const float* __restrict__ w2g, const float* __restrict__ b2g) {
    if (threadIdx.x < 256) gather_role<K, MODE, TSRC>(p, src, srcW, n, y0, x0, ring, posw, posi);
    else dw_role<K, MODE>(p, y0, x0, ring, out1, out2, w1g, b1g, w2g, b2g);
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
template <int K, typename TLR>
__global__ void __launch_bounds__(MTHREADS, 1) creff_mma_kernel(CreffMmaParams p) {
    using Cf = MCfg<K>;
    extern __shared__ __align__(1024) uint8_t msm[];
    uint8_t* sK = msm;                                   // also holds the Q tile during the Q phase
    uint8_t* sV = sK + Cf::KV_BYTES;
    float* ring = reinterpret_cast<float*>(sV + Cf::KV_BYTES);
    float4* posw = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(ring) + Cf::RING_BYTES);
    int* posi = reinterpret_cast<int*>(posw + 2 * Cf::POS_N);
    __half* s_wc = reinterpret_cast<__half*>(posi + 2 * Cf::POS_N);     // [32][MCLS_LD]
    float* s_bc = reinterpret_cast<float*>(s_wc + 32 * MCLS_LD);        // [32]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // frame index fastest: the N frames of a GOP visit the same keyframe tile back to back (L2 reuse)
    int b = blockIdx.x;
    const int n = b % p.N; b /= p.N;
    const int x0 = (b % p.tiles_x) * MT, y0 = (b / p.tiles_x) * MT;
    const int H = p.H, W = p.W;
    const bool do_cls = p.wcls != nullptr;
    const int nct = do_cls ? (p.ncls + 7) >> 3 : 0;     // classifier n-tiles (<= 4)

    if (do_cls) {
        for (int i = tid; i < 32 * MC; i += MTHREADS) {
            const int j = i / MC, c = i % MC;
            s_wc[j * MCLS_LD + c] = __float2half_rn(j < p.ncls ? clamp_h(__ldg(p.wcls + (size_t)j * MC + c)) : 0.f);
        }
        if (tid < 32) s_bc[tid] = (tid < p.ncls && p.bcls) ? __ldg(p.bcls + tid) : 0.f;
    }

    const TLR* lr = reinterpret_cast<const TLR*>(p.lr) + (size_t)n * p.h * p.w * MC;
    const float* hr = p.hr + (p.hr_shared ? 0 : (size_t)n * H * W * MC);

    // ---------------- Q phase: lr_up tile -> Q = dw3x3_q(lr_up) + bq (model/attention.py:191,197) ----------------
    produce<K, 0, TLR>(p, lr, p.w, n, y0, x0, ring, posw, posi, sK, nullptr, p.wq, p.bq, nullptr, nullptr);

    const int wy = warp >> 2, wx = warp & 3;            // this warp's 4x4-pixel block
    const int g = lane >> 2, t = lane & 3;
    uint32_t qa[4][4];
    {
        const int mi = lane >> 3, r = ((mi & 1) << 3) + (lane & 7);
        const int px = (4 * wy + (r >> 2)) * MT + 4 * wx + (r & 3);
        const uint32_t base = s_u32(sK);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) ldsm_x4(qa[ks], base + swz_chunk(px, 2 * ks + (mi >> 1)));
    }
    __syncthreads();   // Q fragments are in registers: the K region may be overwritten

    // ---------------- K/V phase: warped hr tile -> K, V (evaluation.py:61-87, model/attention.py:194-196) ----------
    produce<K, 1, float>(p, hr, W, n, y0, x0, ring, posw, posi, sK, sV, p.wk, p.bk, p.wv, p.bv);

    // ---------------- attention: S = Q K^T (model/attention.py:199) ----------------
    float S[Cf::NT8][4];
#pragma unroll
    for (int j = 0; j < Cf::NT8; ++j) S[j][0] = S[j][1] = S[j][2] = S[j][3] = 0.f;
    {
        const uint32_t kb = s_u32(sK);
        const int mi = lane >> 3;
#pragma unroll
        for (int j = 0; j < Cf::NT8; ++j) {
            int nk = 8 * j + (lane & 7);
            nk = nk < Cf::NK ? nk : Cf::NK - 1;
            const int pos = (4 * wy + nk / Cf::WN) * Cf::KC + 4 * wx + nk % Cf::WN;
            uint32_t b0[4], b1[4];
            ldsm_x4(b0, kb + swz_chunk(pos, mi));
            ldsm_x4(b1, kb + swz_chunk(pos, 4 + mi));
            mma16816(S[j], qa[0], b0[0], b0[1]);
            mma16816(S[j], qa[1], b0[2], b0[3]);
            mma16816(S[j], qa[2], b1[0], b1[1]);
            mma16816(S[j], qa[3], b1[2], b1[3]);
        }
    }
    // ---------------- softmax over the k*k window of every query (model/attention.py:203) ----------------
    // thread (g,t) holds rows g, g+8 and keys 8j+2t, 8j+2t+1; validity bit 2j+e
    uint64_t m0 = 0, m1 = 0;
    {
        const int qy0 = g >> 2, qx0 = g & 3, qy1 = qy0 + 2;
#pragma unroll
        for (int j = 0; j < Cf::NT8; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int nk = 8 * j + 2 * t + e;
                const int ky = nk / Cf::WN, kx = nk % Cf::WN;
                const bool okx = nk < Cf::NK && (unsigned)(kx - qx0) < (unsigned)K;
                if (okx && (unsigned)(ky - qy0) < (unsigned)K) m0 |= 1ull << (2 * j + e);
                if (okx && (unsigned)(ky - qy1) < (unsigned)K) m1 |= 1ull << (2 * j + e);
            }
    }
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < Cf::NT8; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            if ((m0 >> (2 * j + e)) & 1) mx0 = fmaxf(mx0, S[j][e]);
            if ((m1 >> (2 * j + e)) & 1) mx1 = fmaxf(mx1, S[j][2 + e]);
        }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    constexpr float LOG2E = 1.4426950408889634f;
    const float o0 = mx0 * LOG2E, o1 = mx1 * LOG2E;
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int j = 0; j < Cf::NT8; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const float p0 = ((m0 >> (2 * j + e)) & 1) ? exp2f(fmaf(S[j][e], LOG2E, -o0)) : 0.f;
            const float p1 = ((m1 >> (2 * j + e)) & 1) ? exp2f(fmaf(S[j][2 + e], LOG2E, -o1)) : 0.f;
            sum0 += p0; sum1 += p1;
            S[j][e] = p0; S[j][2 + e] = p1;
        }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    const float inv0 = 1.f / sum0, inv1 = 1.f / sum1;

    // ---------------- O = P V (model/attention.py:207); P un-normalised in f16, 1/sum applied in fp32 --------------
    float O[8][4];
#pragma unroll
    for (int c = 0; c < 8; ++c) O[c][0] = O[c][1] = O[c][2] = O[c][3] = 0.f;
    {
        const uint32_t vb = s_u32(sV);
        const int mi = lane >> 3;
#pragma unroll
        for (int i = 0; i < Cf::NT16; ++i) {
            uint32_t pa[4];
            pa[0] = pack_h2(S[2 * i][0], S[2 * i][1]);
            pa[1] = pack_h2(S[2 * i][2], S[2 * i][3]);
            pa[2] = pack_h2(S[2 * i + 1][0], S[2 * i + 1][1]);
            pa[3] = pack_h2(S[2 * i + 1][2], S[2 * i + 1][3]);
            int nk = 16 * i + ((mi & 1) << 3) + (lane & 7);
            nk = nk < Cf::NK ? nk : Cf::NK - 1;
            const int pos = (4 * wy + nk / Cf::WN) * Cf::KC + 4 * wx + nk % Cf::WN;
#pragma unroll
            for (int cp = 0; cp < 4; ++cp) {
                uint32_t v[4];
                ldsm_x4_t(v, vb + swz_chunk(pos, 2 * cp + (mi >> 1)));
                mma16816(O[2 * cp], pa, v[0], v[1]);
                mma16816(O[2 * cp + 1], pa, v[2], v[3]);
            }
        }
    }

    // ---------------- fused = lr_up + O (model/attention.py:210) ----------------
    // thread (g,t): pixels A = block row g>>2, B = A + 2 rows; channels 8c+2t, 8c+2t+1
    const int pyA = y0 + 4 * wy + (g >> 2), pxA = x0 + 4 * wx + (g & 3), pyB = pyA + 2;
    {
        const float lsh = resize_scale(p.h, H, ARSEG_RESIZE_BILINEAR_AC), lsw = resize_scale(p.w, W, ARSEG_RESIZE_BILINEAR_AC);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int py = half ? pyB : pyA;
            const float inv = half ? inv1 : inv0;
            if (py < H && pxA < W) {
                int ya, yb, xa, xb; float lya, lyb, lxa, lxb;
                bilinear_src(lsh, py, p.h, ARSEG_RESIZE_BILINEAR_AC, ya, yb, lya, lyb);
                bilinear_src(lsw, pxA, p.w, ARSEG_RESIZE_BILINEAR_AC, xa, xb, lxa, lxb);
                const TLR* q00 = lr + ((size_t)ya * p.w + xa) * MC + 2 * t;
                const TLR* q01 = lr + ((size_t)ya * p.w + xb) * MC + 2 * t;
                const TLR* q10 = lr + ((size_t)yb * p.w + xa) * MC + 2 * t;
                const TLR* q11 = lr + ((size_t)yb * p.w + xb) * MC + 2 * t;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float2 a = ld2(q00 + 8 * c), bq_ = ld2(q01 + 8 * c), cq = ld2(q10 + 8 * c), d = ld2(q11 + 8 * c);
                    const float u0 = lya * (lxa * a.x + lxb * bq_.x) + lyb * (lxa * cq.x + lxb * d.x);
                    const float u1 = lya * (lxa * a.y + lxb * bq_.y) + lyb * (lxa * cq.y + lxb * d.y);
                    O[c][2 * half] = fmaf(O[c][2 * half], inv, u0);
                    O[c][2 * half + 1] = fmaf(O[c][2 * half + 1], inv, u1);
                }
            }
        }
    }
    if (p.out_p) {
        const size_t plane = (size_t)H * W;
#pragma unroll
        for (int c = 0; c < 8; ++c)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                float* op = p.out_p + ((size_t)n * MC + 8 * c + 2 * t + e) * plane;
                if (pyA < H && pxA < W) op[(size_t)pyA * W + pxA] = O[c][e];
                if (pyB < H && pxA < W) op[(size_t)pyB * W + pxA] = O[c][2 + e];
            }
    }
    if (!do_cls) return;

    // ---------------- classifier (model/pspnet.py:226) as a [16 x 64] x [64 x 8*nct] MMA ----------------
    uint32_t fa[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        fa[ks][0] = pack_h2(clamp_h(O[2 * ks][0]), clamp_h(O[2 * ks][1]));
        fa[ks][1] = pack_h2(clamp_h(O[2 * ks][2]), clamp_h(O[2 * ks][3]));
        fa[ks][2] = pack_h2(clamp_h(O[2 * ks + 1][0]), clamp_h(O[2 * ks + 1][1]));
        fa[ks][3] = pack_h2(clamp_h(O[2 * ks + 1][2]), clamp_h(O[2 * ks + 1][3]));
    }
    float Lg[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        Lg[nt][0] = Lg[nt][1] = Lg[nt][2] = Lg[nt][3] = 0.f;
        if (nt < nct) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const __half* wp = s_wc + (8 * nt + g) * MCLS_LD + 16 * ks + 2 * t;
                mma16816(Lg[nt], fa[ks], *reinterpret_cast<const uint32_t*>(wp), *reinterpret_cast<const uint32_t*>(wp + 8));
            }
        }
    }
    // bias, argmax (first maximum, like torch.argmax) and log-softmax per pixel: values of one pixel live in a quad
    float lmax0 = -INFINITY, lmax1 = -INFINITY;
    int am0 = 0, am1 = 0;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int cls = 8 * nt + 2 * t + e;
            if (nt < nct && cls < p.ncls) {
                Lg[nt][e] += s_bc[cls]; Lg[nt][2 + e] += s_bc[cls];
                if (Lg[nt][e] > lmax0) { lmax0 = Lg[nt][e]; am0 = cls; }
                if (Lg[nt][2 + e] > lmax1) { lmax1 = Lg[nt][2 + e]; am1 = cls; }
            }
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
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int cls = 8 * nt + 2 * t + e;
                if (nt < nct && cls < p.ncls) { lse0 += expf(Lg[nt][e] - lmax0); lse1 += expf(Lg[nt][2 + e] - lmax1); }
            }
        lse0 += __shfl_xor_sync(0xffffffffu, lse0, 1); lse0 += __shfl_xor_sync(0xffffffffu, lse0, 2);
        lse1 += __shfl_xor_sync(0xffffffffu, lse1, 1); lse1 += __shfl_xor_sync(0xffffffffu, lse1, 2);
        lse0 = logf(lse0) + lmax0; lse1 = logf(lse1) + lmax1;
    }
    // stage the logits planes in the (now idle) ring, then store whole 64-byte rows
    float* s_out = ring;                                               // [ncls][MOUT_LD]
    uint8_t* s_arg = reinterpret_cast<uint8_t*>(ring + 32 * MOUT_LD);  // [256]
    const int lpA = (4 * wy + (g >> 2)) * MT + 4 * wx + (g & 3), lpB = lpA + 2 * MT;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int cls = 8 * nt + 2 * t + e;
            if (nt < nct && cls < p.ncls) {
                s_out[cls * MOUT_LD + lpA] = Lg[nt][e] - lse0;
                s_out[cls * MOUT_LD + lpB] = Lg[nt][2 + e] - lse1;
            }
        }
    if (t == 0) { s_arg[lpA] = (uint8_t)am0; s_arg[lpB] = (uint8_t)am1; }
    __syncthreads();
    const size_t plane = (size_t)H * W;
    const bool full = (y0 + MT <= H) && (x0 + MT <= W) && (W % 4 == 0);
    if (p.out_logits) {
        float* ol = p.out_logits + (size_t)n * p.ncls * plane;
        if (full) {
            for (int i = tid; i < p.ncls * 64; i += MTHREADS) {
                const int cls = i >> 6, row = (i >> 2) & 15, q4 = i & 3;
                const float4 v = *reinterpret_cast<const float4*>(s_out + cls * MOUT_LD + row * MT + 4 * q4);
                *reinterpret_cast<float4*>(ol + cls * plane + (size_t)(y0 + row) * W + x0 + 4 * q4) = v;
            }
        } else {
            for (int i = tid; i < p.ncls * 256; i += MTHREADS) {
                const int cls = i >> 8, row = (i >> 4) & 15, col = i & 15;
                if (y0 + row < H && x0 + col < W) ol[cls * plane + (size_t)(y0 + row) * W + x0 + col] = s_out[cls * MOUT_LD + row * MT + col];
            }
        }
    }
    if (p.out_argmax) {
        uint8_t* oa = p.out_argmax + (size_t)n * plane;
        if (full) {
            if (tid < 64) {
                const int row = tid >> 2, q4 = tid & 3;
                *reinterpret_cast<uint32_t*>(oa + (size_t)(y0 + row) * W + x0 + 4 * q4) =
                    *reinterpret_cast<const uint32_t*>(s_arg + row * MT + 4 * q4);
            }
        } else if (tid < 256) {
            const int row = tid >> 4, col = tid & 15;
            if (y0 + row < H && x0 + col < W) oa[(size_t)(y0 + row) * W + x0 + col] = s_arg[row * MT + col];
        }
    }
}

template <int K, typename TLR>
static int creff_mma_launch_t(CreffMmaParams& p, cudaStream_t st) {
    using Cf = MCfg<K>;
    auto kern = creff_mma_kernel<K, TLR>;
    static bool configured[64] = {false};
    int dev = 0;
    ARSEG_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        ARSEG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cf::SMEM));
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    p.tiles_x = ceil_div(p.W, MT);
    p.tiles_y = ceil_div(p.H, MT);
    const long long blocks = (long long)p.tiles_x * p.tiles_y * p.N;
    ARSEG_REQUIRE(blocks > 0 && blocks < 2147483647LL, "creff_mma: grid too large");
    kern<<<(unsigned)blocks, MTHREADS, Cf::SMEM, st>>>(p);
    ARSEG_CHECK_LAUNCH("creff_mma");
    return ARSEG_OK;
}

bool creff_mma_supported(const arseg_creff_args* a) {
    return a->C == MC && a->hr_layout == ARSEG_NHWC && a->lr_layout == ARSEG_NHWC && (a->k == 3 || a->k == 5 || a->k == 7 || a->k == 9) &&
           (!a->wcls || a->ncls <= 32) && ((size_t)a->H * a->W < (1u << 29)) && ((size_t)a->h * a->w < (1u << 29));
}

int creff_march_launch(CreffMmaParams& p, int k, int lr_dtype, cudaStream_t st);   // creff_march.cu

// ARSEG_CREFF_MMA_F16 runs the column-marching engine (creff_march.cu); ARSEG_CREFF_TILE=1 selects the older
// square-tile engine of this file (kept for A/B measurements).
static bool use_tile_engine() {
    const char* e = getenv("ARSEG_CREFF_TILE");
    return e && e[0] == '1';
}

int creff_mma_launch(const arseg_creff_args* a, cudaStream_t st) {
    CreffMmaParams p;
    p.hr = a->hr; p.hr_shared = a->hr_shared; p.flow = a->flow; p.flow_dtype = a->flow_dtype; p.Hm = a->Hm; p.Wm = a->Wm;
    p.lr = a->lr; p.h = a->h; p.w = a->w;
    p.wq = a->wq; p.bq = a->bq; p.wk = a->wk; p.bk = a->bk; p.wv = a->wv; p.bv = a->bv; p.wcls = a->wcls; p.bcls = a->bcls;
    p.ncls = a->ncls; p.log_softmax = a->log_softmax; p.out_p = a->out_p; p.out_logits = a->out_logits;
    p.out_argmax = a->out_argmax; p.N = a->N; p.C = a->C; p.H = a->H; p.W = a->W;
    const bool bf = a->lr_dtype == ARSEG_BF16;
    if (!use_tile_engine() || a->lr_dtype == ARSEG_F16) return creff_march_launch(p, a->k, a->lr_dtype, st);
    switch (a->k) {
        case 3: return bf ? creff_mma_launch_t<3, __nv_bfloat16>(p, st) : creff_mma_launch_t<3, float>(p, st);
        case 5: return bf ? creff_mma_launch_t<5, __nv_bfloat16>(p, st) : creff_mma_launch_t<5, float>(p, st);
        case 7: return bf ? creff_mma_launch_t<7, __nv_bfloat16>(p, st) : creff_mma_launch_t<7, float>(p, st);
        case 9: return bf ? creff_mma_launch_t<9, __nv_bfloat16>(p, st) : creff_mma_launch_t<9, float>(p, st);
        default: ARSEG_UNSUPPORTED("creff_mma: window k=%d", a->k);
    }
}

}  // namespace arseg
