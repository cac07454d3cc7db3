// creff_march.cu -- fused MV-warp + CReFF + classifier, column-marching tensor-core engine (sm_100a).
//
// Same contract and arithmetic as creff_mma.cu (reference: evaluation.py:177-183 MV rescale + warpFeature,
// model/attention.py:184-213 MyAttention.forward, model/pspnet.py:226-229 final_conv + LogSoftmax,
// evaluation.py:204 argmax), C = 64, NHWC operands.  What changes is the decomposition:
//
//   * a CTA owns one 16-pixel-wide column strip of one frame and MARCHES down it 4 rows per step.  K, V live in a
//     shared-memory ring of image rows (f16, 128 B per position), so every warped-hr row is gathered once and
//     every K/V row is convolved once per strip: the only redundancy left is the horizontal halo
//     ((16+k+1)/16 gather, (16+k-1)/16 depthwise) instead of the (16+k+1)^2/256 of a square tile.
//   * three warp-specialised roles run CONCURRENTLY on different steps of the march, handing rows over through
//     mbarriers (4-deep, indexed by step):
//       G (8 warps): bilinear gather of the MV-warped hr rows and the lr_up rows into fp32 row rings
//                    (half-warp per position, float4 per lane; f64 MV arithmetic one position per thread);
//       D (4 warps): the three depthwise 3x3 convolutions (FFMA2, x-marching register window): K,V -> f16
//                    rings, Q -> f16 tile, lr_up centre (the residual) -> fp32 tile;
//       C (4 warps): one 4x4-pixel block each: S = Q K^T (mma.sync m16n8k16 f16, fp32 accumulate) on the
//                    (k+3)^2 key patch, masked softmax in registers, O = resid*sum + P V, classifier MMA,
//                    log-softmax, argmax, stores.
//     Step t of G feeds step t of D feeds step t-1 of C, so the LDG latency of the gather, the FFMA2 stream of
//     the convolutions and the MMA/MUFU stream of the attention overlap inside one SM.
//   * work split: grid = frames x column strips x row segments, frame index fastest (the frames of a GOP read
//     the same keyframe rows back to back -> L2 reuse).
#include "creff_mma_common.cuh"
#include <cstdlib>

namespace arseg {

constexpr int XSW = 16;                 // strip width (pixels)
constexpr int XTHREADS = 512;
constexpr int XG_THREADS = 256, XD_THREADS = 128, XC_THREADS = 128;
constexpr int XHR_RING = 10, XLR_RING = 10;   // fp32 row rings (rows)
constexpr int XRES_LD = 72;             // floats per residual row (bank-conflict pad)
constexpr int XCLS_LD = 72;             // f16 per classifier-weight row
constexpr int XJA = 4;                  // gather positions in flight per half-warp
constexpr int XNB = 4;                  // mbarriers per hand-off (indexed by step & 3)
constexpr int XBAR_G = 1;               // named barrier of the G group

template <int K> struct XCfg {
    static constexpr int R = K / 2;
    static constexpr int KVC = XSW + K - 1;                 // K/V ring columns
    static constexpr int HC = KVC + 2;                      // warped-hr ring columns
    static constexpr int LC = XSW + 2;                      // lr_up ring columns
    static constexpr int P0 = 2 * R - 4 > 0 ? 2 * R - 4 : 0;   // K/V rows of the initial D step
    static constexpr int G0 = P0 + 2;                       // hr rows of the initial G step
    static constexpr int SL = K <= 7 ? 3 : 2;               // D runs at most SL steps ahead of C's K/V reads
    static constexpr int KVR = P0 + 4 * SL;                 // K/V ring rows
    static constexpr int WN = K + 3, NK = WN * WN;          // key patch of a 4x4 block
    static constexpr int NT16 = (NK + 15) / 16, NT8 = 2 * NT16;
    static constexpr int PMAX = (G0 * HC > 4 * HC + 4 * LC) ? G0 * HC : 4 * HC + 4 * LC;   // positions per G step
    static constexpr size_t KV_BYTES = (size_t)KVR * KVC * 128;
    static constexpr size_t HR_BYTES = (size_t)XHR_RING * HC * 256;
    static constexpr size_t LR_BYTES = (size_t)XLR_RING * LC * 256;
    static constexpr size_t Q_BYTES = 64 * 128;
    static constexpr size_t RES_BYTES = 64 * XRES_LD * 4;
    static constexpr size_t POS_BYTES = 2 * (size_t)PMAX * 24;
    static constexpr size_t CLS_BYTES = 32 * XCLS_LD * 2 + 32 * 4;
    static constexpr size_t SMEM = 2 * KV_BYTES + HR_BYTES + LR_BYTES + Q_BYTES + RES_BYTES + POS_BYTES + CLS_BYTES + 4 * XNB * 8;
    static_assert(4 + 2 * R <= KVR, "K/V ring too small for one consumer step");
    static_assert(SMEM <= 232448, "shared memory budget");
};

// ---------------------------------------------------------------------------------------------
// mbarrier helpers (hand-off between the roles; index = march step + 1)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void xbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void xbar_arrive(uint64_t* bars, int step) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(bars + ((step + 1) & (XNB - 1)))) : "memory");
}
__device__ __forceinline__ void xbar_wait(uint64_t* bars, int step) {
    const uint32_t addr = s_u32(bars + ((step + 1) & (XNB - 1)));
    const uint32_t parity = (uint32_t)(((step + 1) / XNB) & 1);
    uint32_t ok = 0;
    long long t0 = 0;
    for (int spin = 0; !ok; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (!ok && (spin & 1023) == 1023) {          // bounded: a protocol bug must not hang the GPU box
            if (t0 == 0) t0 = clock64();
            else if (clock64() - t0 > 4000000000LL) {
                printf("arseg creff_march: mbarrier wait timed out (block %d thread %d step %d)\n", (int)blockIdx.x, (int)threadIdx.x, step);
                __trap();
            }
        }
    }
}

// byte offset of 16-byte chunk `chunk` of K/V ring position (slot row, col): the swizzle key is the column only,
// so a row's offset is a pure function of (slot, col) and ldmatrix over 8 consecutive columns is conflict free
__device__ __forceinline__ uint32_t kv_off(int pos, int col, int chunk) { return (uint32_t)(pos * 128 + (((chunk ^ col) & 7) << 4)); }
// Q tile [row 0..3][col 0..15]: key = (col & 3) | (row & 1) << 2 -> the 8 rows of every ldmatrix 8x8 are distinct
__device__ __forceinline__ uint32_t q_off(int row, int col, int chunk) {
    return (uint32_t)((row * XSW + col) * 128 + (((chunk ^ ((col & 3) | ((row & 1) << 2))) & 7) << 4));
}

struct XSmem {
    uint8_t *sK, *sV, *rings, *sQ;
    float* sRes;
    float4* posw;
    int2* posid;
    __half* s_wc;
    float* s_bc;
    uint64_t *gfull, *ddone, *cdone, *qlempty;
};

// ---------------------------------------------------------------------------------------------
// G role: gather.  Step t (t = -1 .. S): hr rows [h0, h0+nh) into the hr ring, lr_up rows [l0, l0+nl) into the
// lr ring.  Row indices are relative to the segment: hr row r <-> image row ya-R-1+r, lr row r <-> ya-1+r.
// ---------------------------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ void x_step_geom(int t, int& h0, int& nh, int& l0, int& nl) {
    using Cf = XCfg<K>;
    if (t < 0) { h0 = 0; nh = Cf::G0; l0 = 0; nl = 0; }
    else {
        h0 = Cf::G0 + 4 * t; nh = 4;
        if (t == 0) { l0 = 0; nl = 2; } else { l0 = 4 * t - 2; nl = 4; }
    }
}

template <int K, typename TLR>
__device__ __forceinline__ void x_g_role(const CreffMmaParams& p, const XSmem& sm, int n, int x0, int ya, int S) {
    using Cf = XCfg<K>;
    const int gt = threadIdx.x, lane = gt & 31, hw = gt >> 4, cl = lane & 15;
    const float lsh = resize_scale(p.h, p.H, ARSEG_RESIZE_BILINEAR_AC), lsw = resize_scale(p.w, p.W, ARSEG_RESIZE_BILINEAR_AC);
    const float* __restrict__ hr = p.hr + (p.hr_shared ? 0 : (size_t)n * p.H * p.W * MC) + 4 * cl;
    const TLR* __restrict__ lr = reinterpret_cast<const TLR*>(p.lr) + (size_t)n * p.h * p.w * MC + 4 * cl;
    const int hr_rs = p.W * MC, lr_rs = p.w * MC;

    auto compute_pos = [&](int t) {
        int h0, nh, l0, nl;
        x_step_geom<K>(t, h0, nh, l0, nl);
        const int nhp = nh * Cf::HC;
        if (gt < nhp + nl * Cf::LC) {
            PosRec r; int dst;
            if (gt < nhp) {
                const int rr = gt / Cf::HC, cc = gt - rr * Cf::HC, row = h0 + rr;
                r = pos_hr(p, n, ya - Cf::R - 1 + row, x0 - Cf::R - 1 + cc);
                dst = ((row % XHR_RING) * Cf::HC + cc) * 256;
            } else {
                const int q = gt - nhp, rr = q / Cf::LC, cc = q - rr * Cf::LC, row = l0 + rr;
                r = pos_lr(p, lsh, lsw, ya - 1 + row, x0 - 1 + cc);
                dst = ((int)Cf::HR_BYTES + ((row % XLR_RING) * Cf::LC + cc) * 256) | (1 << 30);
            }
            const int buf = (t + 1) & 1;
            sm.posw[buf * Cf::PMAX + gt] = r.w;
            sm.posid[buf * Cf::PMAX + gt] = make_int2(r.info, dst);
        }
    };
    float4 tap[XJA][4];
    auto issue = [&](int buf, int npos, int j0) {
#pragma unroll
        for (int j = 0; j < XJA; ++j) {
            const int i = hw + 16 * (j0 + j);
            int2 id = make_int2(-1, 0);
            if (i < npos) id = sm.posid[buf * Cf::PMAX + i];
            if (id.x >= 0) {
                const int dx = (id.x >> 1) & 1, dy = id.x & 1;
                const size_t pix = (size_t)(id.x >> 2) * MC;
                if (!(id.y >> 30)) {
                    const float* s = hr + pix;
                    const float* s2 = s + dy * hr_rs;
                    tap[j][0] = ld4(s); tap[j][1] = ld4(s + dx * MC); tap[j][2] = ld4(s2); tap[j][3] = ld4(s2 + dx * MC);
                } else {
                    const TLR* s = lr + pix;
                    const TLR* s2 = s + dy * lr_rs;
                    tap[j][0] = ld4(s); tap[j][1] = ld4(s + dx * MC); tap[j][2] = ld4(s2); tap[j][3] = ld4(s2 + dx * MC);
                }
            } else {
                tap[j][0] = tap[j][1] = tap[j][2] = tap[j][3] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    };
    auto commit = [&](int buf, int npos, int j0) {
#pragma unroll
        for (int j = 0; j < XJA; ++j) {
            const int i = hw + 16 * (j0 + j);
            if (i < npos) {
                const float4 w = sm.posw[buf * Cf::PMAX + i];
                const int dst = sm.posid[buf * Cf::PMAX + i].y & 0x3fffffff;
                float4 v;
                v.x = tap[j][0].x * w.x + tap[j][1].x * w.y + tap[j][2].x * w.z + tap[j][3].x * w.w;
                v.y = tap[j][0].y * w.x + tap[j][1].y * w.y + tap[j][2].y * w.z + tap[j][3].y * w.w;
                v.z = tap[j][0].z * w.x + tap[j][1].z * w.y + tap[j][2].z * w.z + tap[j][3].z * w.w;
                v.w = tap[j][0].w * w.x + tap[j][1].w * w.y + tap[j][2].w * w.z + tap[j][3].w * w.w;
                *reinterpret_cast<float4*>(sm.rings + dst + 16 * cl) = v;
            }
        }
    };

    compute_pos(-1);
    nbar_sync(XBAR_G, XG_THREADS);
#pragma unroll 1
    for (int t = -1; t <= S; ++t) {
        const int buf = (t + 1) & 1;
        int h0, nh, l0, nl;
        x_step_geom<K>(t, h0, nh, l0, nl);
        const int npos = nh * Cf::HC + nl * Cf::LC;
        if (t >= 1) xbar_wait(sm.ddone, t - 2);       // D step t-2 done: the ring rows this step overwrites are free
        issue(buf, npos, 0);
        if (t < S) compute_pos(t + 1);                // f64 MV arithmetic overlaps the loads in flight
        commit(buf, npos, 0);
#pragma unroll 1
        for (int j0 = XJA; 16 * j0 < npos; j0 += XJA) {
            issue(buf, npos, j0);
            commit(buf, npos, j0);
        }
        nbar_sync(XBAR_G, XG_THREADS);                // position records of step t+1 visible to the group
        xbar_arrive(sm.gfull, t);
    }
}

// ---------------------------------------------------------------------------------------------
// D role: depthwise 3x3 convolutions.  Warp dw owns row dw of every 4-row strip.
//   NOUT = 2: K/V row from the hr ring (input rows kr..kr+2, columns x..x+2); NOUT = 1: Q row + residual from lr ring.
// ---------------------------------------------------------------------------------------------
template <int NOUT, int INC, int OUTC, typename Store>
__device__ __forceinline__ void x_dw_row(const float* rp0, const float* rp1, const float* rp2, const float2 (&w1)[9], const float2 b1,
                                         const float2 (&w2)[9], const float2 b2, Store&& store) {
    float2 win[3][3];   // [input row][slot]; slot (x + d) % 3 holds input column x + d
    win[0][0] = *reinterpret_cast<const float2*>(rp0);
    win[1][0] = *reinterpret_cast<const float2*>(rp1);
    win[2][0] = *reinterpret_cast<const float2*>(rp2);
    win[0][1] = *reinterpret_cast<const float2*>(rp0 + MC);
    win[1][1] = *reinterpret_cast<const float2*>(rp1 + MC);
    win[2][1] = *reinterpret_cast<const float2*>(rp2 + MC);
#pragma unroll 1
    for (int xb = 0; xb < OUTC; xb += 3) {
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int x = xb + u;
            if (x < OUTC) {
                const int sa = u, sb = (u + 1) % 3, sc = (u + 2) % 3;     // slots of columns x, x+1, x+2
                win[0][sc] = *reinterpret_cast<const float2*>(rp0 + (x + 2) * MC);
                win[1][sc] = *reinterpret_cast<const float2*>(rp1 + (x + 2) * MC);
                win[2][sc] = *reinterpret_cast<const float2*>(rp2 + (x + 2) * MC);
                // three independent row chains per output (shorter dependency chains than one 9-deep chain)
                float2 a0 = __ffma2_rn(w1[0], win[0][sa], b1), a1 = __fmul2_rn(w1[3], win[1][sa]), a2 = __fmul2_rn(w1[6], win[2][sa]);
                a0 = __ffma2_rn(w1[1], win[0][sb], a0); a1 = __ffma2_rn(w1[4], win[1][sb], a1); a2 = __ffma2_rn(w1[7], win[2][sb], a2);
                a0 = __ffma2_rn(w1[2], win[0][sc], a0); a1 = __ffma2_rn(w1[5], win[1][sc], a1); a2 = __ffma2_rn(w1[8], win[2][sc], a2);
                const float2 r1 = __fadd2_rn(__fadd2_rn(a0, a1), a2);
                float2 r2 = make_float2(0.f, 0.f);
                if (NOUT == 2) {
                    float2 c0 = __ffma2_rn(w2[0], win[0][sa], b2), c1 = __fmul2_rn(w2[3], win[1][sa]), c2 = __fmul2_rn(w2[6], win[2][sa]);
                    c0 = __ffma2_rn(w2[1], win[0][sb], c0); c1 = __ffma2_rn(w2[4], win[1][sb], c1); c2 = __ffma2_rn(w2[7], win[2][sb], c2);
                    c0 = __ffma2_rn(w2[2], win[0][sc], c0); c1 = __ffma2_rn(w2[5], win[1][sc], c1); c2 = __ffma2_rn(w2[8], win[2][sc], c2);
                    r2 = __fadd2_rn(__fadd2_rn(c0, c1), c2);
                }
                store(x, r1, r2, win[1][sb]);
            }
        }
    }
}

template <int K>
__device__ __forceinline__ void x_d_role(const CreffMmaParams& p, const XSmem& sm, int x0, int ya, int S) {
    using Cf = XCfg<K>;
    const int lane = threadIdx.x & 31, dw = (threadIdx.x >> 5) - 8;
    float2 wk[9], wv[9], wq[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        wk[t] = make_float2(__ldg(p.wk + (2 * lane) * 9 + t), __ldg(p.wk + (2 * lane + 1) * 9 + t));
        wv[t] = make_float2(__ldg(p.wv + (2 * lane) * 9 + t), __ldg(p.wv + (2 * lane + 1) * 9 + t));
        wq[t] = make_float2(__ldg(p.wq + (2 * lane) * 9 + t), __ldg(p.wq + (2 * lane + 1) * 9 + t));
    }
    const float2 bk = make_float2(__ldg(p.bk + 2 * lane), __ldg(p.bk + 2 * lane + 1));
    const float2 bv = make_float2(__ldg(p.bv + 2 * lane), __ldg(p.bv + 2 * lane + 1));
    const float2 bq = make_float2(__ldg(p.bq + 2 * lane), __ldg(p.bq + 2 * lane + 1));
    const float* hring = reinterpret_cast<const float*>(sm.rings) + 2 * lane;
    const float* lring = reinterpret_cast<const float*>(sm.rings + Cf::HR_BYTES) + 2 * lane;
    const int lane_chunk = lane >> 2;
    const uint32_t lane_off = (uint32_t)((lane & 3) * 4);
#pragma unroll 1
    for (int t = -1; t <= S; ++t) {
        xbar_wait(sm.gfull, t);
        // ---- K/V row kr (relative to image row ya-R) from hr ring rows kr..kr+2 ----
        if (t - Cf::SL >= 0) xbar_wait(sm.cdone, t - Cf::SL);     // the ring rows this step overwrites have been read
        const int kr = t < 0 ? dw : Cf::P0 + 4 * t + dw;
        if (t >= 0 || dw < Cf::P0) {
            const float* rp0 = hring + ((kr) % XHR_RING) * (Cf::HC * MC);
            const float* rp1 = hring + ((kr + 1) % XHR_RING) * (Cf::HC * MC);
            const float* rp2 = hring + ((kr + 2) % XHR_RING) * (Cf::HC * MC);
            const int fy = ya - Cf::R + kr;
            const bool row_ok = fy >= 0 && fy < p.H;
            const int pos0 = (kr % Cf::KVR) * Cf::KVC;
            const int fx0 = x0 - Cf::R;
            x_dw_row<2, Cf::HC, Cf::KVC>(rp0, rp1, rp2, wk, bk, wv, bv, [&](int x, float2 a1, float2 a2, float2) {
                // K / V are exactly 0 outside the image (attention zero padding, model/attention.py:199,207)
                const int fx = fx0 + x;
                if (!(row_ok && fx >= 0 && fx < p.W)) { a1 = make_float2(0.f, 0.f); a2 = make_float2(0.f, 0.f); }
                const uint32_t off = kv_off(pos0 + x, x, lane_chunk) + lane_off;
                *reinterpret_cast<uint32_t*>(sm.sK + off) = pack_h2_sat(a1.x, a1.y);
                *reinterpret_cast<uint32_t*>(sm.sV + off) = pack_h2_sat(a2.x, a2.y);
            });
        }
        // ---- Q row qr = 4(t-1)+dw (relative to ya) from lr ring rows qr..qr+2; residual = lr_up centre ----
        if (t >= 1) {
            if (t - 2 >= 0) xbar_wait(sm.qlempty, t - 2);         // C step t-2 has taken Q / residual into registers
            const int qr = 4 * (t - 1) + dw;
            const float* rp0 = lring + ((qr) % XLR_RING) * (Cf::LC * MC);
            const float* rp1 = lring + ((qr + 1) % XLR_RING) * (Cf::LC * MC);
            const float* rp2 = lring + ((qr + 2) % XLR_RING) * (Cf::LC * MC);
            x_dw_row<1, Cf::LC, XSW>(rp0, rp1, rp2, wq, bq, wq, bq, [&](int x, float2 a1, float2, float2 centre) {
                *reinterpret_cast<uint32_t*>(sm.sQ + q_off(dw, x, lane_chunk) + lane_off) = pack_h2_sat(a1.x, a1.y);
                *reinterpret_cast<float2*>(sm.sRes + (dw * XSW + x) * XRES_LD + 2 * lane) = centre;
            });
        }
        xbar_arrive(sm.ddone, t);
    }
}

// ---------------------------------------------------------------------------------------------
// C role: attention + classifier for one 4x4 block per warp per step.
// ---------------------------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ void x_c_role(const CreffMmaParams& p, const XSmem& sm, int n, int x0, int ya, int yb, int S) {
    using Cf = XCfg<K>;
    const int lane = threadIdx.x & 31, cw = (threadIdx.x >> 5) - 12;
    const int g = lane >> 2, t = lane & 3, mi = lane >> 3;
    const int H = p.H, W = p.W;
    const bool do_cls = p.wcls != nullptr;
    const int nct = do_cls ? (p.ncls + 7) >> 3 : 0;
    const size_t plane = (size_t)H * W;

    // validity masks of this thread's logits: rows g (m0) and g+8 (m1), keys 8j+2t+e -> bit 2j+e
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
    const uint32_t kb = s_u32(sm.sK), vb = s_u32(sm.sV), qb = s_u32(sm.sQ);
    const int pxA = x0 + 4 * cw + (g & 3);
    int s4 = 0;                                           // (4 s) mod KVR
    // the mbarrier phase of a step is (step+1)/XNB: C has no step -1, so arrive for it once (nobody waits on it)
    xbar_arrive(sm.cdone, -1);
    xbar_arrive(sm.qlempty, -1);
#pragma unroll 1
    for (int s = 0; s < S; ++s) {
        xbar_wait(sm.ddone, s + 1);
        // ---------------- Q fragments ----------------
        uint32_t qa[4][4];
        {
            const int r = ((mi & 1) << 3) + (lane & 7);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) ldsm_x4(qa[ks], qb + q_off(r >> 2, 4 * cw + (r & 3), 2 * ks + (mi >> 1)));
        }
        // ---------------- S = Q K^T (model/attention.py:199) ----------------
        float Sx[Cf::NT8][4];
#pragma unroll
        for (int j = 0; j < Cf::NT8; ++j) {
            Sx[j][0] = Sx[j][1] = Sx[j][2] = Sx[j][3] = 0.f;
            int nk = 8 * j + (lane & 7);
            nk = nk < Cf::NK ? nk : Cf::NK - 1;
            const int ky = nk / Cf::WN, kx = nk - ky * Cf::WN;
            int slot = s4 + ky;
            slot = slot >= Cf::KVR ? slot - Cf::KVR : slot;
            const int col = 4 * cw + kx, pos = slot * Cf::KVC + col;
            uint32_t b0[4], b1[4];
            ldsm_x4(b0, kb + kv_off(pos, col, mi));
            ldsm_x4(b1, kb + kv_off(pos, col, 4 + mi));
            mma16816(Sx[j], qa[0], b0[0], b0[1]);
            mma16816(Sx[j], qa[1], b0[2], b0[3]);
            mma16816(Sx[j], qa[2], b1[0], b1[1]);
            mma16816(Sx[j], qa[3], b1[2], b1[3]);
        }
        // ---------------- residual lr_up (model/attention.py:191,210) into the O accumulators ----------------
        // thread (g,t): pixels A = block row g>>2, B = A + 2 rows; channels 8c+2t, 8c+2t+1
        float O[8][4];
        {
            const float* ra = sm.sRes + ((g >> 2) * XSW + 4 * cw + (g & 3)) * XRES_LD + 2 * t;
            const float* rb = ra + 2 * XSW * XRES_LD;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float2 a = *reinterpret_cast<const float2*>(ra + 8 * c), b = *reinterpret_cast<const float2*>(rb + 8 * c);
                O[c][0] = a.x; O[c][1] = a.y; O[c][2] = b.x; O[c][3] = b.y;
            }
        }
        xbar_arrive(sm.qlempty, s);
        // ---------------- softmax over the k*k window of every query (model/attention.py:203) ----------------
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < Cf::NT8; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                if ((m0 >> (2 * j + e)) & 1) mx0 = fmaxf(mx0, Sx[j][e]);
                if ((m1 >> (2 * j + e)) & 1) mx1 = fmaxf(mx1, Sx[j][2 + e]);
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
                const float p0 = ((m0 >> (2 * j + e)) & 1) ? exp2f(fmaf(Sx[j][e], LOG2E, -o0)) : 0.f;
                const float p1 = ((m1 >> (2 * j + e)) & 1) ? exp2f(fmaf(Sx[j][2 + e], LOG2E, -o1)) : 0.f;
                sum0 += p0; sum1 += p1;
                Sx[j][e] = p0; Sx[j][2 + e] = p1;
            }
        sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
        sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
        const float inv0 = 1.f / sum0, inv1 = 1.f / sum1;
        // ---------------- O = resid * sum + P V (model/attention.py:207); P un-normalised f16, 1/sum in fp32 --------
#pragma unroll
        for (int c = 0; c < 8; ++c) { O[c][0] *= sum0; O[c][1] *= sum0; O[c][2] *= sum1; O[c][3] *= sum1; }
#pragma unroll
        for (int i = 0; i < Cf::NT16; ++i) {
            uint32_t pa[4];
            pa[0] = pack_h2(Sx[2 * i][0], Sx[2 * i][1]);
            pa[1] = pack_h2(Sx[2 * i][2], Sx[2 * i][3]);
            pa[2] = pack_h2(Sx[2 * i + 1][0], Sx[2 * i + 1][1]);
            pa[3] = pack_h2(Sx[2 * i + 1][2], Sx[2 * i + 1][3]);
            int nk = 16 * i + ((mi & 1) << 3) + (lane & 7);
            nk = nk < Cf::NK ? nk : Cf::NK - 1;
            const int ky = nk / Cf::WN, kx = nk - ky * Cf::WN;
            int slot = s4 + ky;
            slot = slot >= Cf::KVR ? slot - Cf::KVR : slot;
            const int col = 4 * cw + kx, pos = slot * Cf::KVC + col;
#pragma unroll
            for (int cp = 0; cp < 4; ++cp) {
                uint32_t v[4];
                ldsm_x4_t(v, vb + kv_off(pos, col, 2 * cp + (mi >> 1)));
                mma16816(O[2 * cp], pa, v[0], v[1]);
                mma16816(O[2 * cp + 1], pa, v[2], v[3]);
            }
        }
        xbar_arrive(sm.cdone, s);
        s4 += 4; s4 = s4 >= Cf::KVR ? s4 - Cf::KVR : s4;
#pragma unroll
        for (int c = 0; c < 8; ++c) { O[c][0] *= inv0; O[c][1] *= inv0; O[c][2] *= inv1; O[c][3] *= inv1; }

        const int pyA = ya + 4 * s + (g >> 2), pyB = pyA + 2;
        const bool okA = pyA < yb && pxA < W, okB = pyB < yb && pxA < W;
        if (p.out_p) {
#pragma unroll
            for (int c = 0; c < 8; ++c)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    float* op = p.out_p + ((size_t)n * MC + 8 * c + 2 * t + e) * plane;
                    if (okA) op[(size_t)pyA * W + pxA] = O[c][e];
                    if (okB) op[(size_t)pyB * W + pxA] = O[c][2 + e];
                }
        }
        if (!do_cls) continue;

        // ---------------- classifier (model/pspnet.py:226) as a [16 x 64] x [64 x 8*nct] MMA ----------------
        uint32_t fa[4][4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            fa[ks][0] = pack_h2_sat(O[2 * ks][0], O[2 * ks][1]);
            fa[ks][1] = pack_h2_sat(O[2 * ks][2], O[2 * ks][3]);
            fa[ks][2] = pack_h2_sat(O[2 * ks + 1][0], O[2 * ks + 1][1]);
            fa[ks][3] = pack_h2_sat(O[2 * ks + 1][2], O[2 * ks + 1][3]);
        }
        float Lg[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            Lg[nt][0] = Lg[nt][1] = Lg[nt][2] = Lg[nt][3] = 0.f;
            if (nt < nct) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const __half* wp = sm.s_wc + (8 * nt + g) * XCLS_LD + 16 * ks + 2 * t;
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
                    const float bc = sm.s_bc[cls];
                    Lg[nt][e] += bc; Lg[nt][2 + e] += bc;
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
        if (p.out_logits) {
            float* ol = p.out_logits + (size_t)n * p.ncls * plane;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int cls = 8 * nt + 2 * t + e;
                    if (nt < nct && cls < p.ncls) {
                        if (okA) ol[cls * plane + (size_t)pyA * W + pxA] = Lg[nt][e] - lse0;
                        if (okB) ol[cls * plane + (size_t)pyB * W + pxA] = Lg[nt][2 + e] - lse1;
                    }
                }
        }
        if (p.out_argmax && t == 0) {
            uint8_t* oa = p.out_argmax + (size_t)n * plane;
            if (okA) oa[(size_t)pyA * W + pxA] = (uint8_t)am0;
            if (okB) oa[(size_t)pyB * W + pxA] = (uint8_t)am1;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
template <int K, typename TLR>
__global__ void __launch_bounds__(XTHREADS, 1) creff_march_kernel(CreffMmaParams p) {
    using Cf = XCfg<K>;
    extern __shared__ __align__(1024) uint8_t xsm[];
    XSmem sm;
    sm.sK = xsm;
    sm.sV = sm.sK + Cf::KV_BYTES;
    sm.rings = sm.sV + Cf::KV_BYTES;                                        // hr ring, then lr ring
    sm.sQ = sm.rings + Cf::HR_BYTES + Cf::LR_BYTES;
    sm.sRes = reinterpret_cast<float*>(sm.sQ + Cf::Q_BYTES);
    sm.posw = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(sm.sRes) + Cf::RES_BYTES);
    sm.posid = reinterpret_cast<int2*>(sm.posw + 2 * Cf::PMAX);
    sm.s_wc = reinterpret_cast<__half*>(sm.posid + 2 * Cf::PMAX);           // [32][XCLS_LD]
    sm.s_bc = reinterpret_cast<float*>(sm.s_wc + 32 * XCLS_LD);             // [32]
    sm.gfull = reinterpret_cast<uint64_t*>(sm.s_bc + 32);
    sm.ddone = sm.gfull + XNB;
    sm.cdone = sm.ddone + XNB;
    sm.qlempty = sm.cdone + XNB;

    const int tid = threadIdx.x, warp = tid >> 5;
    // frame index fastest: the N frames of a GOP visit the same keyframe rows back to back (L2 reuse)
    int b = blockIdx.x;
    const int n = b % p.N; b /= p.N;
    const int x0 = (b % p.ncols) * XSW;
    const int ya = (b / p.ncols) * p.seg_rows;
    const int yb = min(ya + p.seg_rows, p.H);
    const int S = (yb - ya + 3) >> 2;

    if (tid == 0) {
        for (int i = 0; i < XNB; ++i) {
            xbar_init(sm.gfull + i, XG_THREADS);
            xbar_init(sm.ddone + i, XD_THREADS);
            xbar_init(sm.cdone + i, XC_THREADS);
            xbar_init(sm.qlempty + i, XC_THREADS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (p.wcls) {
        for (int i = tid; i < 32 * MC; i += XTHREADS) {
            const int j = i / MC, c = i % MC;
            sm.s_wc[j * XCLS_LD + c] = __float2half_rn(j < p.ncls ? clamp_h(__ldg(p.wcls + (size_t)j * MC + c)) : 0.f);
        }
        if (tid < 32) sm.s_bc[tid] = (tid < p.ncls && p.bcls) ? __ldg(p.bcls + tid) : 0.f;
    }
    __syncthreads();

    if (warp < 8) x_g_role<K, TLR>(p, sm, n, x0, ya, S);
    else if (warp < 12) x_d_role<K>(p, sm, x0, ya, S);
    else x_c_role<K>(p, sm, n, x0, ya, yb, S);
}

template <int K, typename TLR>
static int creff_march_launch_t(CreffMmaParams& p, cudaStream_t st) {
    using Cf = XCfg<K>;
    auto kern = creff_march_kernel<K, TLR>;
    static bool configured[64] = {false};
    int dev = 0;
    ARSEG_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        ARSEG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cf::SMEM));
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    p.ncols = ceil_div(p.W, XSW);
    // row segments: enough CTAs for >= ~6 waves of one-CTA-per-SM, but segments of >= 48 rows (each segment pays
    // ~K+5 redundant halo rows and a 3-step pipeline fill)
    const int sms = sm_count() > 0 ? sm_count() : 148;
    int nseg = 1;
    const char* e = getenv("ARSEG_CREFF_SEG_ROWS");
    if (e && atoi(e) >= 4) nseg = ceil_div(p.H, (atoi(e) + 3) / 4 * 4);
    else while ((long long)p.N * p.ncols * nseg < 6LL * sms && ceil_div(p.H, nseg + 1) >= 48) ++nseg;
    p.seg_rows = (ceil_div(p.H, nseg) + 3) / 4 * 4;
    p.nseg = ceil_div(p.H, p.seg_rows);
    const long long blocks = (long long)p.N * p.ncols * p.nseg;
    ARSEG_REQUIRE(blocks > 0 && blocks < 2147483647LL, "creff_march: grid too large");
    kern<<<(unsigned)blocks, XTHREADS, Cf::SMEM, st>>>(p);
    ARSEG_CHECK_LAUNCH("creff_march");
    return ARSEG_OK;
}

int creff_march_launch(CreffMmaParams& p, int k, bool lr_bf16, cudaStream_t st) {
    switch (k) {
        case 3: return lr_bf16 ? creff_march_launch_t<3, __nv_bfloat16>(p, st) : creff_march_launch_t<3, float>(p, st);
        case 5: return lr_bf16 ? creff_march_launch_t<5, __nv_bfloat16>(p, st) : creff_march_launch_t<5, float>(p, st);
        case 7: return lr_bf16 ? creff_march_launch_t<7, __nv_bfloat16>(p, st) : creff_march_launch_t<7, float>(p, st);
        case 9: return lr_bf16 ? creff_march_launch_t<9, __nv_bfloat16>(p, st) : creff_march_launch_t<9, float>(p, st);
        default: ARSEG_UNSUPPORTED("creff_march: window k=%d", k);
    }
}

}  // namespace arseg
