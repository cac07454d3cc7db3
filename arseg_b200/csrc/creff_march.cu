// creff_march.cu -- fused MV-warp + CReFF + classifier, column-marching tensor-core engine (sm_100a).
//
// Contract and arithmetic (reference: evaluation.py:177-183 MV rescale + warpFeature,
// model/attention.py:184-213 MyAttention.forward, model/pspnet.py:226-229 final_conv + LogSoftmax,
// evaluation.py:204 argmax), C = 64, NHWC operands.  What changes is the decomposition:
//
//   * a CTA owns one 16-pixel-wide column strip of one frame and MARCHES down it 4 rows per step.  K, V live in a
//     shared-memory ring of image rows (f16, 128 B per position), so every warped-hr row is gathered once and
//     every K/V row is convolved once per strip: the only redundancy left is the horizontal halo
//     ((16+k+1)/16 gather, (16+k-1)/16 depthwise) instead of the (16+k+1)^2/256 of a square tile.
//   * three warp-specialised roles run CONCURRENTLY on different steps of the march, handing rows over through
//     mbarriers (4-deep, indexed by step):
//       G (6 warps): bilinear gather of the MV-warped hr rows and the lr_up rows into fp32 row rings
//                    (half-warp per position, float4 per lane; f64 MV arithmetic one position per thread);
//       D (6 warps): the three depthwise 3x3 convolutions (FFMA2, x-marching register window): K,V -> f16
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
constexpr int XG_WARPS = 6, XD_WARPS = 6, XC_WARPS = 4;                 // warps 0-5 / 6-11 / 12-15
static_assert((XG_WARPS + XD_WARPS) % 4 == 0 && XC_WARPS % 4 == 0, "setmaxnreg works on whole warpgroups");
constexpr int XG_THREADS = 32 * XG_WARPS, XD_THREADS = 32 * XD_WARPS, XC_THREADS = 32 * XC_WARPS;
constexpr int XNHW = XG_THREADS / 16;   // gather half-warps: one position each per slot
constexpr int XHR_RING = 10, XLR_RING = 10;   // fp32 row rings (rows)
constexpr int XJA = 4;                  // gather positions in flight per half-warp (ring of single-position slots)
constexpr int XRES_LD = 72;             // floats per residual row (bank-conflict pad)
constexpr int XCLS_LD = 72;             // f16 per classifier-weight row
constexpr int XNB = 4;                  // mbarriers per hand-off (indexed by step & 3)
constexpr int XBAR_G = 1;               // named barrier of the G group
constexpr int XQG = 2;                  // QK n-tiles interleaved (the legacy tensor pipe takes ~17 clk per MMA per SMSP)

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
    static constexpr int PREAL = (G0 * HC > 4 * HC + 4 * LC) ? G0 * HC : 4 * HC + 4 * LC;  // positions per G step
    static constexpr int PMAX = ((PREAL + XNHW * XJA - 1) / (XNHW * XJA)) * (XNHW * XJA);  // padded with no-op records
    static constexpr size_t KV_BYTES = (size_t)KVR * KVC * 128;
    static constexpr size_t HR_BYTES = (size_t)XHR_RING * HC * 256;
    static constexpr size_t LR_BYTES = (size_t)XLR_RING * LC * 256;
    static constexpr size_t Q_BYTES = 64 * 128;
    static constexpr size_t RES_BYTES = 64 * XRES_LD * 4;
    static constexpr size_t POS_BYTES = (size_t)PMAX * 32;
    static constexpr size_t CLS_BYTES = 32 * XCLS_LD * 2 + 32 * 4;
    static constexpr size_t SCRATCH_BYTES = 256;                  // ring-like slot the no-op gather records write to
    static constexpr size_t DWW_BYTES = 3 * 10 * 32 * 8;          // depthwise weights + bias, float2 per lane
    static_assert(CLS_BYTES <= Q_BYTES, "classifier staging aliases the Q tile");
    static constexpr size_t SMEM = 2 * KV_BYTES + HR_BYTES + LR_BYTES + SCRATCH_BYTES + Q_BYTES + RES_BYTES + POS_BYTES + DWW_BYTES + 4 * XNB * 8;
    static_assert(PMAX <= XG_THREADS, "one position record per G thread");
    static_assert(4 + 2 * R <= KVR, "K/V ring too small for one consumer step");
    static_assert(SMEM <= 232448, "shared memory budget");
};

// ---------------------------------------------------------------------------------------------
// mbarrier helpers (hand-off between the roles; index = march step + 1)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void xbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count));
}
// one arrival per WARP (barrier counts = warps of the producing role): a waiter parked in try_wait is woken by every
// arrival on its barrier, so per-thread arrivals make every waiting warp spin through 32x more wake-ups.  __syncwarp
// orders the lanes' shared-memory accesses before the elected lane's release-arrive.
// -DARSEG_ARRIVE_ALL (compute-sanitizer racecheck builds): every lane arrives (barrier counts x 32).  racecheck orders only the
// accesses of threads that arrive themselves; it does not follow the __syncwarp -> elected-lane release-arrive chain.
#ifdef ARSEG_ARRIVE_ALL
constexpr uint32_t XARRIVALS = 32;
#else
constexpr uint32_t XARRIVALS = 1;
#endif
__device__ __forceinline__ void xbar_arrive(uint64_t* bars, int step) {
#ifdef ARSEG_ARRIVE_ALL
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(bars + ((step + 1) & (XNB - 1)))) : "memory");
#else
    __syncwarp();
    if ((threadIdx.x & 31) == 0)
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(bars + ((step + 1) & (XNB - 1)))) : "memory");
#endif
}
__device__ __forceinline__ void xbar_wait(uint64_t* bars, int step) {
    const uint32_t addr = s_u32(bars + ((step + 1) & (XNB - 1)));
    const uint32_t parity = (uint32_t)(((step + 1) / XNB) & 1);
    uint32_t ok;
    // try_wait suspends the thread in hardware until the phase completes or a time limit expires; the poll loop is
    // kept minimal (a spinning warp takes issue slots from the working roles) and bounded (a protocol bug must trap,
    // not hang the GPU box)
#pragma unroll 1
    for (int spin = 0; spin < (1 << 22); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(addr), "r"(parity), "r"(100000u) : "memory");
        if (ok) return;
    }
    printf("arseg creff_march: mbarrier wait timed out (block %d thread %d step %d)\n", (int)blockIdx.x, (int)threadIdx.x, step);
    __trap();
}

// byte offset of 16-byte chunk `chunk` of K/V ring position `pos` = slot * KVC + col.  Swizzle key = col + WN * kr, kr =
// the K/V row index (image row - (ya - R)): an ldmatrix 8x8 reads 8 consecutive keys of a WN-wide key patch, i.e. the
// tail of one patch row and the head of the next; with the row term the head continues the tail's residues mod 8, so
// the 8 rows always land in 8 different 16-byte bank groups (with the column alone, columns 8, 9 aliased 0, 1: +45 %
// wavefronts on every K / V fragment load).  WN = k + 3 is even, so the key is the same for rows 4 steps apart: per
// lane it is constant over the march on both the writing (D) and the reading (C) side.
__device__ __forceinline__ uint32_t kv_off(int pos, int key, int chunk) { return (uint32_t)(pos * 128 + (((chunk ^ key) & 7) << 4)); }
// Q tile [row 0..3][col 0..15]: key = (col & 3) | (row & 1) << 2 -> the 8 rows of every ldmatrix 8x8 are distinct
__device__ __forceinline__ uint32_t q_off(int row, int col, int chunk) {
    return (uint32_t)((row * XSW + col) * 128 + (((chunk ^ ((col & 3) | ((row & 1) << 2))) & 7) << 4));
}

// Optional timeline trace (compile with -DARSEG_XTRACE): lane 0 of the first warp of every role of CTA XTRACE_CTA
// records (role, step, tag, clock64) tuples; read back with arseg_debug_creff_trace().
#ifdef ARSEG_XTRACE
constexpr int XTRACE_STEPS = 256, XTRACE_TAGS = 8, XTRACE_N = 3 * XTRACE_STEPS * XTRACE_TAGS;
__device__ long long g_xtrace[XTRACE_N];
__device__ __forceinline__ void xtrace(int role, int step, int tag) {
    if (blockIdx.x == 148 * 2 + 7 && (threadIdx.x == 0 || threadIdx.x == 32 * XG_WARPS || threadIdx.x == 384) && step + 1 < XTRACE_STEPS)
        g_xtrace[(role * XTRACE_STEPS + step + 1) * XTRACE_TAGS + tag] = clock64();
}
#define XTRACE(role, step, tag) xtrace(role, step, tag)
#else
#define XTRACE(role, step, tag)
#endif

struct XSmem {
    uint8_t *sK, *sV, *rings, *sQ;
    float* sRes;
    float4* posw;
    uint4* posa;
    __half* s_wc;
    float* s_bc;
    float2* s_dw;     // [3 convs: k, v, q][9 taps + bias][32 lanes] (channels 2*lane, 2*lane+1)
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

// A gather record names a 2x2 source block that lies INSIDE the image (top-left pixel clamped to [0, W-2] x [0, H-2]) plus
// the four weights of its pixels, so the taps are base, base + one pixel, base + one row, base + one row + one pixel:
// three of the four addresses are immediates / one add.  PosRec (pos_hr / pos_lr) names the taps by a clamped NW
// tap and dx / dy flags instead; where a flag is 0 both taps of that direction read the same pixel (image border), and
// their weights are merged onto whichever block column / row holds it (at most one of the two is non-zero for the
// warp, grid_sample zero padding; for lr_up the two add up to the same bilinear weight).
__device__ __forceinline__ void x_block_of(const PosRec& r, int Wimg, int Himg, float4& w, int& bx, int& by) {
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

__device__ __forceinline__ void sts_f4(uint32_t a, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int K, typename TLR, bool MVF>
__device__ __forceinline__ void x_g_role(const CreffMmaParams& p, const XSmem& sm, int n, int x0, int ya, int S) {
    using Cf = XCfg<K>;
    constexpr int NHW = XNHW;                             // half-warps: one gather position each
    constexpr int LR_ES = (int)sizeof(TLR);
    constexpr uint32_t SCRATCH_OFF = (uint32_t)(Cf::HR_BYTES + Cf::LR_BYTES);
    const int gt = threadIdx.x, lane = gt & 31, hw = gt >> 4, cl = lane & 15;
    const float lsh = resize_scale(p.h, p.H, ARSEG_RESIZE_BILINEAR_AC), lsw = resize_scale(p.w, p.W, ARSEG_RESIZE_BILINEAR_AC);
    const char* const hrb = reinterpret_cast<const char*>(p.hr + (p.hr_shared ? 0 : (size_t)n * p.H * p.W * MC));
    const char* const lrb = reinterpret_cast<const char*>(reinterpret_cast<const TLR*>(p.lr) + (size_t)n * p.h * p.w * MC);
    const double rcp_w = 2.0 / (double)max(p.W - 1, 1), rcp_h = 2.0 / (double)max(p.H - 1, 1);
    const uint32_t hr_rs = (uint32_t)p.W * MC * 4, lr_rs = (uint32_t)p.w * MC * LR_ES;
    const uint32_t ring_st = s_u32(sm.rings) + 16 * cl;   // this lane's 4 channels of ring position 0

    // position record q of a step (one per G thread): posw = the block's four weights; posa = {block address (64 bit),
    // row stride in bytes, byte offset of the destination ring position | lr << 31}.  The record list is padded to a
    // multiple of NHW * XJA with no-op records (zero weights, scratch destination), so the gather loop is branch-free.
    // Positions outside the image are zero-weight records onto a valid address (exact zeros for finite inputs).
    // Single-buffered: the records of step t+1 are written after the last use of those of step t.
    // MVF (compile time): the MV field is the on-disk int16 quarter-pel map at feature resolution -- the generic MV code
    // (f32 / f64 fields, f64 bilinear resize of the field, evaluation.py:177-180) is then not even instantiated, which keeps
    // the once-per-step record code short (it runs cold in the instruction cache every step)
    constexpr bool mv_fast = MVF;
    const int* const mvp = reinterpret_cast<const int*>(p.flow) + (size_t)n * p.H * p.W;
    auto padded = [](int npos) { return ((npos + NHW * XJA - 1) / (NHW * XJA)) * (NHW * XJA); };
    // the int16 MV pair of this thread's hr position of step t (loaded one step ahead of compute_pos)
    auto mv_of = [&](int t) -> int {
        int h0, nh, l0, nl;
        x_step_geom<K>(t, h0, nh, l0, nl);
        if (!mv_fast || gt >= nh * Cf::HC) return 0;
        const int rr = gt / Cf::HC, cc = gt - rr * Cf::HC;
        const int fy = ya - Cf::R - 1 + h0 + rr, fx = x0 - Cf::R - 1 + cc;
        return (fy >= 0 && fy < p.H && fx >= 0 && fx < p.W) ? __ldg(mvp + (size_t)fy * p.W + fx) : 0;
    };
    auto compute_pos = [&](int t, int mv) {
        int h0, nh, l0, nl;
        x_step_geom<K>(t, h0, nh, l0, nl);
        const int nhp = nh * Cf::HC, npos = nhp + nl * Cf::LC;
        const int q = gt;
        if (q >= padded(npos)) return;
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        const char* src = hrb;
        uint32_t rs = hr_rs, dst = SCRATCH_OFF;
        if (q < nhp) {
            const int rr = q / Cf::HC, cc = q - rr * Cf::HC, row = h0 + rr;
            const int fy = ya - Cf::R - 1 + row, fx = x0 - Cf::R - 1 + cc;
            const PosRec r = pos_hr(p, n, fy, fx, rcp_w, rcp_h, mv_fast ? &mv : nullptr);
            dst = (uint32_t)(((row % XHR_RING) * Cf::HC + cc) * 256);
            if (r.info >= 0) {
                int bx, by;
                x_block_of(r, p.W, p.H, w, bx, by);
                src = hrb + ((size_t)by * p.W + bx) * (MC * 4);
            }
        } else if (q < npos) {
            const int q2 = q - nhp, rr = q2 / Cf::LC, cc = q2 - rr * Cf::LC, row = l0 + rr;
            const PosRec r = pos_lr(p, lsh, lsw, ya - 1 + row, x0 - 1 + cc);
            dst = (uint32_t)(Cf::HR_BYTES + ((row % XLR_RING) * Cf::LC + cc) * 256);
            if (r.info >= 0) {
                int bx, by;
                x_block_of(r, p.w, p.h, w, bx, by);
                src = lrb + ((size_t)by * p.w + bx) * (MC * LR_ES);
                rs = lr_rs;
                dst |= 0x80000000u;
            }
        }
        sm.posw[q] = w;
        const unsigned long long a = reinterpret_cast<unsigned long long>(src);
        sm.posa[q] = make_uint4((uint32_t)a, (uint32_t)(a >> 32), rs, dst);
    };
    float4 tap[XJA][4];
    auto issue = [&](float4 (&tp)[4], int j) {
#ifdef ARSEG_XTRACE
        if (p.dbg & 4) return;
#endif
        const uint4 id = sm.posa[hw + NHW * j];
        const char* a0 = reinterpret_cast<const char*>(((unsigned long long)id.y << 32) | id.x);
        if (LR_ES == 4 || !(id.w >> 31)) {
            a0 += 16 * cl;
            const char* a1 = a0 + id.z;
            tp[0] = ld4(reinterpret_cast<const float*>(a0));
            tp[1] = ld4(reinterpret_cast<const float*>(a0 + MC * 4));
            tp[2] = ld4(reinterpret_cast<const float*>(a1));
            tp[3] = ld4(reinterpret_cast<const float*>(a1 + MC * 4));
        } else {
            a0 += LR_ES * 4 * cl;
            const char* a1 = a0 + id.z;
            tp[0] = ld4(reinterpret_cast<const TLR*>(a0));
            tp[1] = ld4(reinterpret_cast<const TLR*>(a0 + MC * LR_ES));
            tp[2] = ld4(reinterpret_cast<const TLR*>(a1));
            tp[3] = ld4(reinterpret_cast<const TLR*>(a1 + MC * LR_ES));
        }
    };
    auto commit = [&](const float4 (&tp)[4], int j) {
        const int i = hw + NHW * j;
        const float4 w = sm.posw[i];
        const uint32_t dst = sm.posa[i].w & 0x7fffffffu;
        float4 v;
        v.x = tp[0].x * w.x + tp[1].x * w.y + tp[2].x * w.z + tp[3].x * w.w;
        v.y = tp[0].y * w.x + tp[1].y * w.y + tp[2].y * w.z + tp[3].y * w.w;
        v.z = tp[0].z * w.x + tp[1].z * w.y + tp[2].z * w.z + tp[3].z * w.w;
        v.w = tp[0].w * w.x + tp[1].w * w.y + tp[2].w * w.z + tp[3].w * w.w;
        sts_f4(ring_st + dst, v);
    };

    compute_pos(-1, mv_of(-1));
    nbar_sync(XBAR_G, XG_THREADS);
#pragma unroll 1
    for (int t = -1; t <= S; ++t) {
        int h0, nh, l0, nl;
        x_step_geom<K>(t, h0, nh, l0, nl);
        const int nj = padded(nh * Cf::HC + nl * Cf::LC) / NHW;   // positions of this half-warp (multiple of XJA)
        XTRACE(0, t, 0);
        // rolling pipeline: XJA positions' loads are always in flight while the oldest one is combined and stored
#pragma unroll
        for (int j = 0; j < XJA; ++j) issue(tap[j], j);
        const int mv_next = t < S ? mv_of(t + 1) : 0;     // in flight during the gather loop
        if (t >= 1) xbar_wait(sm.ddone, t - 2);           // D step t-2 done: the ring rows this step overwrites are free
        XTRACE(0, t, 1);
#pragma unroll 1
        for (int j0 = 0; j0 < nj; j0 += XJA) {
            if (j0 + XJA < nj) {
#pragma unroll
                for (int j = 0; j < XJA; ++j) { commit(tap[j], j0 + j); issue(tap[j], j0 + j + XJA); }
            } else {
#pragma unroll
                for (int j = 0; j < XJA; ++j) commit(tap[j], j0 + j);
            }
        }
        XTRACE(0, t, 2);
        xbar_arrive(sm.gfull, t);
        nbar_sync(XBAR_G, XG_THREADS);                    // every G thread is done with the records of step t
        if (t < S) compute_pos(t + 1, mv_next);
        nbar_sync(XBAR_G, XG_THREADS);                    // records of step t+1 visible to the group
        XTRACE(0, t, 3);
    }
}

// ---------------------------------------------------------------------------------------------
// D role: depthwise 3x3 convolutions.  Warp d = 3 * rp + third owns column third `third` of rows rp and rp + 2 of
// every 4-row strip and convolves the two rows in one pass (the two rows share input row 2 of the five they read).
// Everything is addressed with 32-bit shared-window addresses whose per-column parts are compile-time immediates:
// the column loop is fully unrolled, the 3-column register window rotates by renaming.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 lds_f2(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_f2(uint32_t a, float2 v) { asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(a), "f"(v.x), "f"(v.y) : "memory"); }

// one 3x3 depthwise output (two channels): three independent row chains (shorter dependency chains than one 9-deep chain)
__device__ __forceinline__ float2 x_dw9(const float2 (&w)[10], const float2 (&r0)[3], const float2 (&r1)[3], const float2 (&r2)[3],
                                        int sa, int sb, int sc) {
    float2 a0 = __ffma2_rn(w[0], r0[sa], w[9]), a1 = __fmul2_rn(w[3], r1[sa]), a2 = __fmul2_rn(w[6], r2[sa]);
    a0 = __ffma2_rn(w[1], r0[sb], a0); a1 = __ffma2_rn(w[4], r1[sb], a1); a2 = __ffma2_rn(w[7], r2[sb], a2);
    a0 = __ffma2_rn(w[2], r0[sc], a0); a1 = __ffma2_rn(w[5], r1[sc], a1); a2 = __ffma2_rn(w[8], r2[sc], a2);
    return __fadd2_rn(__fadd2_rn(a0, a1), a2);
}

// Two output rows A (input rows 0..2) and B (input rows 2..4) x NC columns; ra[i] = shared address of this lane's channel
// pair at input column 0 of input row i.  Columns >= NMIN are computed only when `full` (warp-uniform).
// emit(x, a1, a2, b1, b2, centreA, centreB): conv 1 / conv 2 results of rows A / B at output column x (a compile-time
// constant after unrolling), centre = the input sample under the kernel centre.
template <int NOUT, int NC, int NMIN, typename Emit>
__device__ __forceinline__ void x_dw_rows5(const uint32_t (&ra)[5], bool full, const float2 (&w1)[10], const float2 (&w2)[10], Emit&& emit) {
    float2 win[5][3];   // [input row][slot]; slot (x + d) % 3 holds input column x + d
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < 5; ++i) win[i][c] = lds_f2(ra[i] + c * (MC * 4));
#pragma unroll
    for (int x = 0; x < NC; ++x) {
        if (x < NMIN || full) {
            const int sa = x % 3, sb = (x + 1) % 3, sc = (x + 2) % 3;
#pragma unroll
            for (int i = 0; i < 5; ++i) win[i][sc] = lds_f2(ra[i] + (x + 2) * (MC * 4));
            const float2 a1 = x_dw9(w1, win[0], win[1], win[2], sa, sb, sc), b1 = x_dw9(w1, win[2], win[3], win[4], sa, sb, sc);
            float2 a2 = make_float2(0.f, 0.f), b2 = a2;
            if (NOUT == 2) { a2 = x_dw9(w2, win[0], win[1], win[2], sa, sb, sc); b2 = x_dw9(w2, win[2], win[3], win[4], sa, sb, sc); }
            emit(x, a1, a2, b1, b2, win[1][sb], win[3][sb]);
        }
    }
}

template <int K>
__device__ __forceinline__ void x_d_role(const CreffMmaParams& p, const XSmem& sm, int x0, int ya, int S) {
    using Cf = XCfg<K>;
    const int lane = threadIdx.x & 31, d = (threadIdx.x >> 5) - XG_WARPS, third = d % 3, rp = d / 3;
    constexpr int KV3 = (Cf::KVC + 2) / 3, Q3 = (XSW + 2) / 3;              // columns per third
    constexpr int KVL = Cf::KVC - 2 * KV3, QL = XSW - 2 * Q3;               // columns of the last third
    static_assert(KVL > 0 && KVL <= KV3 && QL > 0 && QL <= Q3, "column thirds");
    const int kv_lo = third * KV3, q_lo = third * Q3;
    const bool full = third != 2;
    const int lane_chunk = lane >> 2;
    const uint32_t lane_off = (uint32_t)((lane & 3) * 4);
    const float2* const swl = sm.s_dw + lane;
    // shared-window addresses: this lane's channel pair at the first input column of its third, ring row 0
    const uint32_t hring = s_u32(sm.rings) + (uint32_t)(kv_lo * (MC * 4) + lane * 8);
    const uint32_t lring = s_u32(sm.rings) + (uint32_t)(Cf::HR_BYTES + q_lo * (MC * 4) + lane * 8);
    const uint32_t kbase = s_u32(sm.sK), qbase = s_u32(sm.sQ);
    const uint32_t resbase = s_u32(sm.sRes) + (uint32_t)(((rp * XSW + q_lo) * XRES_LD + 2 * lane) * 4);
    // per-column store offsets inside the Q tile (the swizzle key depends on the column)
    uint32_t qst[Q3];
#pragma unroll
    for (int x = 0; x < Q3; ++x) qst[x] = q_off(rp, q_lo + x, lane_chunk) + lane_off;      // row rp + 2: + 2 * XSW * 128 (same key)
    // columns of this third inside the image (constant over the march)
    const int fx_lo = x0 - Cf::R + kv_lo, kv_n = full ? KV3 : KVL;
    const bool cols_ok = fx_lo >= 0 && fx_lo + kv_n <= p.W;
    // ring slots of the first row of a step (advance by 4 per step)
    int hs = rp, ks = rp, ls = rp;          // hr ring slot of K/V row krA, K/V ring slot of krA, lr ring slot of Q row qr
#pragma unroll 1
    for (int t = -1; t <= S; ++t) {
        XTRACE(1, t, 0);
        xbar_wait(sm.gfull, t);
        XTRACE(1, t, 1);
        // ---- K/V rows kr, kr+2 (relative to image row ya-R) from hr ring rows kr..kr+2 / kr+2..kr+4 ----
        if (t - Cf::SL >= 0) xbar_wait(sm.cdone, t - Cf::SL);     // the ring rows this step overwrites have been read
        XTRACE(1, t, 2);
#ifdef ARSEG_XTRACE
        if (!(p.dbg & 1))
#endif
        if (t >= 0 || rp < Cf::P0) {
            const int krA = (t < 0 ? 0 : Cf::P0 + 4 * t) + rp;
            const bool haveB = t >= 0 || rp + 2 < Cf::P0;          // initial step: only P0 rows exist
            float2 wk[10], wv[10];
#pragma unroll
            for (int i = 0; i < 10; ++i) { wk[i] = swl[i * 32]; wv[i] = swl[(10 + i) * 32]; }
            uint32_t ra[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                int slot = hs + i;
                slot = slot >= XHR_RING ? slot - XHR_RING : slot;
                ra[i] = hring + (uint32_t)(slot * (Cf::HC * MC * 4));
            }
            int ksB = ks + 2;
            ksB = ksB >= Cf::KVR ? ksB - Cf::KVR : ksB;
            const uint32_t kA = kbase + (uint32_t)(ks * (Cf::KVC * 128)) + lane_off, kB = kbase + (uint32_t)(ksB * (Cf::KVC * 128)) + lane_off;
            const int keyA = kv_lo + Cf::WN * krA, keyB = keyA + 2 * Cf::WN;            // swizzle keys of column kv_lo in rows A / B
            x_dw_rows5<2, KV3, KVL>(ra, full, wk, wv, [&](int x, float2 a1, float2 a2, float2 b1, float2 b2, float2, float2) {
                const uint32_t oA = kA + kv_off(kv_lo + x, keyA + x, lane_chunk), oB = kB + kv_off(kv_lo + x, keyB + x, lane_chunk);
                sts_u32(oA, pack_h2_sat(a1.x, a1.y));
                sts_u32(oA + (uint32_t)Cf::KV_BYTES, pack_h2_sat(a2.x, a2.y));
                if (haveB) {
                    sts_u32(oB, pack_h2_sat(b1.x, b1.y));
                    sts_u32(oB + (uint32_t)Cf::KV_BYTES, pack_h2_sat(b2.x, b2.y));
                }
            });
            // K / V are exactly 0 outside the image (attention zero padding, model/attention.py:199,207): border strips and
            // the first / last rows overwrite what the convolution (bias included) produced there
            const int fyA = ya - Cf::R + krA;
            const bool okA = fyA >= 0 && fyA < p.H, okB = fyA + 2 >= 0 && fyA + 2 < p.H;
            if (!(cols_ok && okA && (okB || !haveB))) {
#pragma unroll 1
                for (int x = 0; x < kv_n; ++x) {
                    const bool okx = fx_lo + x >= 0 && fx_lo + x < p.W;
                    const uint32_t oA = kA + kv_off(kv_lo + x, keyA + x, lane_chunk), oB = kB + kv_off(kv_lo + x, keyB + x, lane_chunk);
                    if (!(okA && okx)) { sts_u32(oA, 0u); sts_u32(oA + (uint32_t)Cf::KV_BYTES, 0u); }
                    if (haveB && !(okB && okx)) { sts_u32(oB, 0u); sts_u32(oB + (uint32_t)Cf::KV_BYTES, 0u); }
                }
            }
        }
        if (t >= 0) {
            hs += 4; hs = hs >= XHR_RING ? hs - XHR_RING : hs;
            ks += 4; ks = ks >= Cf::KVR ? ks - Cf::KVR : ks;
        } else {
            hs = (Cf::P0 + rp) % XHR_RING; ks = (Cf::P0 + rp) % Cf::KVR;
        }
        // ---- Q rows qr, qr+2 (qr = 4(t-1)+rp, relative to ya) from the lr ring; residual = lr_up centre ----
        XTRACE(1, t, 3);
        if (t >= 1) {
            xbar_wait(sm.qlempty, t - 2);                         // C step t-2 (-1: prologue) has taken Q / residual into registers
            XTRACE(1, t, 4);
#ifdef ARSEG_XTRACE
            if (!(p.dbg & 1))
#endif
            {
                float2 wq[10];
#pragma unroll
                for (int i = 0; i < 10; ++i) wq[i] = swl[(20 + i) * 32];
                uint32_t ra[5];
#pragma unroll
                for (int i = 0; i < 5; ++i) {
                    int slot = ls + i;
                    slot = slot >= XLR_RING ? slot - XLR_RING : slot;
                    ra[i] = lring + (uint32_t)(slot * (Cf::LC * MC * 4));
                }
                x_dw_rows5<1, Q3, QL>(ra, full, wq, wq, [&](int x, float2 a1, float2, float2 b1, float2, float2 ca, float2 cb) {
                    sts_u32(qbase + qst[x], pack_h2_sat(a1.x, a1.y));
                    sts_u32(qbase + qst[x] + 2 * XSW * 128, pack_h2_sat(b1.x, b1.y));
                    sts_f2(resbase + (uint32_t)(x * XRES_LD * 4), ca);
                    sts_f2(resbase + (uint32_t)((2 * XSW + x) * XRES_LD * 4), cb);
                });
            }
            ls += 4; ls = ls >= XLR_RING ? ls - XLR_RING : ls;
        }
        xbar_arrive(sm.ddone, t);
        XTRACE(1, t, 5);
    }
}

// ---------------------------------------------------------------------------------------------
// C role: attention + classifier for one 4x4 block per warp per step.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int K, int NCT>
__device__ __forceinline__ void x_c_role(const CreffMmaParams& p, const XSmem& sm, int n, int x0, int ya, int yb, int S) {
    using Cf = XCfg<K>;
    constexpr int NT8Q = (Cf::NK + 7) / 8;                // key n-tiles that hold at least one real key
    constexpr int MW = (2 * NT8Q + 31) / 32;              // 32-bit mask words per query row
    const int lane = threadIdx.x & 31, cw = (threadIdx.x >> 5) - 12;
    const int g = lane >> 2, t = lane & 3, mi = lane >> 3;
    const int W = p.W;
    const size_t plane = (size_t)p.H * W;

    // validity of this thread's logits: rows g (mA) and g+8 (mB), keys 8j+2t+e -> bit 2j+e.  An invalid logit
    // (outside the k x k window of its query, or a padding key) starts its accumulator at -inf instead of 0, so
    // the MMA result is already masked and the softmax needs no selects (ex2(-inf) = 0).
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
    // ldmatrix row addresses per n-tile (QK: 8 keys per tile, lanes 0-7 of each address group; PV: 16 keys per tile).
    // Ring byte offsets of this lane's ldmatrix rows, one register per n-tile: (ring slot of the key row) * ROWB + the
    // column part.  The column part is < ROWB, so the slot is offset / ROWB and the per-step update is "add four rows,
    // wrap at the ring size" -- no per-step row table, no selects.
    constexpr uint32_t ROWB = Cf::KVC * 128, RINGB = Cf::KVR * ROWB;
    uint32_t aq[NT8Q], av[Cf::NT16];
#pragma unroll
    for (int j = 0; j < NT8Q; ++j) {
        int nk = 8 * j + (lane & 7);
        nk = nk < Cf::NK ? nk : Cf::NK - 1;
        const int ky = nk / Cf::WN, kx = nk - ky * Cf::WN, col = 4 * cw + kx;
        aq[j] = (uint32_t)ky * ROWB + kv_off(col, col + Cf::WN * ky, mi);                            // chunk mi; chunk 4+mi = ^ 64
    }
#pragma unroll
    for (int i = 0; i < Cf::NT16; ++i) {
        int nk = 16 * i + ((mi & 1) << 3) + (lane & 7);
        nk = nk < Cf::NK ? nk : Cf::NK - 1;
        const int ky = nk / Cf::WN, kx = nk - ky * Cf::WN, col = 4 * cw + kx;
        av[i] = (uint32_t)ky * ROWB + kv_off(col, col + Cf::WN * ky, mi >> 1);                       // chunk (mi>>1) + 2cp = ^ (cp << 5)
    }
    static_assert(Cf::WN <= Cf::KVR, "a key patch fits the ring");
    // classifier B fragments (final_conv weights, f16) and biases stay in registers for the whole march
    constexpr int NCTA = NCT > 0 ? NCT : 1;
    uint32_t wf[NCTA][4][2];
    float bcl[NCTA][2];
#pragma unroll
    for (int nt = 0; nt < NCT; ++nt) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const __half* wp = sm.s_wc + (8 * nt + g) * XCLS_LD + 16 * ks + 2 * t;
            wf[nt][ks][0] = *reinterpret_cast<const uint32_t*>(wp);
            wf[nt][ks][1] = *reinterpret_cast<const uint32_t*>(wp + 8);
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) bcl[nt][e] = 8 * nt + 2 * t + e < p.ncls ? sm.s_bc[8 * nt + 2 * t + e] : -INFINITY;   // padding classes: -inf
    }
    const uint32_t kb = s_u32(sm.sK), vb = s_u32(sm.sV);
    uint32_t qaddr[4];
    {
        const int r = ((mi & 1) << 3) + (lane & 7);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) qaddr[ks] = s_u32(sm.sQ) + q_off(r >> 2, 4 * cw + (r & 3), 2 * ks + (mi >> 1));
    }
    const float* const resa = sm.sRes + ((g >> 2) * XSW + 4 * cw + (g & 3)) * XRES_LD + 2 * t;
    const int pxA = x0 + 4 * cw + (g & 3);
    // output offsets of pixel A (block row g>>2) at step 0; pixel B is 2 rows below; each step advances 4 rows
    size_t offA = (size_t)(ya + (g >> 2)) * W + pxA;
    float* const ol = p.out_logits ? p.out_logits + (size_t)n * p.ncls * plane + (size_t)(2 * t) * plane : nullptr;
    float* const op = p.out_p ? p.out_p + (size_t)n * MC * plane + (size_t)(2 * t) * plane : nullptr;
    uint8_t* const oa = p.out_argmax ? p.out_argmax + (size_t)n * plane : nullptr;
    // the mbarrier phase of a step is (step+1)/XNB: C has no step -1, so arrive for it once (nobody waits on it)
    xbar_arrive(sm.cdone, -1);
    xbar_arrive(sm.qlempty, -1);
#pragma unroll 1
    for (int s = 0; s < S; ++s, offA += 4 * (size_t)W) {
        XTRACE(2, s, 0);
        xbar_wait(sm.ddone, s + 1);
        XTRACE(2, s, 1);
#ifdef ARSEG_XTRACE
        if (p.dbg & 2) { xbar_arrive(sm.qlempty, s); xbar_arrive(sm.cdone, s); continue; }
#endif
        // ---------------- S = Q K^T (model/attention.py:199) ----------------
        // n-tiles in groups of XQG with the k-step loop outside: XQG independent accumulator chains in flight
        float Sx[NT8Q][4];
        {
            uint32_t qa[4][4];
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) ldsm_x4(qa[ks], qaddr[ks]);
#pragma unroll
            for (int jg = 0; jg < NT8Q; jg += XQG) {
                uint32_t bf[XQG][2][4];
#pragma unroll
                for (int jj = 0; jj < XQG; ++jj) {
                    const int j = jg + jj;
                    if (j < NT8Q) {
                        const uint32_t a0 = kb + aq[j];
                        ldsm_x4(bf[jj][0], a0);
                        ldsm_x4(bf[jj][1], a0 ^ 64u);
                        const int b0 = 2 * j, b1 = 2 * j + 1;
                        Sx[j][0] = (mA[b0 >> 5] >> (b0 & 31)) & 1u ? 0.f : -INFINITY;
                        Sx[j][1] = (mA[b1 >> 5] >> (b1 & 31)) & 1u ? 0.f : -INFINITY;
                        Sx[j][2] = (mB[b0 >> 5] >> (b0 & 31)) & 1u ? 0.f : -INFINITY;
                        Sx[j][3] = (mB[b1 >> 5] >> (b1 & 31)) & 1u ? 0.f : -INFINITY;
                    }
                }
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
#pragma unroll
                    for (int jj = 0; jj < XQG; ++jj)
                        if (jg + jj < NT8Q) mma16816(Sx[jg + jj], qa[ks], bf[jj][ks >> 1][2 * (ks & 1)], bf[jj][ks >> 1][2 * (ks & 1) + 1]);
            }
        }
        // ---------------- residual lr_up (model/attention.py:191,210), kept in registers until the end ----------
        // thread (g,t): pixels A = block row g>>2, B = A + 2 rows; channels 8c+2t, 8c+2t+1
        float2 resA[8], resB[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            resA[c] = *reinterpret_cast<const float2*>(resa + 8 * c);
            resB[c] = *reinterpret_cast<const float2*>(resa + 2 * XSW * XRES_LD + 8 * c);
        }
        xbar_arrive(sm.qlempty, s);
        XTRACE(2, s, 2);
        // ---------------- softmax over the k*k window of every query (model/attention.py:203) ----------------
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
        // un-normalised P in f16 (A fragments of P V); the row sums come out of the P V MMAs (ones column below)
        uint32_t pa[Cf::NT16][4];
#pragma unroll
        for (int i = 0; i < Cf::NT16; ++i)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = 2 * i + h;
                if (j < NT8Q) {
                    pa[i][2 * h] = pack_h2(ex2_approx(fmaf(Sx[j][0], LOG2E, -o0)), ex2_approx(fmaf(Sx[j][1], LOG2E, -o0)));
                    pa[i][2 * h + 1] = pack_h2(ex2_approx(fmaf(Sx[j][2], LOG2E, -o1)), ex2_approx(fmaf(Sx[j][3], LOG2E, -o1)));
                } else {
                    pa[i][2 * h] = 0u; pa[i][2 * h + 1] = 0u;
                }
            }
        XTRACE(2, s, 3);
        // ---------------- O = P V (model/attention.py:207); row sums = P . 1 on the same tensor pipe ----------------
        float O[8][4], Ssum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 8; ++c) O[c][0] = O[c][1] = O[c][2] = O[c][3] = 0.f;
        {
            // V fragments one key tile ahead of the MMAs that consume them (ldmatrix latency off the MMA chain)
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
                mma16816(Ssum, pa[i], 0x3C003C00u, 0x3C003C00u);      // B = all ones (f16): every column = row sum of P
            }
        }
        xbar_arrive(sm.cdone, s);
        XTRACE(2, s, 4);
#pragma unroll
        for (int j = 0; j < NT8Q; ++j) { aq[j] += 4 * ROWB; aq[j] = aq[j] >= RINGB ? aq[j] - RINGB : aq[j]; }
#pragma unroll
        for (int i = 0; i < Cf::NT16; ++i) { av[i] += 4 * ROWB; av[i] = av[i] >= RINGB ? av[i] - RINGB : av[i]; }
        // fused = lr_up + O / sum (model/attention.py:210); the sum is that of the f16-rounded P the MMA consumed
        {
            const float inv0 = rcp_approx(Ssum[0]), inv1 = rcp_approx(Ssum[2]);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                O[c][0] = fmaf(O[c][0], inv0, resA[c].x); O[c][1] = fmaf(O[c][1], inv0, resA[c].y);
                O[c][2] = fmaf(O[c][2], inv1, resB[c].x); O[c][3] = fmaf(O[c][3], inv1, resB[c].y);
            }
        }
        const int pyA = ya + 4 * s + (g >> 2);
        const bool okA = pyA < yb && pxA < W, okB = pyA + 2 < yb && pxA < W;
        const size_t offB = offA + 2 * (size_t)W;
        if (op) {
#pragma unroll
            for (int c = 0; c < 8; ++c)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    float* o = op + (size_t)(8 * c + e) * plane;
                    if (okA) o[offA] = O[c][e];
                    if (okB) o[offB] = O[c][2 + e];
                }
        }
        XTRACE(2, s, 5);
        if (NCT == 0) continue;

        // ---------------- classifier (model/pspnet.py:226) as a [16 x 64] x [64 x 8*NCT] MMA ----------------
        float Lg[NCT > 0 ? NCT : 1][4];
        {
            uint32_t fa[4][4];
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                fa[ks][0] = pack_h2_sat(O[2 * ks][0], O[2 * ks][1]);
                fa[ks][1] = pack_h2_sat(O[2 * ks][2], O[2 * ks][3]);
                fa[ks][2] = pack_h2_sat(O[2 * ks + 1][0], O[2 * ks + 1][1]);
                fa[ks][3] = pack_h2_sat(O[2 * ks + 1][2], O[2 * ks + 1][3]);
            }
#pragma unroll
            for (int nt = 0; nt < NCT; ++nt) { Lg[nt][0] = Lg[nt][2] = bcl[nt][0]; Lg[nt][1] = Lg[nt][3] = bcl[nt][1]; }
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
#pragma unroll
                for (int nt = 0; nt < NCT; ++nt) mma16816(Lg[nt], fa[ks], wf[nt][ks][0], wf[nt][ks][1]);
        }
        // argmax (first maximum, like torch.argmax) and log-softmax per pixel: the values of one pixel live in a quad
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
                    lse0 += ex2_approx(fmaf(Lg[nt][e], LOG2E, -q0));          // padding classes: ex2(-inf) = 0
                    lse1 += ex2_approx(fmaf(Lg[nt][2 + e], LOG2E, -q1));
                }
            lse0 += __shfl_xor_sync(0xffffffffu, lse0, 1); lse0 += __shfl_xor_sync(0xffffffffu, lse0, 2);
            lse1 += __shfl_xor_sync(0xffffffffu, lse1, 1); lse1 += __shfl_xor_sync(0xffffffffu, lse1, 2);
            lse0 = __logf(lse0) + lmax0; lse1 = __logf(lse1) + lmax1;
        }
        XTRACE(2, s, 6);
        if (ol) {
#pragma unroll
            for (int nt = 0; nt < NCT; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    if (8 * nt + 2 * t + e < p.ncls) {
                        float* o = ol + (size_t)(8 * nt + e) * plane;
                        if (okA) o[offA] = Lg[nt][e] - lse0;
                        if (okB) o[offB] = Lg[nt][2 + e] - lse1;
                    }
                }
        }
        if (oa && t == 0) {
            if (okA) oa[offA] = (uint8_t)am0;
            if (okB) oa[offB] = (uint8_t)am1;
        }
        XTRACE(2, s, 7);
    }
}

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
template <int K, typename TLR, int NCT, bool MVF>
__global__ void __launch_bounds__(XTHREADS, 1) creff_march_kernel(CreffMmaParams p) {
    using Cf = XCfg<K>;
    extern __shared__ __align__(1024) uint8_t xsm[];
    XSmem sm;
    sm.sK = xsm;
    sm.sV = sm.sK + Cf::KV_BYTES;
    sm.rings = sm.sV + Cf::KV_BYTES;                                        // hr ring, then lr ring
    sm.sQ = sm.rings + Cf::HR_BYTES + Cf::LR_BYTES + Cf::SCRATCH_BYTES;
    sm.sRes = reinterpret_cast<float*>(sm.sQ + Cf::Q_BYTES);
    sm.posw = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(sm.sRes) + Cf::RES_BYTES);
    sm.posa = reinterpret_cast<uint4*>(sm.posw + Cf::PMAX);
    // classifier weights are staged in the Q tile: C moves them to registers in its prologue, and D's first Q write
    // (step 1) waits for C's prologue arrival on qlempty(-1)
    sm.s_wc = reinterpret_cast<__half*>(sm.sQ);                             // [32][XCLS_LD]
    sm.s_bc = reinterpret_cast<float*>(sm.s_wc + 32 * XCLS_LD);             // [32]
    sm.s_dw = reinterpret_cast<float2*>(sm.posa + Cf::PMAX);
    sm.gfull = reinterpret_cast<uint64_t*>(sm.s_dw + 3 * 10 * 32);
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
            xbar_init(sm.gfull + i, XG_WARPS * XARRIVALS);
            xbar_init(sm.ddone + i, XD_WARPS * XARRIVALS);
            xbar_init(sm.cdone + i, XC_WARPS * XARRIVALS);
            xbar_init(sm.qlempty + i, XC_WARPS * XARRIVALS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 3 * 10 * 32; i += XTHREADS) {
        const int cv = i / 320, tp = (i / 32) % 10, ln = i % 32;
        const float* w = cv == 0 ? p.wk : (cv == 1 ? p.wv : p.wq);
        const float* b = cv == 0 ? p.bk : (cv == 1 ? p.bv : p.bq);
        sm.s_dw[i] = tp < 9 ? make_float2(__ldg(w + (2 * ln) * 9 + tp), __ldg(w + (2 * ln + 1) * 9 + tp)) : make_float2(__ldg(b + 2 * ln), __ldg(b + 2 * ln + 1));
    }
    if (p.wcls) {
        for (int i = tid; i < 32 * MC; i += XTHREADS) {
            const int j = i / MC, c = i % MC;
            sm.s_wc[j * XCLS_LD + c] = __float2half_rn(j < p.ncls ? clamp_h(__ldg(p.wcls + (size_t)j * MC + c)) : 0.f);
        }
        if (tid < 32) sm.s_bc[tid] = (tid < p.ncls && p.bcls) ? __ldg(p.bcls + tid) : 0.f;
    }
    __syncthreads();

    // register file re-balance (warpgroup-aligned): the kernel starts with 128 per thread; G and D give 16 each to C
    if (warp < XG_WARPS + XD_WARPS) asm volatile("setmaxnreg.dec.sync.aligned.u32 112;");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 176;");
    if (warp < XG_WARPS) x_g_role<K, TLR, MVF>(p, sm, n, x0, ya, S);
    else if (warp < XG_WARPS + XD_WARPS) x_d_role<K>(p, sm, x0, ya, S);
    else x_c_role<K, NCT>(p, sm, n, x0, ya, yb, S);
}

template <int K, typename TLR, int NCT, bool MVF>
static int creff_march_launch_n(CreffMmaParams& p, cudaStream_t st) {
    using Cf = XCfg<K>;
    auto kern = creff_march_kernel<K, TLR, NCT, MVF>;
    static bool configured[64] = {false};
    int dev = 0;
    ARSEG_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        ARSEG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cf::SMEM));
        // keep what is left of the 228 KB as L1: the gather's memory-level parallelism is bounded by the L1 lines
        // its outstanding misses can allocate
        ARSEG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)((Cf::SMEM + 1024) * 100 / (228 * 1024)) + 1));
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    p.ncols = ceil_div(p.W, XSW);
    { const char* d = getenv("ARSEG_CREFF_DBG"); p.dbg = d ? atoi(d) : 0; }
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

// classifier n-tiles held in registers: none, <= 16 classes (CamVid 12), <= 32 classes (Cityscapes 19)
template <int K, typename TLR>
static int creff_march_launch_t(CreffMmaParams& p, cudaStream_t st) {
    const bool mvf = p.flow && p.flow_dtype == ARSEG_I16 && p.Hm == p.H && p.Wm == p.W;
    if (mvf) {
        if (!p.wcls) return creff_march_launch_n<K, TLR, 0, true>(p, st);
        if (p.ncls <= 16) return creff_march_launch_n<K, TLR, 2, true>(p, st);
        return creff_march_launch_n<K, TLR, 4, true>(p, st);
    }
    if (!p.wcls) return creff_march_launch_n<K, TLR, 0, false>(p, st);
    if (p.ncls <= 16) return creff_march_launch_n<K, TLR, 2, false>(p, st);
    return creff_march_launch_n<K, TLR, 4, false>(p, st);
}

template <int K>
static int creff_march_launch_k(CreffMmaParams& p, int lr_dtype, cudaStream_t st) {
    if (lr_dtype == ARSEG_BF16) return creff_march_launch_t<K, __nv_bfloat16>(p, st);
    if (lr_dtype == ARSEG_F16) return creff_march_launch_t<K, __half>(p, st);
    return creff_march_launch_t<K, float>(p, st);
}

int creff_march_launch(CreffMmaParams& p, int k, int lr_dtype, cudaStream_t st) {
    if (p.H < 2 || p.W < 2 || p.h < 2 || p.w < 2) ARSEG_UNSUPPORTED("creff_march: maps must be at least 2x2 (hr %dx%d, lr %dx%d)", p.H, p.W, p.h, p.w);
    switch (k) {
        case 3: return creff_march_launch_k<3>(p, lr_dtype, st);
        case 5: return creff_march_launch_k<5>(p, lr_dtype, st);
        case 7: return creff_march_launch_k<7>(p, lr_dtype, st);
        case 9: return creff_march_launch_k<9>(p, lr_dtype, st);
        default: ARSEG_UNSUPPORTED("creff_march: window k=%d", k);
    }
}

bool creff_mma_supported(const arseg_creff_args* a) {
    return a->C == MC && a->hr_layout == ARSEG_NHWC && a->lr_layout == ARSEG_NHWC && (a->k == 3 || a->k == 5 || a->k == 7 || a->k == 9) &&
           (!a->wcls || a->ncls <= 32) && ((size_t)a->H * a->W < (1u << 29)) && ((size_t)a->h * a->w < (1u << 29));
}

int creff_tc_launch(CreffMmaParams& p, int k, int hr_dtype, int phase, void* ws, size_t ws_bytes, cudaStream_t st);   // creff_tc.cu

// C = 64: ARSEG_CREFF_TCGEN05 (or ARSEG_CREFF_MMA_F16 with an f16 keyframe feature) -> the tcgen05 engine (creff_tc.cu, f16 LR
// feature, k <= 7); ARSEG_CREFF_MMA_F16 -> the column-marching mma.sync engine of this file (any LR dtype, k <= 9).
int creff_mma_launch(const arseg_creff_args* a, cudaStream_t st) {
    CreffMmaParams p;
    p.hr = reinterpret_cast<const float*>(a->hr); p.hr_shared = a->hr_shared; p.flow = a->flow; p.flow_dtype = a->flow_dtype; p.Hm = a->Hm; p.Wm = a->Wm;
    p.lr = a->lr; p.h = a->h; p.w = a->w;
    p.wq = a->wq; p.bq = a->bq; p.wk = a->wk; p.bk = a->bk; p.wv = a->wv; p.bv = a->bv; p.wcls = a->wcls; p.bcls = a->bcls;
    p.ncls = a->ncls; p.log_softmax = a->log_softmax; p.out_p = a->out_p; p.out_logits = a->out_logits;
    p.out_argmax = a->out_argmax; p.N = a->N; p.C = a->C; p.H = a->H; p.W = a->W;
    if (a->engine == ARSEG_CREFF_TCGEN05 || a->hr_dtype == ARSEG_F16) {
        if (a->lr_dtype != ARSEG_F16 || a->k > 7)
            ARSEG_UNSUPPORTED("creff: the tcgen05 engine needs an f16 LR feature and k <= 7 (lr dtype %d, k = %d)", a->lr_dtype, a->k);
        return creff_tc_launch(p, a->k, a->hr_dtype, a->phase, a->workspace, a->workspace_bytes, st);
    }
    return creff_march_launch(p, a->k, a->lr_dtype, st);
}

}  // namespace arseg
#ifdef ARSEG_XTRACE
extern "C" int arseg_debug_creff_trace(long long* host, int cap) {
    const int n = cap < arseg::XTRACE_N ? cap : arseg::XTRACE_N;
    cudaMemcpyFromSymbol(host, arseg::g_xtrace, sizeof(long long) * n);
    return n;
}
#endif
