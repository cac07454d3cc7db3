// creff_tc.cu -- fused MV-warp + CReFF + classifier on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a, C = 64.
//
// Same contract and arithmetic as creff_march.cu (reference: evaluation.py:177-183 MV rescale + warpFeature,
// model/attention.py:184-213 MyAttention.forward, model/pspnet.py:226-229 final_conv + LogSoftmax, evaluation.py:204
// argmax).  The window contractions S = Q K^T and O = P V and the classifier run as tcgen05.mma with the accumulators,
// P and the classifier input in tensor memory; nothing of the attention goes through ldmatrix / mma.sync any more.
//
//   * A CTA owns a 16-pixel-wide column strip of one frame and marches down it one TILE = 8 rows x 16 pixels = 128
//     queries at a time: query (qy, qx) is row qy*16+qx of the M = 128 UMMA tile, i.e. TMEM lane qy*16+qx.
//   * K and V live in shared-memory rings of image rows, every key one 128-byte row (64 x f16) in the canonical
//     SWIZZLE_128B layout, TKP = 24 keys per ring row (16 + k - 1 used), so consecutive ring rows are contiguous
//     8-key groups (SBO = 1024): ONE K-major B descriptor spans up to 8 key rows (N = 24 * rows), and the same bytes
//     read as an MN-major B operand are V[key][channel] for O = P V.  Layouts verified by tools/umma_probe.cu.
//   * S (128 lanes x NK fp32 columns, NK = 24 * (8 + k - 1)) is read back by four softmax warps (TMEM lane quarter =
//     warp index): every lane owns one query, masks the columns outside its own k x k window, and writes the
//     un-normalised P as packed f16 pairs over the S columns it has consumed; P is then the TMEM A operand of P V.
//     The row sums come from one more MMA against a ones tile, the residual lr_up + O / sum is formed in registers,
//     goes back to TMEM as f16 and is the A operand of the classifier MMA.
//   * Producer roles as in the march engine, handing rows over through mbarriers:
//       W (pre-pass kernel, creff_tc_warp_kernel): the MV warp of the keyframe feature (bilinear gather, f16, mixed-precision
//                    FHFMA) written ONCE per frame to a zero-bordered f16 NHWC workspace -- a latency-bound gather belongs in a
//                    full-occupancy kernel, not in five warps of a one-CTA-per-SM engine;
//       G (5 warps): one thread streams the warped rows into the f16 row ring with cp.async.bulk (3 KB per row, completion
//                    counted on the row hand-off mbarrier); the warps gather the lr_up rows (quarter-warp per position, 8
//                    channels = 16 bytes per lane);
//       D (6 warps): depthwise 3x3 convolutions (FFMA2): two K warps, two V warps (4 rows x half the columns each),
//                    two Q warps; K, V -> rings, Q -> the A tile;
//       M (1 warp):  MMA issue (one elected lane);
//       S (4 warps): softmax: S -> P in tensor memory;
//       E (4 warps): residual, classifier input, log-softmax / argmax, stores -- one tile behind S.
#include "creff_mma_common.cuh"
#include <cstdlib>

namespace arseg {

constexpr int TSW = 16;                    // strip width (pixels)
// 20 warps, no register re-balancing (640 threads x 96 registers).  Scheduler = warp id % 4 = TMEM lane quarter.  The depthwise
// warps are latency-bound (dependent FFMA2 chains: ncu shows them issuing a quarter of the time), so they get the most warps --
// two per scheduler -- and the latency-critical chain M -> S -> M -> E has the highest warp ids (the arbiter prefers them):
//   0 K0  1 K1  2 K2  3 V0 | 4 V1  5 V2  6 Qa  7 Qb | 8 G0  9 G1  10 G2  11 M | 12..15 E | 16..19 S
constexpr int TTHREADS = 640;
#ifndef ARSEG_TC_KVSPLIT
#define ARSEG_TC_KVSPLIT 3
#endif
constexpr int TKV_SPLIT = ARSEG_TC_KVSPLIT;    // column parts of a K / V row (3: the layout above; 2: K0 K1 V0 V1 Qa Qb + five G warps -- within 2 %)
#ifndef ARSEG_TC_QROWSPLIT
#define ARSEG_TC_QROWSPLIT 1
#endif
constexpr int TQ_RSPLIT = ARSEG_TC_QROWSPLIT;  // Q warps per column half: 1 = four query rows per warp and half-step, 2 = two rows each
constexpr int TQ_ROWS = 4 / TQ_RSPLIT;
constexpr int TQ_WARP0 = 2 * TKV_SPLIT;        // Q warps: two column halves x TQ_RSPLIT row groups
constexpr int TD_WARPS = TQ_WARP0 + 2 * TQ_RSPLIT;
static_assert(TD_WARPS <= 9, "at least two gather warps");
constexpr int TG_WARP0 = TD_WARPS, TG_WARPS = 11 - TD_WARPS;
constexpr int TM_WARP = 11;                // MMA issuer (and TMEM allocator)
constexpr int TE_W0 = 12, TS_W0 = 16, TC_WARPS = 4;   // epilogue warps 12..15, softmax warps 16..19 (lane quarter = warp % 4)
constexpr int TG_THREADS = 32 * TG_WARPS;
constexpr int TNQW = 4 * TG_WARPS;         // gather quarter-warps: one position each per slot
constexpr int TJA = 4;                     // gather positions in flight per quarter-warp
constexpr int TKP = 24;                    // keys per K/V ring row
constexpr int TROWB = TKP * 128;           // bytes per K/V ring row
constexpr int THRR = 10, TLRR = 16;        // warped-hr / lr_up ring rows
constexpr int TLC = TSW + 2;               // lr_up ring columns
constexpr int TNB = 4;                     // mbarriers per hand-off (indexed by step & 3)
constexpr int TBAR_G = 1;                  // named barrier of the G group
constexpr int TNCLS = 32;                  // classifier rows staged (>= ncls)
constexpr int TWB_DEFAULT = 8;             // workspace rows per band of the MV-warp pre-pass
// 128-thread CTAs: 0.340 ms against 0.358 ms with 256 threads (finer-grained tail); small enough to sit beside a resident conv CTA,
// which however did not change how much of the pre-pass hides under the LR branch (~0.1 ms either way)
constexpr int TWARP_THREADS = 128;

template <int K> struct TCfg {
    static constexpr int R = K / 2;
    static constexpr int NKR = 8 + 2 * R;                          // key rows of a tile
    static constexpr int PADR = (4 - (2 * R) % 4) % 4;             // rows produced before the first key row (4-row half-steps)
    static constexpr int HP = (2 * R + PADR) / 4;                  // half-steps before the first query row
    static constexpr int KVC = TSW + 2 * R;                        // K/V ring columns in use (<= TKP)
    static constexpr int HC = KVC + 2;                             // warped-hr ring columns
    // K/V ring rows: a tile's window + the 8 new rows of the NEXT tile, so that the depthwise warps never wait for P V of the
    // current tile before producing the next one (with NKR + 4 rows the second half-step of tile i+1 overwrote rows tile i
    // still reads: D, S = Q K^T, softmax and P V ran strictly one after the other, 14.7k cycles per tile instead of ~11k)
    static constexpr int KVR = NKR + 8;
    static constexpr int NK = NKR * TKP;                           // S columns
    static constexpr int NITEM = 4 * TLC;                          // lr_up gather positions of a step
    static constexpr int PT = R + PADR + 1, PB = 8 + R, PX = R + 1;   // zero border of the warped-keyframe workspace (rows above / below, columns left)
    static constexpr uint32_t HROW_BYTES = HC * 128;               // one warped-hr ring row = one bulk copy
    static constexpr int NJ = (NITEM + TNQW - 1) / TNQW;
    static constexpr int PMAX = NJ * TNQW;                         // padded with no-op records
    static constexpr int KVC3 = (KVC + TKV_SPLIT - 1) / TKV_SPLIT; // columns of a K / V warp (the last one takes what is left)
    static constexpr uint32_t COL_O = NK, COL_SUM = NK + 64, COL_A = NK + 80, COL_L = NK + 112;
    static constexpr size_t KV_BYTES = (size_t)KVR * TROWB;
    static constexpr size_t Q_BYTES = 128 * 128;
    static constexpr size_t W_BYTES = TNCLS * 128;
    static constexpr size_t ONES_BYTES = 2048;
    static constexpr size_t HR_BYTES = (size_t)THRR * HC * 128;
    static constexpr size_t LR_BYTES = (size_t)TLRR * TLC * 128;
    static constexpr size_t SCRATCH_BYTES = 128;
    static constexpr size_t REC_BYTES = (size_t)2 * PMAX * 16;     // double-buffered gather records (uint4)
    static constexpr size_t SMEM = 1024 + 2 * KV_BYTES + Q_BYTES + W_BYTES + ONES_BYTES + HR_BYTES + LR_BYTES + SCRATCH_BYTES + REC_BYTES +
                                   TNCLS * 4 + 10 * TNB * 8 + 16;
    static_assert(KVC <= TKP, "window too wide for the key pitch");
    static_assert(NKR % 2 == 0 && KVR % 2 == 0 && PADR % 2 == 0, "even row counts: a 16-key MMA step never straddles the ring end");
    static_assert(COL_L + TNCLS <= 512, "tensor memory columns");
    static_assert(PMAX <= 2 * TG_THREADS, "at most two gather records per G thread");
    static_assert(SMEM <= 232448, "shared memory budget");
};

// workspace = [H][W] lr_up gather records (16 bytes) | [N][Hp][Wp][64] f16 MV-warped keyframe feature with a zero border
// (Hp = PT + H + PB rows, Wp = 16 * ceil(W / 16) + HC - 16 columns: every ring row of every strip is one in-bounds 128-byte
// aligned run of HC pixels)
template <int K> __host__ __device__ constexpr int t_hp(int H) { return TCfg<K>::PT + H + TCfg<K>::PB; }
template <int K> __host__ __device__ constexpr int t_wp(int W) { return (W + TSW - 1) / TSW * TSW + TCfg<K>::HC - TSW; }
static size_t t_rec_bytes(int H, int W) { return ((size_t)H * W * sizeof(uint4) + 1023) / 1024 * 1024; }
size_t creff_tc_workspace_bytes(int N, int H, int W, int k) {
    const size_t px = k <= 3 ? (size_t)t_hp<3>(H) * t_wp<3>(W) : k <= 5 ? (size_t)t_hp<5>(H) * t_wp<5>(W) : (size_t)t_hp<7>(H) * t_wp<7>(W);
    return t_rec_bytes(H, W) + (size_t)N * px * 128;
}

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count));
}
// hand-off barriers are indexed by a step counter >= 0: slot idx & 3, phase parity (idx / 4) & 1
// -DARSEG_ARRIVE_ALL: see creff_march.cu (racecheck builds)
#ifdef ARSEG_ARRIVE_ALL
constexpr uint32_t TARRIVALS = 32;
#else
constexpr uint32_t TARRIVALS = 1;
#endif
__device__ __forceinline__ void tbar_arrive(uint64_t* bars, int idx) {
#ifdef ARSEG_ARRIVE_ALL
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(bars + (idx & (TNB - 1)))) : "memory");
#else
    __syncwarp();
    if ((threadIdx.x & 31) == 0)
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(bars + (idx & (TNB - 1)))) : "memory");
#endif
}
__device__ __forceinline__ bool tbar_test(uint64_t* bars, int idx) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(s_u32(bars + (idx & (TNB - 1)))), "r"((uint32_t)((idx / TNB) & 1)) : "memory");
    return ok != 0;
}
__device__ __noinline__ void tbar_timeout(int tag, int idx) {
    if ((threadIdx.x & 31) == 0)
        printf("arseg creff_tc: mbarrier wait timed out (block %d warp %d barrier %d index %d)\n", (int)blockIdx.x, (int)(threadIdx.x >> 5), tag, idx);
    __trap();
}
// Optional wait-time trace (compile with -DARSEG_TTRACE): lane 0 of every warp of CTA TTRACE_CTA accumulates the cycles it
// spends in tbar_wait per barrier tag, [15] = cycles of its whole role; read back with arseg_debug_creff_tc_trace().
#ifdef ARSEG_TTRACE
constexpr int TTRACE_CTA = 148 * 3 + 11;
__device__ long long g_ttrace[32 * 16];
__device__ long long g_tev[3 * 64 * 12];     // [M | S0 | E0][tile][event] clock64 stamps
__device__ __forceinline__ void tev(int role, int tile, int ev) {
    if (blockIdx.x == TTRACE_CTA && (threadIdx.x & 31) == 0 && tile < 64) g_tev[(role * 64 + tile) * 12 + ev] = clock64();
}
#define TEV(role, tile, ev) tev(role, tile, ev)
#else
#define TEV(role, tile, ev)
#endif
// (the SLEEP template flag is kept for the call sites: producer and consumer roles now wait the same way)
template <bool SLEEP = false>
__device__ __forceinline__ void tbar_wait(uint64_t* bars, int idx, int tag = 0) {
#ifdef ARSEG_TTRACE
    const long long t_in = clock64();
#endif
    const uint32_t addr = s_u32(bars + (idx & (TNB - 1)));
    const uint32_t parity = (uint32_t)((idx / TNB) & 1);
    uint32_t ok;
    // bounded: a protocol bug must trap, not hang the GPU box
    // try_wait with a suspend-time hint: the thread is parked in hardware until the phase completes or the hint (ns) expires, so a
    // waiting warp issues one instruction per time-out instead of a poll loop (nanosleep back-off loops were 45 % of all
    // executed instructions; 2.75 -> 2.72 ms)
#pragma unroll 1
    for (int spin = 0; spin < (1 << 16); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(addr), "r"(parity), "r"(20000u) : "memory");
        if (ok) {
#ifdef ARSEG_TTRACE
            if (blockIdx.x == TTRACE_CTA && (threadIdx.x & 31) == 0) g_ttrace[(threadIdx.x >> 5) * 16 + tag] += clock64() - t_in;
#endif
            return;
        }
    }
    tbar_timeout(tag, idx);
}
__device__ __forceinline__ void t_commit(uint64_t* bars, int idx) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(bars + (idx & (TNB - 1)))) : "memory");
}
__device__ __forceinline__ void t_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void t_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void t_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// SWIZZLE_128B shared-memory matrix descriptor, SBO = 1024 (8 rows of 128 bytes), version 1 (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t t_desc(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A / B f16, b_major @16 (1 = MN-major), N >> 3 @17, M >> 4 @24
__host__ __device__ constexpr uint32_t t_idesc(int N, int b_mn) {
    return (1u << 4) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void t_mma_ss(uint64_t da, uint64_t db, uint32_t td, uint32_t acc, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(td), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void t_mma_ts(uint32_t ta, uint64_t db, uint32_t td, uint32_t acc, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(td), "r"(ta), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void t_ld32(uint32_t ta, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(ta));
}
__device__ __forceinline__ void t_ld16(uint32_t ta, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(ta));
}
__device__ __forceinline__ void t_ld8(uint32_t ta, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(ta));
}
__device__ __forceinline__ void t_ld1(uint32_t ta, uint32_t& r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(ta));
}
__device__ __forceinline__ void t_st16(uint32_t ta, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(ta), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void t_st8(uint32_t ta, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(ta), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void t_st4(uint32_t ta, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ta), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void t_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void t_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float t_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float t_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t t_lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint4 t_lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void t_sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void t_sts128(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void t_fhfma(float& acc, uint16_t a, uint16_t b) {
    asm("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(acc) : "h"(a), "h"(b));
}
__device__ __forceinline__ float2 t_h2f(uint32_t u) { return __half22float2(*reinterpret_cast<const __half2*>(&u)); }
#ifdef ARSEG_TC_NOALLOC
__device__ __forceinline__ uint4 t_ldg128(const char* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
#else
__device__ __forceinline__ uint4 t_ldg128(const char* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
#endif

struct TSmem {
    uint32_t sK, sV, sQ, sW, sOnes, rings;       // shared-window addresses; rings = hr ring, lr ring, scratch
    uint4* posa;                                 // [2][PMAX] gather records: {w0 | w1 << 16, w2 | w3 << 16 (f16), source pixel | lr << 31, ring offset | swizzle << 28}
    float* s_bc;                                 // [TNCLS]
    uint64_t *gfull /* lr_up rows gathered, index g + 1 */, *hfull /* warped-hr rows landed (bulk copies), index g + 1 */, *ddone, *lrfree, *sfull, *pfull, *ofull,
             *ofree, *afull, *lfull;
};

// ---------------------------------------------------------------------------------------------
// G role: gather.  Step g (g = -1 .. NH-1) fills warped-hr ring rows [h0, h0+nh) and lr_up ring rows [l0, l0+nl).
// hr row r <-> image row ya - R - PADR - 1 + r (column c <-> x0 - R - 1 + c); lr row r <-> image row ya - 1 + r
// (column c <-> x0 - 1 + c).
// ---------------------------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ void t_step_geom(int g, int& h0, int& nh, int& l0, int& nl) {
    using Cf = TCfg<K>;
    if (g < 0) { h0 = 0; nh = 2; l0 = 0; nl = 0; return; }
    h0 = 4 * g + 2; nh = 4;
    if (g >= Cf::HP) { l0 = 4 * (g - Cf::HP) + 2; nl = 4; }
    else if (g == Cf::HP - 1) { l0 = 0; nl = 2; }
    else { l0 = 0; nl = 0; }
}

// Warped-hr rows: bulk copies from the pre-pass workspace (ring row r <-> workspace row ya + r, columns x0 .. x0 + HC).
// lr_up rows: the per-position gather records (2x2 source block + four bilinear weights) are computed once per pixel by
// creff_tc_rec_kernel into the caller's workspace; the G warps only copy the records of the next step into shared memory
// (one 16-byte load per thread, in flight during the gather loop) and stream the taps.
__device__ __forceinline__ void t_bulk_row(uint32_t dst, const char* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
template <int K>
__device__ __forceinline__ void t_g_role(const CreffMmaParams& p, const TSmem& sm, const uint4* __restrict__ rec, const char* __restrict__ warped,
                                         int n, int x0, int ya, int NH, int S) {
    using Cf = TCfg<K>;
    const int gt = (int)threadIdx.x - 32 * TG_WARP0, lane = gt & 31, qw = gt >> 3, l8 = lane & 7;    // thread / quarter-warp index inside the G group
    const char* const lrb = reinterpret_cast<const char*>(p.lr) + (size_t)n * p.h * p.w * 128;
    const uint32_t lr_rs = (uint32_t)p.w * 128;
    constexpr uint32_t SCRATCH_OFF = (uint32_t)(Cf::HR_BYTES + Cf::LR_BYTES);
    const uint4* const rec_lr = rec;                                   // [H][W] (frame independent)
    const size_t wrow = (size_t)t_wp<K>(p.W) * 128;                    // workspace row pitch
    const char* const wsrc = warped + ((size_t)n * t_hp<K>(p.H) + ya) * wrow + (size_t)x0 * 128;

    // record slot q of step g: destination ring offset and the global record it copies (nullptr: a zero record -- positions
    // outside the image are the depthwise convolutions' zero padding; slots beyond the step's positions are no-ops)
    auto slot_of = [&](int g, int q, uint32_t& dst) -> const uint4* {
        int h0, nh, l0, nl;
        t_step_geom<K>(g, h0, nh, l0, nl);
        const int npos = nl * TLC;
        dst = SCRATCH_OFF;
        if (q < npos) {
            const int rr = q / TLC, cc = q - rr * TLC, row = l0 + rr;
            const int fy = ya - 1 + row, fx = x0 - 1 + cc;
            dst = (uint32_t)(Cf::HR_BYTES + ((row % TLRR) * TLC + cc) * 128) | ((uint32_t)(cc & 7) << 28);
            return (fy >= 0 && fy < p.H && fx >= 0 && fx < p.W) ? rec_lr + (size_t)fy * p.W + fx : nullptr;
        }
        return nullptr;
    };
    constexpr int NR = (Cf::PMAX + TG_THREADS - 1) / TG_THREADS;       // records per thread and step (<= 2)
    uint4 rnext[NR];
    uint32_t dnext[NR];
    auto fetch = [&](int g) {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const int q = gt + r * TG_THREADS;
            rnext[r] = make_uint4(0u, 0u, 0u, 0u);
            dnext[r] = SCRATCH_OFF;
            if (q < Cf::PMAX) {
                const uint4* src = slot_of(g, q, dnext[r]);
                if (src) rnext[r] = __ldg(src);
            }
        }
    };
    auto publish = [&](int g) {
        const int buf = (g + 1) & 1;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const int q = gt + r * TG_THREADS;
            if (q < Cf::PMAX) sm.posa[buf * Cf::PMAX + q] = make_uint4(rnext[r].x, rnext[r].y, rnext[r].z, dnext[r]);
        }
    };
    uint4 tap[TJA][4];
    auto issue = [&](uint4 (&tp)[4], int buf, int j) {
        const uint4 id = sm.posa[buf * Cf::PMAX + qw + TNQW * j];
        const char* a0 = lrb + (size_t)(id.z & 0x7fffffffu) * 128 + 16 * l8;
        const char* a1 = a0 + lr_rs;
        tp[0] = t_ldg128(a0);
        tp[1] = t_ldg128(a0 + 128);
        tp[2] = t_ldg128(a1);
        tp[3] = t_ldg128(a1 + 128);
    };
    auto commit = [&](const uint4 (&tp)[4], int buf, int j) {
        const uint4 id = sm.posa[buf * Cf::PMAX + qw + TNQW * j];
        const uint32_t* t0 = reinterpret_cast<const uint32_t*>(&tp[0]);
        const uint32_t* t1 = reinterpret_cast<const uint32_t*>(&tp[1]);
        const uint32_t* t2 = reinterpret_cast<const uint32_t*>(&tp[2]);
        const uint32_t* t3 = reinterpret_cast<const uint32_t*>(&tp[3]);
        // mixed-precision FMA (f16 x f16 + f32 -> f32, SASS FHFMA): the taps are consumed as they are, no f16 -> f32 conversions;
        // the bilinear weights are f16 (<= 2^-12 relative, below the f16 rounding of the result itself)
        const uint16_t w0 = (uint16_t)(id.x & 0xffffu), w1 = (uint16_t)(id.x >> 16), w2 = (uint16_t)(id.y & 0xffffu), w3 = (uint16_t)(id.y >> 16);
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float vx = 0.f, vy = 0.f;
            t_fhfma(vx, (uint16_t)(t0[e] & 0xffffu), w0); t_fhfma(vy, (uint16_t)(t0[e] >> 16), w0);
            t_fhfma(vx, (uint16_t)(t1[e] & 0xffffu), w1); t_fhfma(vy, (uint16_t)(t1[e] >> 16), w1);
            t_fhfma(vx, (uint16_t)(t2[e] & 0xffffu), w2); t_fhfma(vy, (uint16_t)(t2[e] >> 16), w2);
            t_fhfma(vx, (uint16_t)(t3[e] & 0xffffu), w3); t_fhfma(vy, (uint16_t)(t3[e] >> 16), w3);
            o[e] = pack_h2_sat(vx, vy);
        }
        const uint32_t d = id.w;
        t_sts128(sm.rings + (d & 0x0fffffffu) + (((uint32_t)l8 ^ (d >> 28)) << 4), make_uint4(o[0], o[1], o[2], o[3]));
    };

    fetch(-1);
    publish(-1);
    nbar_sync(TBAR_G, TG_THREADS);
#pragma unroll 1
    for (int g = -1; g < NH; ++g) {
        const int buf = (g + 1) & 1;
        // rolling pipeline: TJA positions' loads are always in flight while the oldest one is combined and stored
#pragma unroll
        for (int j = 0; j < TJA; ++j) issue(tap[j], buf, j);
        if (g + 1 < NH) fetch(g + 1);                               // the next step's records: in flight during the gather loop
        if (g >= 2) tbar_wait<true>(sm.ddone, g - 2, 2);            // D half-step g-2 done: the hr / lr rows this step overwrites are read
        if (gt == 0) {
            // the step's warped-hr rows: one bulk copy each, completing on their own hand-off barrier -- the K / V warps (the
            // producers that set the pace) wait for these only, never for the lr_up gather below
            int h0, nh, l0, nl;
            t_step_geom<K>(g, h0, nh, l0, nl);
            const uint32_t bar = s_u32(sm.hfull + ((g + 1) & (TNB - 1)));
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)nh * Cf::HROW_BYTES) : "memory");
            for (int r = 0; r < nh; ++r)
                t_bulk_row(sm.rings + (uint32_t)((h0 + r) % THRR) * Cf::HROW_BYTES, wsrc + (size_t)(h0 + r) * wrow, Cf::HROW_BYTES, bar);
        }
        {
            // lr_up rows double as the residual: E takes the rows of tile i at the start of its step
            const int num = 4 * (g - Cf::HP) + 5 - TLRR - 1;        // last overwritten lr row - 1
            if (num >= 0) { const int im = num >> 3; tbar_wait<true>(sm.lrfree, im < S - 1 ? im : S - 1, 3); }
        }
#ifdef ARSEG_TTRACE
        if (!(p.dbg & 4))
#endif
#pragma unroll 1
        for (int j0 = 0; j0 < Cf::NJ; j0 += TJA) {
#pragma unroll
            for (int j = 0; j < TJA; ++j) {
                if (j0 + j < Cf::NJ) {
                    commit(tap[j], buf, j0 + j);
                    if (j0 + j + TJA < Cf::NJ) issue(tap[j], buf, j0 + j + TJA);
                }
            }
        }
        tbar_arrive(sm.gfull, g + 1);
        if (g + 1 < NH) publish(g + 1);
        nbar_sync(TBAR_G, TG_THREADS);                              // records of step g+1 visible; everyone is done with those of step g-1
    }
}

// ---------------------------------------------------------------------------------------------
// D role: depthwise 3x3 convolutions, four output rows per pass with an x-marching 6-row x 3-column register window.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 t_dw9(const float2 (&w)[10], const float2 (&r0)[3], const float2 (&r1)[3], const float2 (&r2)[3], int sa, int sb, int sc) {
    float2 a0 = __ffma2_rn(w[0], r0[sa], w[9]), a1 = __fmul2_rn(w[3], r1[sa]), a2 = __fmul2_rn(w[6], r2[sa]);
    a0 = __ffma2_rn(w[1], r0[sb], a0); a1 = __ffma2_rn(w[4], r1[sb], a1); a2 = __ffma2_rn(w[7], r2[sb], a2);
    a0 = __ffma2_rn(w[2], r0[sc], a0); a1 = __ffma2_rn(w[5], r1[sc], a1); a2 = __ffma2_rn(w[8], r2[sc], a2);
    return __fadd2_rn(__fadd2_rn(a0, a1), a2);
}

// jmax: the last tile whose K/V rows are overwritten by half-step h (K/V row kr lives in ring slot kr % KVR)
template <int K>
__device__ __forceinline__ int t_kv_last_reader(int h) {
    using Cf = TCfg<K>;
    const int num = 4 * h + 3 - Cf::KVR - Cf::PADR;
    return num >= 0 ? (num >> 3) : -1;
}

template <int K>
__device__ __forceinline__ void t_d_role(const CreffMmaParams& p, const TSmem& sm, int x0, int ya, int NH, int S) {
    using Cf = TCfg<K>;
    const int lane = threadIdx.x & 31, dw = (int)(threadIdx.x >> 5), qhalf = (dw - TQ_WARP0) & 1, qrow = ((dw - TQ_WARP0) >> 1) * TQ_ROWS;
    const uint32_t lane_sub = (uint32_t)((lane & 3) * 4), lane_chunk = (uint32_t)(lane >> 2);
    if (dw < TQ_WARP0) {
        // ---------------- K (warps 0..2) or V (3..5): columns [c_lo, c_lo + ncol) of four K/V rows per half-step ----------------
        const int isv = dw / TKV_SPLIT, part = dw % TKV_SPLIT;
        const int c_lo = part * Cf::KVC3, ncol = min(Cf::KVC3, Cf::KVC - c_lo);
        float2 w[10];
        {
            const float* wp = isv ? p.wv : p.wk;
            const float* bp = isv ? p.bv : p.bk;
#pragma unroll
            for (int i = 0; i < 9; ++i) w[i] = make_float2(__ldg(wp + (2 * lane) * 9 + i), __ldg(wp + (2 * lane + 1) * 9 + i));
            w[9] = make_float2(__ldg(bp + 2 * lane), __ldg(bp + 2 * lane + 1));
        }
        const uint32_t ring = isv ? sm.sV : sm.sK;
        const uint32_t hbase = sm.rings + (uint32_t)(c_lo * 128 + lane * 4);
#pragma unroll 1
        for (int h = 0; h < NH; ++h) {
            if (h == 0) tbar_wait<true>(sm.hfull, 0, 1);                        // rows 0, 1 (step -1) complete on their own barrier phase
            tbar_wait<true>(sm.hfull, h + 1, 1);
            {
                const int jm = t_kv_last_reader<K>(h);
                if (jm >= 0) tbar_wait<true>(sm.ofull, jm < S - 1 ? jm : S - 1, 4);     // P V of the last tile that read these ring slots has retired
            }
            uint32_t ra[6], ko[4];
            uint32_t rok[4];                                                   // all-ones / zero masks: the select must stay branch-free
#pragma unroll
            for (int i = 0; i < 6; ++i) ra[i] = hbase + (uint32_t)(((4 * h + i) % THRR) * (Cf::HC * 128));
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int kr = 4 * h + r, fy = ya - Cf::R - Cf::PADR + kr;
                ko[r] = ring + (uint32_t)((kr % Cf::KVR) * TROWB) + lane_sub;
                rok[r] = (fy >= 0 && fy < p.H) ? 0xffffffffu : 0u;             // K / V are exactly 0 outside the image (attention zero padding)
            }
#ifdef ARSEG_TTRACE
            if (p.dbg & 1) { t_fence_async_smem(); tbar_arrive(sm.ddone, h); continue; }
#endif
            float2 win[6][3];
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int i = 0; i < 6; ++i) win[i][c] = t_h2f(t_lds32(ra[i] + c * 128));
            // three columns per trip (the register window rotates by renaming inside a trip); the loop stays rolled: the five
            // roles of the CTA share one instruction cache
#pragma unroll 1
            for (int x3 = 0; x3 < ncol; x3 += 3) {
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    const int x = x3 + u;
                    if (x < ncol) {
                        const int sa = u, sb = (u + 1) % 3, sc = (u + 2) % 3;
#pragma unroll
                        for (int i = 0; i < 6; ++i) win[i][sc] = t_h2f(t_lds32(ra[i] + (x + 2) * 128));
                        const int col = c_lo + x, fx = x0 - Cf::R + col;
                        const uint32_t okx = (fx >= 0 && fx < p.W) ? 0xffffffffu : 0u;
                        const uint32_t co = (uint32_t)(col * 128) + ((lane_chunk ^ (uint32_t)(col & 7)) << 4);
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            const float2 v = t_dw9(w, win[r], win[r + 1], win[r + 2], sa, sb, sc);
                            t_sts32(ko[r] + co, pack_h2_sat(v.x, v.y) & okx & rok[r]);
                        }
                    }
                }
            }
            t_fence_async_smem();                                              // generic-proxy writes -> visible to the UMMA reads
            tbar_arrive(sm.ddone, h);
        }
    } else {
        // ---------------- Q: four query rows per half-step (h >= HP) into the A tile ----------------
        float2 w[10];
#pragma unroll
        for (int i = 0; i < 9; ++i) w[i] = make_float2(__ldg(p.wq + (2 * lane) * 9 + i), __ldg(p.wq + (2 * lane + 1) * 9 + i));
        w[9] = make_float2(__ldg(p.bq + 2 * lane), __ldg(p.bq + 2 * lane + 1));
        uint32_t lx[2];     // byte offset of this lane's channel pair inside a 128-byte position whose column & 7 = c (chunk swizzle)
#pragma unroll
        for (int c = 0; c < 2; ++c) lx[c] = ((lane_chunk ^ (uint32_t)c) << 4) + lane_sub;
        const uint32_t lbase = sm.rings + (uint32_t)Cf::HR_BYTES;
#pragma unroll 1
        for (int h = 0; h < NH; ++h) {
            tbar_wait<true>(sm.gfull, h + 1, 1);
#ifdef ARSEG_TTRACE
            if (p.dbg & 1) { t_fence_async_smem(); tbar_arrive(sm.ddone, h); continue; }
#endif
            if (h >= Cf::HP) {
                const int q0 = 4 * (h - Cf::HP), ti = q0 >> 3, qy0 = q0 & 7;
                if (ti >= 1) tbar_wait<true>(sm.sfull, ti - 1, 5);                      // S = Q K^T of the previous tile has retired: the A tile is free
                uint32_t ra[TQ_ROWS + 2];
#pragma unroll
                for (int i = 0; i < TQ_ROWS + 2; ++i) ra[i] = lbase + (uint32_t)(((q0 + qrow + i) % TLRR) * (TLC * 128));
                const uint32_t qo = sm.sQ + (uint32_t)((qy0 + qrow) * TSW * 128);
                // this warp's column half: output columns [8 qhalf, 8 qhalf + 8) from lr ring columns [8 qhalf, 8 qhalf + 10)
                const int xlo = 8 * qhalf;
                float2 win[TQ_ROWS + 2][3];
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int i = 0; i < TQ_ROWS + 2; ++i) win[i][c] = t_h2f(t_lds32(ra[i] + (xlo + c) * 128 + lx[c]));     // (xlo + c) & 7 = c
#pragma unroll 1
                for (int x3 = 0; x3 < TSW / 2; x3 += 3) {
#pragma unroll
                    for (int u = 0; u < 3; ++u) {
                        if (x3 + u < TSW / 2) {
                            const int x = xlo + x3 + u;
                            const int sa = u, sb = (u + 1) % 3, sc = (u + 2) % 3;
                            const uint32_t lin = (uint32_t)((x + 2) * 128) + (((lane_chunk ^ (uint32_t)((x + 2) & 7)) << 4) + lane_sub);
#pragma unroll
                            for (int i = 0; i < TQ_ROWS + 2; ++i) win[i][sc] = t_h2f(t_lds32(ra[i] + lin));
                            const uint32_t qst = qo + (uint32_t)(x * 128) + (((lane_chunk ^ (uint32_t)(x & 7)) << 4) + lane_sub);
#pragma unroll
                            for (int r = 0; r < TQ_ROWS; ++r) {
                                const float2 v = t_dw9(w, win[r], win[r + 1], win[r + 2], sa, sb, sc);
                                t_sts32(qst + (uint32_t)(r * TSW * 128), pack_h2_sat(v.x, v.y));
                            }
                        }
                    }
                }
            }
            t_fence_async_smem();
            tbar_arrive(sm.ddone, h);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// M role: the MMA-issuing thread.
// ---------------------------------------------------------------------------------------------
// The issuing thread is ONE lane: every instruction it executes is a dependent scalar chain, so the descriptors are advanced
// incrementally (one add per MMA) instead of being rebuilt from row / column arithmetic.
template <int K>
__device__ __forceinline__ void t_issue_qk(const TSmem& sm, uint32_t tmem, int i) {
    using Cf = TCfg<K>;
    // key rows [8i + PADR, + NKR) of the ring in as few MMAs as possible: up to TEN rows (N = 240 <= 256) per instruction, cut at
    // the ring end (row counts are even, so every piece is a multiple of N = 48).  Small-N UMMA is issue-bound (28 MMAs of
    // N = 48 retire in ~1060 cycles against ~340 of math); ARSEG_TC_QK48 restores the fixed two-row pieces.
    const int b0 = (8 * i + Cf::PADR) % Cf::KVR;
    const uint64_t dq = t_desc(sm.sQ), dk = t_desc(sm.sK);
    uint32_t td = tmem;
    int row = b0, left = Cf::NKR;
#pragma unroll 1
    while (left > 0) {
#ifdef ARSEG_TC_QK48
        const int seg = 2;
#else
        int seg = min(left, Cf::KVR - row);
        seg = seg > 10 ? 10 : seg;
#endif
        const uint32_t id = t_idesc(seg * TKP, 0);
        const uint64_t dkr = dk + (uint64_t)((uint32_t)row * (TROWB >> 4));
#pragma unroll
        for (int k = 0; k < 4; ++k) t_mma_ss(dq + (uint64_t)(2 * k), dkr + (uint64_t)(2 * k), td, k != 0, id);
        td += (uint32_t)(seg * TKP);
        row += seg;
        row = row >= Cf::KVR ? row - Cf::KVR : row;
        left -= seg;
    }
}
template <int K>
__device__ __forceinline__ void t_issue_pv(const TSmem& sm, uint32_t tmem, int i) {
    using Cf = TCfg<K>;
    const int b0 = (8 * i + Cf::PADR) % Cf::KVR;
    constexpr uint32_t IDV = t_idesc(64, 1), IDS = t_idesc(16, 0);
    constexpr uint32_t RING16 = (uint32_t)(Cf::KVR * TROWB) >> 4;      // ring size in descriptor units (16 bytes)
    const uint64_t ones = t_desc(sm.sOnes), dv = t_desc(sm.sV);
    // 16 keys (2 KB of ring) per MMA: P columns [8 ks, 8 ks + 8) x V rows; the row sums come from the same A against ones.
    // An even number of rows precedes the ring end, so a 16-key step never straddles it: the offset just wraps.
    uint32_t off = (uint32_t)(b0 * (TROWB >> 4)), ta = tmem;
#pragma unroll
    for (int ks = 0; ks < Cf::NK / 16; ++ks) {
        t_mma_ts(ta, dv + off, tmem + Cf::COL_O, ks != 0, IDV);
        t_mma_ts(ta, ones, tmem + Cf::COL_SUM, ks != 0, IDS);
        off += 2048 >> 4;
        off = off >= RING16 ? off - RING16 : off;
        ta += 8;
    }
}
template <int K, int NCP>
__device__ __forceinline__ void t_issue_cls(const TSmem& sm, uint32_t tmem) {
    using Cf = TCfg<K>;
    constexpr uint32_t ID = t_idesc(NCP, 0);
#pragma unroll
    for (int k = 0; k < 4; ++k) t_mma_ts(tmem + Cf::COL_A + (uint32_t)(k * 8), t_desc(sm.sW + k * 32), tmem + Cf::COL_L, k != 0, ID);
}

__device__ __forceinline__ bool t_elect() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// The whole warp runs the control flow (warp-uniform operands live in uniform registers, which tcgen05.mma reads
// directly); one elected lane issues.  Per tile: S = Q K^T -> (softmax warps) -> O = P V; then, in whichever order they
// become possible, the classifier MMA of this tile (after the epilogue warps have written its input) and S of the NEXT
// tile (its columns are free once P V has been issued -- the tensor pipe runs in order).
template <int K, int NCP>
__device__ __forceinline__ void t_m_role(const TSmem& sm, uint32_t tmem, int S, int dbg) {
    using Cf = TCfg<K>;
    bool qk_ahead = false;          // S of this tile already issued
#pragma unroll 1
    for (int i = 0; i < S; ++i) {
        if (!qk_ahead) {
            tbar_wait(sm.ddone, 2 * i + Cf::HP + 1, 6);           // K/V rows and Q of tile i are in shared memory
            TEV(0, i, 0);
            t_fence_after();
            if (t_elect()) { t_issue_qk<K>(sm, tmem, i); t_commit(sm.sfull, i); }
            __syncwarp();
            TEV(0, i, 1);
        }
        qk_ahead = false;
        tbar_wait(sm.pfull, i, 7);                                   // P is in tensor memory
        if (i >= 1) tbar_wait(sm.ofree, i - 1, 12);                  // the epilogue warps have read O of the previous tile
        TEV(0, i, 2);
        t_fence_after();
        if (t_elect()) { t_issue_pv<K>(sm, tmem, i); t_commit(sm.ofull, i); }
        __syncwarp();
        TEV(0, i, 3);
        bool need_cls = NCP > 0, need_qk = i + 1 < S;
#ifdef ARSEG_TTRACE
        if (dbg & 8) { need_qk = false; if (i >= 0) tbar_wait(sm.ofree, i, 12); }     // no overlap between tiles (debugging)
#endif
#pragma unroll 1
        for (int spin = 0; need_cls || need_qk; ++spin) {
            if (need_cls && tbar_test(sm.afull, i)) {                // residual + O / sum is back in tensor memory (f16)
                TEV(0, i, 4);
                t_fence_after();
                if (t_elect()) { t_issue_cls<K, NCP>(sm, tmem); t_commit(sm.lfull, i); }
                __syncwarp();
                TEV(0, i, 5);
                need_cls = false;
            }
            if (need_qk && tbar_test(sm.ddone, 2 * (i + 1) + Cf::HP + 1)) {
                t_fence_after();
                if (t_elect()) { t_issue_qk<K>(sm, tmem, i + 1); t_commit(sm.sfull, i + 1); }
                __syncwarp();
                need_qk = false;
                qk_ahead = true;
            }
            if (spin > (1 << 22)) tbar_timeout(13, i);
            if (spin >= 8) __nanosleep(40);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// S role: softmax.  Warp of lane quarter cw owns TMEM lanes 32cw .. 32cw+31 = query rows 2cw, 2cw+1 of the tile.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void t_ld_row(uint32_t ta, uint32_t* s) { t_ld16(ta, s); t_ld8(ta + 16, s + 16); }

template <int K>
__device__ __forceinline__ void t_s_role(const TSmem& sm, uint32_t tmem, int S, int dbg) {
    using Cf = TCfg<K>;
    constexpr float LOG2E = 1.4426950408889634f;
    static_assert((K + 1) % 2 == 0, "the row loops are unrolled by two");
    const int lane = threadIdx.x & 31, cw = (int)(threadIdx.x >> 5) & 3;
    const int b = lane >> 4, qx = lane & 15;
    const uint32_t tq = tmem + ((uint32_t)(32 * cw) << 16);
    // cbm[c]: the column mask of this lane's window (key column c is inside iff qx <= c < qx + K), kept in ONE register array:
    // 0 / -inf during pass 1 (max of s + cbm), -max * log2(e) / -inf during pass 2 (ex2(s * log2e + cbm)); "+= ml" after pass 2
    // restores exact zeros (-ml + ml), so a tile's arithmetic does not depend on the tiles before it (bit-exact row segments).
    float cbm[Cf::KVC];
#pragma unroll
    for (int c = 0; c < Cf::KVC; ++c) cbm[c] = ((unsigned)(c - qx) < (unsigned)K) ? 0.f : -INFINITY;
    // the warp reads key rows 2cw .. 2cw + K of the tile; the first is outside the window of its odd query row, the last outside that of the even one
    const float rb_first = b ? -INFINITY : 0.f, rb_last = b ? 0.f : -INFINITY;
    const uint32_t srow = tq + (uint32_t)(2 * cw * TKP);            // S columns of the warp's first key row
    const uint32_t prow = tq + (uint32_t)(2 * cw * (TKP / 2));      // P columns of the warp's first key row

    auto rowmax = [&](const uint32_t* s) -> float {
        float m0 = -INFINITY, m1 = -INFINITY;                       // two chains
#pragma unroll
        for (int c = 0; c < Cf::KVC; c += 2) {
            m0 = fmaxf(m0, __uint_as_float(s[c]) + cbm[c]);
            if (c + 1 < Cf::KVC) m1 = fmaxf(m1, __uint_as_float(s[c + 1]) + cbm[c + 1 < Cf::KVC ? c + 1 : c]);
        }
        return fmaxf(m0, m1);
    };
    auto exprow = [&](const uint32_t* s, int j) {
        const uint32_t keep = ((j == 0 && b) || (j == K && !b)) ? 0u : 0xffffffffu;
        uint32_t pk[12];
#pragma unroll
        for (int c = 0; c < 24; c += 2) {
            if (c < Cf::KVC) {
                const float e0 = t_ex2(fmaf(__uint_as_float(s[c]), LOG2E, cbm[c]));
                const float e1 = c + 1 < Cf::KVC ? t_ex2(fmaf(__uint_as_float(s[c + 1]), LOG2E, cbm[c + 1 < Cf::KVC ? c + 1 : c])) : 0.f;
                pk[c >> 1] = pack_h2(e0, e1) & keep;
            } else {
                pk[c >> 1] = 0u;
            }
        }
        t_st8(prow + j * (TKP / 2), pk);
        t_st4(prow + j * (TKP / 2) + 8, pk[8], pk[9], pk[10], pk[11]);
    };

#pragma unroll 1
    for (int i = 0; i < S; ++i) {
        tbar_wait(sm.sfull, i, 9);
        if (cw == 0) TEV(1, i, 1);
        t_fence_after();
#ifdef ARSEG_TTRACE
        if (dbg & 16) { t_fence_before(); tbar_arrive(sm.pfull, i); continue; }      // timing experiment: no softmax work (results are garbage)
#endif
        // ---------------- pass 1: the shift of the maximum (the next row's TMEM load is in flight while a row is reduced) ----------------
        uint32_t sa[24], sb[24];
        float d = -INFINITY;
        t_ld_row(srow, sa);
#pragma unroll 1
        for (int j = 0; j <= K; j += 2) {
            t_ld_wait();
            t_ld_row(srow + (j + 1) * TKP, sb);
            d = fmaxf(d, rowmax(sa) + (j == 0 ? rb_first : 0.f));
            t_ld_wait();
            if (j + 2 <= K) t_ld_row(srow + (j + 2) * TKP, sa); else t_ld_row(srow, sa);      // last trip: row 0 again, for pass 2
            d = fmaxf(d, rowmax(sb) + (j + 1 == K ? rb_last : 0.f));
        }
        const float ml = d * LOG2E;
#pragma unroll
        for (int c = 0; c < Cf::KVC; ++c) cbm[c] -= ml;
        if (cw == 0) TEV(1, i, 2);
        // ---------------- pass 2: P = exp(s - max) as f16 pairs over the S columns already consumed ----------------
#pragma unroll 1
        for (int j = 0; j <= K; j += 2) {
            t_ld_wait();
            t_ld_row(srow + (j + 1) * TKP, sb);
            exprow(sa, j);
            t_ld_wait();
            if (j + 2 <= K) t_ld_row(srow + (j + 2) * TKP, sa);
            exprow(sb, j + 1);
        }
#pragma unroll
        for (int c = 0; c < Cf::KVC; ++c) cbm[c] += ml;              // back to exact 0 / -inf
        if (cw == 0) TEV(1, i, 3);
        // P of the key rows outside the warp's range is zero (after the S reads: the P columns alias S): rows [0, 2cw) and
        // (2cw + K, NKR) = 3cw + (9 - 3cw) chunks of 8 columns
        {
            const uint32_t z[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            constexpr int NZ = (Cf::NKR - K - 1) * (TKP / 2) / 8;
            static_assert((Cf::NKR - K - 1) * (TKP / 2) % 8 == 0 && (2 * (TKP / 2)) % 8 == 0, "zero fill in chunks of 8 columns");
#pragma unroll 1
            for (int t = 0; t < NZ; ++t) {
                const int lo = 3 * cw;                               // chunks below the warp's rows (24 cw columns)
                const uint32_t col = t < lo ? (uint32_t)(8 * t) : (uint32_t)((2 * cw + K + 1) * (TKP / 2) + 8 * (t - lo));
                t_st8(tq + col, z);
            }
        }
        t_st_wait();
        t_fence_before();
        tbar_arrive(sm.pfull, i);
        if (cw == 0) TEV(1, i, 4);
    }
}

// ---------------------------------------------------------------------------------------------
// E role: residual, classifier input, log-softmax / argmax, stores.  Same lane ownership as S.
// ---------------------------------------------------------------------------------------------
template <int K, int NCP>
__device__ __forceinline__ void t_e_role(const CreffMmaParams& p, const TSmem& sm, uint32_t tmem, int n, int x0, int ya, int yb, int S) {
    using Cf = TCfg<K>;
    constexpr float LOG2E = 1.4426950408889634f;
    const int lane = threadIdx.x & 31, cw = (int)(threadIdx.x >> 5) & 3;
    const int b = lane >> 4, qx = lane & 15, qy = 2 * cw + b;
    const uint32_t tq = tmem + ((uint32_t)(32 * cw) << 16);
    const int W = p.W;
    const size_t plane = (size_t)p.H * W;
    const int px = x0 + qx;
    float* const ol = p.out_logits ? p.out_logits + (size_t)n * p.ncls * plane : nullptr;
    float* const op = p.out_p ? p.out_p + (size_t)n * MC * plane : nullptr;
    uint8_t* const oa = p.out_argmax ? p.out_argmax + (size_t)n * plane : nullptr;
    const uint32_t res_col = (uint32_t)((1 + qx) * 128), res_sw = (uint32_t)((1 + qx) & 7);
    const uint32_t lbase = sm.rings + (uint32_t)Cf::HR_BYTES;

#pragma unroll 1
    for (int i = 0; i < S; ++i) {
        // ---------------- residual lr_up (model/attention.py:191,210): lr ring row 8i + 1 + qy, column 1 + qx ----------------
        // (a parity wait may only name a phase the barrier has already entered: every role's FIRST wait on a hand-off uses an
        // index < TNB and later ones advance by < TNB, so "tile ready" = ddone(2i + HP + 1) is used here, not gfull(2i + HP + 2))
        tbar_wait(sm.ddone, 2 * i + Cf::HP + 1, 1);
        if (cw == 0) TEV(2, i, 0);
        uint4 res[8];
        {
            const uint32_t ra = lbase + (uint32_t)(((8 * i + 1 + qy) % TLRR) * (TLC * 128)) + res_col;
#pragma unroll
            for (int c = 0; c < 8; ++c) res[c] = t_lds128(ra + (((uint32_t)c ^ res_sw) << 4));
        }
        tbar_arrive(sm.lrfree, i);
        // ---------------- fused = lr_up + O / sum (model/attention.py:207,210) ----------------
        tbar_wait(sm.ofull, i, 10);
#ifdef ARSEG_TTRACE
        if (p.dbg & 32) {      // timing experiment: no epilogue work (results are garbage)
            t_fence_before(); tbar_arrive(sm.ofree, i);
            if (NCP > 0) { tbar_arrive(sm.afull, i); tbar_wait(sm.lfull, i, 11); }
            continue;
        }
#endif
        if (cw == 0) TEV(2, i, 5);
        t_fence_after();
        const int py = ya + 8 * i + qy;
        const bool ok = py < yb && px < W;
        const size_t off = (size_t)py * W + px;
        uint32_t su;
        t_ld1(tq + Cf::COL_SUM, su);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            uint32_t o[32];
            t_ld32(tq + Cf::COL_O + 32 * hf, o);
            t_ld_wait();
            if (hf == 1 && !op) { t_fence_before(); tbar_arrive(sm.ofree, i); }       // O and the row sum are in registers
            const float inv = t_rcp(__uint_as_float(su));
            uint32_t pk[16];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint32_t* rr = reinterpret_cast<const uint32_t*>(&res[4 * hf + c]);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 r2 = t_h2f(rr[e]);
                    const int ch = 8 * c + 2 * e;                              // channel 32 hf + ch
                    const float f0 = fmaf(__uint_as_float(o[ch]), inv, r2.x), f1 = fmaf(__uint_as_float(o[ch + 1]), inv, r2.y);
                    pk[4 * c + e] = pack_h2_sat(f0, f1);
                }
            }
            if (NCP > 0) t_st16(tq + Cf::COL_A + 16 * hf, pk);
        }
        if (op) {
            // the fused feature itself (the second value of forward_phase2, discarded by evaluation.py:193): a cold, warp-uniform
            // branch that re-reads O instead of predicating 64 stores into the hot path
            float* o_ptr = op + off;
            const float inv = t_rcp(__uint_as_float(su));
#pragma unroll 1
            for (int q8 = 0; q8 < 8; ++q8) {
                uint32_t o[8];
                t_ld8(tq + Cf::COL_O + 8 * q8, o);
                t_ld_wait();
                const uint32_t* rr = reinterpret_cast<const uint32_t*>(&res[0]);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 r2 = t_h2f(rr[e]);                            // res[0] = channels 8 q8 .. 8 q8 + 7 (rotated below)
                    if (ok) {
                        o_ptr[(size_t)(2 * e) * plane] = fmaf(__uint_as_float(o[2 * e]), inv, r2.x);
                        o_ptr[(size_t)(2 * e + 1) * plane] = fmaf(__uint_as_float(o[2 * e + 1]), inv, r2.y);
                    }
                }
                o_ptr += 8 * plane;
                const uint4 t0 = res[0];
#pragma unroll
                for (int c = 0; c < 7; ++c) res[c] = res[c + 1];
                res[7] = t0;
            }
            t_fence_before();
            tbar_arrive(sm.ofree, i);
        }
        if (NCP == 0) continue;
        t_st_wait();
        t_fence_before();
        tbar_arrive(sm.afull, i);
        if (cw == 0) TEV(2, i, 6);
        // ---------------- classifier (model/pspnet.py:226), log-softmax (:229), argmax (evaluation.py:204) ----------------
        tbar_wait(sm.lfull, i, 11);
        if (cw == 0) TEV(2, i, 7);
        t_fence_after();
        constexpr int NCPA = NCP > 0 ? NCP : 16;
        uint32_t lg[NCPA];
        t_ld16(tq + Cf::COL_L, lg);
        if (NCPA > 16) t_ld16(tq + Cf::COL_L + 16, lg + (NCPA > 16 ? 16 : 0));
        t_ld_wait();
        float lmax = -INFINITY;
        int am = 0;
#pragma unroll
        for (int j = 0; j < NCPA; ++j) {
            float v = __uint_as_float(lg[j]) + sm.s_bc[j];
            v = j < p.ncls ? v : -INFINITY;                                    // padding classes
            lg[j] = __float_as_uint(v);
            if (v > lmax) { lmax = v; am = j; }                                // first maximum, like torch.argmax
        }
        float lse = 0.f;
        if (p.log_softmax) {
            const float q0 = lmax * LOG2E;
#pragma unroll
            for (int j = 0; j < NCPA; ++j) lse += t_ex2(fmaf(__uint_as_float(lg[j]), LOG2E, -q0));
            lse = __logf(lse) + lmax;
        }
        if (ok) {
            if (ol) {
#pragma unroll
                for (int j = 0; j < NCPA; ++j)
                    if (j < p.ncls) ol[(size_t)j * plane + off] = __uint_as_float(lg[j]) - lse;
            }
            if (oa) oa[off] = (uint8_t)am;
        }
        if (cw == 0) TEV(2, i, 8);
    }
}

// ---------------------------------------------------------------------------------------------
// pre-pass 1: lr_up gather records, once per pixel ([H][W], frame independent)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) creff_tc_rec_kernel(CreffMmaParams p, uint4* __restrict__ rec) {
    const long long total = (long long)p.H * p.W;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int fy = (int)(i / p.W), fx = (int)(i % p.W);
    const PosRec r = pos_lr(p, resize_scale(p.h, p.H, ARSEG_RESIZE_BILINEAR_AC), resize_scale(p.w, p.W, ARSEG_RESIZE_BILINEAR_AC), fy, fx);
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (r.info >= 0) {
        float4 w; int bx, by;
        rec_block_of(r, p.w, p.h, w, bx, by);
        o.x = pack_h2(w.x, w.y);
        o.y = pack_h2(w.z, w.w);
        o.z = (uint32_t)(by * p.w + bx);
    }
    rec[i] = o;
}

// ---------------------------------------------------------------------------------------------
// pre-pass 2: the MV warp of the keyframe feature (evaluation.py:177-183 MV rescale, :61-87 grid_sample geometry; f64 position
// arithmetic as in pos_hr) into the zero-bordered f16 workspace [N][Hp][Wp][64].  A warp owns 32 consecutive workspace pixels:
// every lane derives ONE pixel's 2x2 source block + weights, then the warp walks the 32 pixels four at a time (8 lanes x 16
// bytes per pixel, fully coalesced 512-byte stores; the 4 x 8 tap loads of a lane are all in flight together).
// ---------------------------------------------------------------------------------------------
template <int K, bool HR32>
__global__ void __launch_bounds__(TWARP_THREADS) creff_tc_warp_kernel(CreffMmaParams p, uint4* __restrict__ dst, int TWB) {
    using Cf = TCfg<K>;
    constexpr int PXB = HR32 ? 256 : 128;          // bytes per source pixel
    constexpr int NB = HR32 ? 2 : 4;               // quads of pixels per pass: 16 tap loads (16 bytes each) per lane in flight
    const int Hp = t_hp<K>(p.H), Wp = t_wp<K>(p.W);
    const long long plane = (long long)Hp * Wp;
    const int lane = threadIdx.x & 31, l8 = lane & 7;
    // work order: bands of TWB workspace rows, and inside a band the N frames one after the other -- the frames of a GOP read the
    // same keyframe rows (displaced by their MVs), so a band of the keyframe feature is fetched from DRAM once and then served
    // by L2 (frame-major order streamed the whole feature, 177 MB in fp32 = more than L2, once per frame)
    const int gpc = (TWB * Wp + 31) / 32;                                   // 32-pixel groups per (band, frame) chunk
    const long long wg = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long chunk = wg / gpc;
    const int band = (int)(chunk / p.N), n = (int)(chunk % p.N);
    const int r0 = band * TWB;
    if (r0 >= Hp) return;
    const int npx = min(TWB, Hp - r0) * Wp;                                 // pixels of this chunk
    const int j0 = (int)(wg % gpc) * 32;
    if (j0 >= npx) return;
    const long long base = (long long)n * plane + (long long)r0 * Wp + j0;  // first workspace pixel of the warp
    const int nval = min(32, npx - j0);
    float4 wf = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t src = 0xffffffffu;
    if (lane < nval) {
        const int j = j0 + lane;
        const int fy = r0 + j / Wp - Cf::PT, fx = j % Wp - Cf::PX;
        // info < 0 outside the image (depthwise zero padding) and for samples with no tap inside; the grid normalisation is an
        // f64 multiply, as in the march engine (creff_march.cu)
        const PosRec r = pos_hr<true>(p, n, fy, fx, 2.0 / (double)max(p.W - 1, 1), 2.0 / (double)max(p.H - 1, 1));
        if (r.info >= 0) {
            int bx, by;
            rec_block_of(r, p.W, p.H, wf, bx, by);
            src = (uint32_t)((p.hr_shared ? 0 : n) * (p.H * p.W) + by * p.W + bx);
        }
    }
    // f16 source: the taps are consumed as they are by mixed-precision FMAs (f16 x f16 + f32 -> f32, SASS FHFMA) with f16 bilinear
    // weights (<= 2^-12 relative, below the f16 rounding of the result itself); fp32 source: plain fp32 FMAs
    const uint32_t wxy = pack_h2(wf.x, wf.y), wzw = pack_h2(wf.z, wf.w);
    const char* const hrb = reinterpret_cast<const char*>(p.hr) + (PXB / 8) * l8;
    const uint32_t hr_rs = (uint32_t)p.W * PXB;
#pragma unroll 1
    for (int pass = 0; pass < 8 / NB; ++pass) {
        uint4 tap[NB][HR32 ? 8 : 4];
        float4 wq[NB];
#pragma unroll
        for (int it = 0; it < NB; ++it) {
            const int sl = (pass * NB + it) * 4 + (lane >> 3);
            if (HR32) {
                wq[it].x = __shfl_sync(0xffffffffu, wf.x, sl); wq[it].y = __shfl_sync(0xffffffffu, wf.y, sl);
                wq[it].z = __shfl_sync(0xffffffffu, wf.z, sl); wq[it].w = __shfl_sync(0xffffffffu, wf.w, sl);
            } else {
                wq[it].x = __uint_as_float(__shfl_sync(0xffffffffu, wxy, sl));
                wq[it].y = __uint_as_float(__shfl_sync(0xffffffffu, wzw, sl));
            }
            const uint32_t s = __shfl_sync(0xffffffffu, src, sl);
            if (s != 0xffffffffu) {
                const char* a0 = hrb + (size_t)s * PXB;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const char* at = a0 + (t & 1) * PXB + (t >> 1) * (size_t)hr_rs;
                    if (HR32) { tap[it][2 * t] = t_ldg128(at); tap[it][2 * t + 1] = t_ldg128(at + 16); }
                    else tap[it][t] = t_ldg128(at);
                }
            } else {
#pragma unroll
                for (int t = 0; t < (HR32 ? 8 : 4); ++t) tap[it][t] = make_uint4(0u, 0u, 0u, 0u);
                wq[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int it = 0; it < NB; ++it) {
            uint32_t o[4];
            if (HR32) {
                const float wt[4] = {wq[it].x, wq[it].y, wq[it].z, wq[it].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {                       // channels 2e, 2e + 1 of this lane's eight
                    float vx = 0.f, vy = 0.f;
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const float* f = reinterpret_cast<const float*>(&tap[it][2 * t + (e >> 1)]);
                        vx = fmaf(f[2 * (e & 1)], wt[t], vx);
                        vy = fmaf(f[2 * (e & 1) + 1], wt[t], vy);
                    }
                    o[e] = pack_h2_sat(vx, vy);
                }
            } else {
                const uint32_t wa = __float_as_uint(wq[it].x), wb = __float_as_uint(wq[it].y);
                const uint16_t w0 = (uint16_t)(wa & 0xffffu), w1 = (uint16_t)(wa >> 16), w2 = (uint16_t)(wb & 0xffffu), w3 = (uint16_t)(wb >> 16);
                const uint32_t* t0 = reinterpret_cast<const uint32_t*>(&tap[it][0]);
                const uint32_t* t1 = reinterpret_cast<const uint32_t*>(&tap[it][1]);
                const uint32_t* t2 = reinterpret_cast<const uint32_t*>(&tap[it][2]);
                const uint32_t* t3 = reinterpret_cast<const uint32_t*>(&tap[it][3]);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float vx = 0.f, vy = 0.f;
                    t_fhfma(vx, (uint16_t)(t0[e] & 0xffffu), w0); t_fhfma(vy, (uint16_t)(t0[e] >> 16), w0);
                    t_fhfma(vx, (uint16_t)(t1[e] & 0xffffu), w1); t_fhfma(vy, (uint16_t)(t1[e] >> 16), w1);
                    t_fhfma(vx, (uint16_t)(t2[e] & 0xffffu), w2); t_fhfma(vy, (uint16_t)(t2[e] >> 16), w2);
                    t_fhfma(vx, (uint16_t)(t3[e] & 0xffffu), w3); t_fhfma(vy, (uint16_t)(t3[e] >> 16), w3);
                    o[e] = pack_h2_sat(vx, vy);
                }
            }
            const int q = (pass * NB + it) * 4 + (lane >> 3);
            if (q < nval) dst[(base + q) * 8 + l8] = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
template <int K, int NCP>
__global__ void __launch_bounds__(TTHREADS, 1) creff_tc_kernel(CreffMmaParams p, const uint4* __restrict__ rec, const char* __restrict__ warped) {
    using Cf = TCfg<K>;
    extern __shared__ __align__(1024) uint8_t tsm_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tsm_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* pK = base;
    uint8_t* pV = pK + Cf::KV_BYTES;
    uint8_t* pQ = pV + Cf::KV_BYTES;
    uint8_t* pW = pQ + Cf::Q_BYTES;
    uint8_t* pOnes = pW + Cf::W_BYTES;
    uint8_t* pRings = pOnes + Cf::ONES_BYTES;
    uint8_t* pRec = pRings + Cf::HR_BYTES + Cf::LR_BYTES + Cf::SCRATCH_BYTES;
    TSmem sm;
    sm.sK = s_u32(pK); sm.sV = s_u32(pV); sm.sQ = s_u32(pQ); sm.sW = s_u32(pW); sm.sOnes = s_u32(pOnes); sm.rings = s_u32(pRings);
    sm.posa = reinterpret_cast<uint4*>(pRec);
    sm.s_bc = reinterpret_cast<float*>(pRec + Cf::REC_BYTES);
    sm.gfull = reinterpret_cast<uint64_t*>(sm.s_bc + TNCLS);
    sm.hfull = sm.gfull + TNB; sm.ddone = sm.hfull + TNB; sm.lrfree = sm.ddone + TNB; sm.sfull = sm.lrfree + TNB; sm.pfull = sm.sfull + TNB;
    sm.ofull = sm.pfull + TNB; sm.ofree = sm.ofull + TNB; sm.afull = sm.ofree + TNB; sm.lfull = sm.afull + TNB;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm.lfull + TNB);

    const int tid = threadIdx.x, warp = tid >> 5;
    // frame index fastest: the N frames of a GOP visit the same keyframe rows back to back (L2 reuse)
    int bi = blockIdx.x;
    const int n = bi % p.N; bi /= p.N;
    const int x0 = (bi % p.ncols) * TSW;
    const int ya = (bi / p.ncols) * p.seg_rows;
    const int yb = min(ya + p.seg_rows, p.H);
    const int S = (yb - ya + 7) >> 3;             // tiles
    const int NH = 2 * S + Cf::HP;                // D half-steps (G steps -1 .. NH-1)

    if (tid == 0) {
        for (int i = 0; i < TNB; ++i) {
            tbar_init(sm.gfull + i, TG_WARPS * TARRIVALS); tbar_init(sm.hfull + i, 1);   /* the bulk-copy issuer's expect_tx arrival */ tbar_init(sm.ddone + i, TD_WARPS * TARRIVALS); tbar_init(sm.lrfree + i, TC_WARPS * TARRIVALS);
            tbar_init(sm.sfull + i, 1); tbar_init(sm.pfull + i, TC_WARPS * TARRIVALS); tbar_init(sm.ofull + i, 1);
            tbar_init(sm.afull + i, TC_WARPS * TARRIVALS); tbar_init(sm.lfull + i, 1); tbar_init(sm.ofree + i, TC_WARPS * TARRIVALS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TM_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // K / V rings start as zeros: the pad columns (KVC .. TKP-1) are never written and must stay finite (P = 0 there)
    for (int i = tid; i < (int)(2 * Cf::KV_BYTES / 16); i += TTHREADS) reinterpret_cast<uint4*>(pK)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < (int)(Cf::ONES_BYTES / 4); i += TTHREADS) reinterpret_cast<uint32_t*>(pOnes)[i] = 0x3C003C00u;   // f16 ones
    // classifier weights [TNCLS][64] f16, K-major SWIZZLE_128B rows (16-byte chunk c of row j at (c ^ (j & 7)) << 4)
    for (int i = tid; i < TNCLS * MC; i += TTHREADS) {
        const int j = i / MC, c = i % MC;
        const float v = (p.wcls && j < p.ncls) ? clamp_h(__ldg(p.wcls + (size_t)j * MC + c)) : 0.f;
        *reinterpret_cast<__half*>(pW + j * 128 + ((((c >> 3) ^ j) & 7) << 4) + (c & 7) * 2) = __float2half_rn(v);
    }
    if (tid < TNCLS) sm.s_bc[tid] = (p.wcls && p.bcls && tid < p.ncls) ? __ldg(p.bcls + tid) : 0.f;
    t_fence_async_smem();
    t_fence_before();
    __syncthreads();
    t_fence_after();
    const uint32_t tmem = *tmem_slot;

#ifdef ARSEG_TTRACE
    const long long t_role = clock64();
#endif
    if (warp >= TS_W0) t_s_role<K>(sm, tmem, S, p.dbg);
    else if (warp >= TE_W0) t_e_role<K, NCP>(p, sm, tmem, n, x0, ya, yb, S);
    else if (warp == TM_WARP) t_m_role<K, NCP>(sm, tmem, S, p.dbg);
    else if (warp < TG_WARP0) t_d_role<K>(p, sm, x0, ya, NH, S);
    else t_g_role<K>(p, sm, rec, warped, n, x0, ya, NH, S);
#ifdef ARSEG_TTRACE
    if (blockIdx.x == TTRACE_CTA && (tid & 31) == 0) { g_ttrace[warp * 16 + 15] += clock64() - t_role; g_ttrace[warp * 16 + 14] += S; }
#endif

    t_fence_before();
    __syncthreads();
    if (warp == TM_WARP) {
        t_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

template <int K, int NCP>
static int creff_tc_launch_n(CreffMmaParams& p, int hr_dtype, int phase, void* ws, size_t ws_bytes, cudaStream_t st) {
    using Cf = TCfg<K>;
    auto kern = creff_tc_kernel<K, NCP>;
    // per device, written once per process: an idempotent attribute, so the unsynchronised flag is a benign race between
    // nn.DataParallel worker threads
    static bool configured[64] = {false};
    int dev = 0;
    ARSEG_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        ARSEG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cf::SMEM));
        // keep what is left of the 228 KB as L1: the gather's memory-level parallelism is bounded by the L1 lines its
        // outstanding misses can allocate
        ARSEG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)((Cf::SMEM + 1024) * 100 / (228 * 1024)) + 1));
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    ARSEG_REQUIRE(ws && ws_bytes >= creff_tc_workspace_bytes(p.N, p.H, p.W, K) && ((uintptr_t)ws % 128) == 0,
                  "creff_tc: needs a 128-byte aligned workspace of %zu bytes (arseg_creff_workspace_bytes)", creff_tc_workspace_bytes(p.N, p.H, p.W, K));
    uint4* rec = reinterpret_cast<uint4*>(ws);
    char* warped = reinterpret_cast<char*>(ws) + t_rec_bytes(p.H, p.W);
    if (phase != ARSEG_CREFF_PHASE_MAIN) {
        // pre-pass: depends on the keyframe feature and the MV fields only (not on the LR feature)
        creff_tc_rec_kernel<<<(unsigned)ceil_div_ll((long long)p.H * p.W, 256), 256, 0, st>>>(p, rec);
        ARSEG_CHECK_LAUNCH("creff_tc_rec");
        int TWB = TWB_DEFAULT;
        { const char* e = getenv("ARSEG_TC_TWB"); if (e && atoi(e) > 0) TWB = atoi(e); }
        const long long warps = (long long)ceil_div(t_hp<K>(p.H), TWB) * p.N * ceil_div(TWB * t_wp<K>(p.W), 32);
        ARSEG_REQUIRE((long long)p.N * p.H * p.W < 0xffffffffLL && ceil_div_ll(warps, TWARP_THREADS / 32) < 2147483647LL, "creff_tc: too many pixels");
        if (hr_dtype == ARSEG_F16) creff_tc_warp_kernel<K, false><<<(unsigned)ceil_div_ll(warps, TWARP_THREADS / 32), TWARP_THREADS, 0, st>>>(p, reinterpret_cast<uint4*>(warped), TWB);
        else creff_tc_warp_kernel<K, true><<<(unsigned)ceil_div_ll(warps, TWARP_THREADS / 32), TWARP_THREADS, 0, st>>>(p, reinterpret_cast<uint4*>(warped), TWB);
        ARSEG_CHECK_LAUNCH("creff_tc_warp");
        if (phase == ARSEG_CREFF_PHASE_PREPASS) return ARSEG_OK;
    }
    p.ncols = ceil_div(p.W, TSW);
    { const char* d = getenv("ARSEG_CREFF_DBG"); p.dbg = d ? atoi(d) : 0; }      // -DARSEG_TTRACE builds only
    // row segments: enough CTAs for >= ~6 waves of one CTA per SM, but segments of >= 48 rows (each segment pays
    // ~K+5 redundant halo rows and the pipeline fill)
    const int sms = sm_count() > 0 ? sm_count() : 148;
    int nseg = 1;
    const char* e = getenv("ARSEG_CREFF_SEG_ROWS");
    if (e && atoi(e) >= 8) nseg = ceil_div(p.H, (atoi(e) + 7) / 8 * 8);
    else while ((long long)p.N * p.ncols * nseg < 6LL * sms && ceil_div(p.H, nseg + 1) >= 48) ++nseg;
    p.seg_rows = (ceil_div(p.H, nseg) + 7) / 8 * 8;
    p.nseg = ceil_div(p.H, p.seg_rows);
    const long long blocks = (long long)p.N * p.ncols * p.nseg;
    ARSEG_REQUIRE(blocks > 0 && blocks < 2147483647LL, "creff_tc: grid too large");
    kern<<<(unsigned)blocks, TTHREADS, Cf::SMEM, st>>>(p, rec, warped);
    ARSEG_CHECK_LAUNCH("creff_tc");
    return ARSEG_OK;
}

template <int K>
static int creff_tc_launch_k(CreffMmaParams& p, int hr_dtype, int phase, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (!p.wcls) return creff_tc_launch_n<K, 0>(p, hr_dtype, phase, ws, ws_bytes, st);
    if (p.ncls <= 16) return creff_tc_launch_n<K, 16>(p, hr_dtype, phase, ws, ws_bytes, st);
    return creff_tc_launch_n<K, 32>(p, hr_dtype, phase, ws, ws_bytes, st);
}

// hr NHWC fp32 or f16 ([.,H,W,64]), lr NHWC f16 ([N,h,w,64]); k in {3, 5, 7}; ws = lr_up gather records + warped keyframe rows
// (creff_tc_workspace_bytes); phase: ARSEG_CREFF_PHASE_* (pre-pass and attention kernel together or separately)
int creff_tc_launch(CreffMmaParams& p, int k, int hr_dtype, int phase, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (p.H < 2 || p.W < 2 || p.h < 2 || p.w < 2) ARSEG_UNSUPPORTED("creff_tc: maps must be at least 2x2 (hr %dx%d, lr %dx%d)", p.H, p.W, p.h, p.w);
    ARSEG_REQUIRE((long long)p.H * p.W < (1LL << 31) && (long long)p.h * p.w < (1LL << 31), "creff_tc: map too large");
    switch (k) {
        case 3: return creff_tc_launch_k<3>(p, hr_dtype, phase, ws, ws_bytes, st);
        case 5: return creff_tc_launch_k<5>(p, hr_dtype, phase, ws, ws_bytes, st);
        case 7: return creff_tc_launch_k<7>(p, hr_dtype, phase, ws, ws_bytes, st);
        default: ARSEG_UNSUPPORTED("creff_tc: window k=%d (3, 5, 7)", k);
    }
}

}  // namespace arseg
#ifdef ARSEG_TTRACE
extern "C" int arseg_debug_creff_tc_events(long long* host) {
    cudaMemcpyFromSymbol(host, arseg::g_tev, sizeof(long long) * 3 * 64 * 12);
    return 3 * 64 * 12;
}
extern "C" int arseg_debug_creff_tc_trace(long long* host, int reset) {
    cudaMemcpyFromSymbol(host, arseg::g_ttrace, sizeof(long long) * 512);
    if (reset) { long long z[512] = {0}; cudaMemcpyToSymbol(arseg::g_ttrace, z, sizeof(z)); }
    return 512;
}
#endif
