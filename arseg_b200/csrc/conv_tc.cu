// conv_tc.cu -- implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// GEMM view (stride-1 convs, any dilation / padding):
//   D[128 output pixels, BLOCK_N couts] += sum over taps (ky,kx) and Cin chunks  A_tap[128, BK] * W_tap[BLOCK_N, BK]^T
//   * M tile = a TH x TW spatial patch of one image (TH*TW = 128).  For every tap the A operand is ONE 4-D
//     TMA box load {BK channels, TW, TH, 1} from the NHWC activation at the tap-shifted coordinate; the TMA
//     unit zero-fills out-of-image elements, which implements the conv padding with no im2col buffer.
//     The box lands in shared memory as 128 rows x 128 bytes with the 128B swizzle = the canonical K-major
//     UMMA operand layout (8-row atoms, SBO = 1024 B).
//   * B operand = weights [Cout][KH*KW*Cin] (K contiguous), 2-D TMA box {BK, BLOCK_N}, same layout.
//   * Accumulator: 128 lanes x BLOCK_N fp32 columns of TMEM; one elected thread issues tcgen05.mma
//     (kind::f16 for bf16 operands, kind::tf32 for fp32 operands), tcgen05.commit releases smem stages
//     and signals the epilogue through mbarriers.
//   * Epilogue (4 warps, one TMEM lane quarter each): tcgen05.ld -> folded-BN scale/shift (+bias)
//     -> (+residual) -> ReLU/PReLU -> NHWC store (channel slice of the destination = fused torch.cat).
//   * Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue.
//   Persistent: one CTA per SM walks the (m-tile, n-tile) list; two TMEM accumulator buffers let the epilogue of
//   one tile overlap the main loop of the next.
#include "common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <mutex>
#include <cstdlib>

namespace arseg {

constexpr int TC_THREADS = 192;
constexpr int TC_BM = 128;
constexpr int TC_ROW_BYTES = 128;           // one swizzle row = BK elements
constexpr int TC_A_STAGE = TC_BM * TC_ROW_BYTES;  // 16 KB

struct TcParams {
    const float* scale; const float* shift; const void* res; void* out;
    int N, Ho, Wo, Cin, Cout, KH, KW, pad, dil, ocs, oco, act;
    float slope;
    int TH, TW, tiles_x, tiles_y, n_tiles, kchunks;  // kchunks = Cin / BK
    int stride;                                      // 1 or 2 (tap-box kernel only: TMA element strides)
    int cout_pad;                                    // Cout rounded up to the n-tile size (scale/shift staging)
    int out_f32;                                     // fp32 output from 16-bit operands
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a mis-programmed descriptor must not hang the GPU box -- trap after ~4e9 cycles.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {
            printf("arseg conv_tc: mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4 | [32,46) stride byte offset >> 4 (8 rows * 128 B = 1024) | [46,48) version = 1
//   | [61,64) layout type = 2 (SWIZZLE_128B).  LBO is unused for swizzled K-major operands.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) @4, a/b format @7/@10
// (BF16 = 1, TF32 = 2), K-major A and B, N>>3 @17, M>>4 @24.
__host__ __device__ constexpr uint32_t umma_idesc(int fmt, int M, int N) {
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// cute::UMMA::F16F32Format: F16 = 0, BF16 = 1, TF32 = 2
template <typename T> struct UmmaFmt;
template <> struct UmmaFmt<float> { static constexpr int v = 2; };
template <> struct UmmaFmt<__nv_bfloat16> { static constexpr int v = 1; };
template <> struct UmmaFmt<__half> { static constexpr int v = 0; };

template <bool TF32>
__device__ __forceinline__ void umma_ss(uint64_t da, uint64_t db, uint32_t tmem_d, uint32_t accumulate, uint32_t idesc) {
    if (TF32)
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// eight fp32 -> eight 16-bit values (bf16: round to nearest; fp16: round to nearest, saturate)
template <typename T> __device__ __forceinline__ uint4 pack8(const float* v);
template <> __device__ __forceinline__ uint4 pack8<__nv_bfloat16>(const float* v) {
    uint4 pk;
    __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
    pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
    pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
    return pk;
}
template <> __device__ __forceinline__ uint4 pack8<__half>(const float* v) {
    uint4 pk;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(pk.x) : "f"(v[1]), "f"(v[0]));
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(pk.y) : "f"(v[3]), "f"(v[2]));
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(pk.z) : "f"(v[5]), "f"(v[4]));
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(pk.w) : "f"(v[7]), "f"(v[6]));
    return pk;
}
template <> __device__ __forceinline__ uint4 pack8<float>(const float*) { return make_uint4(0, 0, 0, 0); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ---------------------------------------------------------------------------------------------
// kernel: persistent, warp-specialised.  One CTA per SM walks the tile list (tile = blockIdx.x + i * gridDim.x,
// n-tile fastest so that CTAs working at the same time share A tiles in L2).  Three pipelines:
//   smem ring   full/empty   TMA producer (warp 0)  <->  MMA issuer (warp 1)
//   TMEM        tfull/tempty MMA issuer             <->  epilogue warps 2..5 (two accumulator buffers, so the
//                                                        epilogue of tile i overlaps the main loop of tile i+1)
// ---------------------------------------------------------------------------------------------
constexpr int TC_NACC = 2;
constexpr int TC_MAX_COUT = 2048;   // scale/shift staged in shared memory once per CTA

template <typename T, int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                const __grid_constant__ CUtensorMap map_b, TcParams p) {
    constexpr bool TF32 = sizeof(T) == 4;
    constexpr int BK = TC_ROW_BYTES / (int)sizeof(T);     // elements per swizzle row: 64 (bf16) / 32 (tf32)
    constexpr int UMMA_K = 32 / (int)sizeof(T);           // 16 / 8
    constexpr int B_STAGE = BLOCK_N * TC_ROW_BYTES;
    constexpr uint32_t STAGE_BYTES = TC_A_STAGE + B_STAGE;
    constexpr uint32_t IDESC = umma_idesc(UmmaFmt<T>::v, TC_BM, BLOCK_N);
    constexpr int ACC_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;
    constexpr int TMEM_COLS = TC_NACC * ACC_COLS;         // power of two >= 64

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * TC_A_STAGE;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + TC_NACC;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + TC_NACC);
    float* s_scale = reinterpret_cast<float*>(tmem_slot + 4);
    float* s_shift = s_scale + p.cout_pad;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < TC_NACC; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 128); }
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < p.cout_pad; i += TC_THREADS) {
        s_scale[i] = (p.scale && i < p.Cout) ? __ldg(p.scale + i) : 1.f;
        s_shift[i] = (p.shift && i < p.Cout) ? __ldg(p.shift + i) : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int taps = p.KH * p.KW;
    const int kiters = taps * p.kchunks;
    const int total_tiles = p.N * p.tiles_y * p.tiles_x * p.n_tiles;

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                int b = tile;
                const int nt = b % p.n_tiles; b /= p.n_tiles;
                const int txi = b % p.tiles_x; b /= p.tiles_x;
                const int tyi = b % p.tiles_y;
                const int img = b / p.tiles_y;
                const int x0 = txi * p.TW, y0 = tyi * p.TH, n0 = nt * BLOCK_N;
                int tap = 0, kc = 0, ky = 0, kx = 0;
                for (int it = 0; it < kiters; ++it) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
                    tma_load_4d(&map_a, &full_bar[stage], smem_a + stage * TC_A_STAGE, kc * BK, x0 * p.stride - p.pad + kx * p.dil,
                                y0 * p.stride - p.pad + ky * p.dil, img);
                    tma_load_2d(&map_b, &full_bar[stage], smem_b + stage * B_STAGE, tap * p.Cin + kc * BK, n0);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    if (++kc == p.kchunks) { kc = 0; ++tap; if (++kx == p.KW) { kx = 0; ++ky; } }
                }
            }
        }
    } else if (warp == 1) {
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);              // the epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(acc * ACC_COLS);
            for (int it = 0; it < kiters; ++it) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t da = umma_desc_sw128(smem_u32(smem_a + stage * TC_A_STAGE));
                    const uint64_t db = umma_desc_sw128(smem_u32(smem_b + stage * B_STAGE));
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)   // +32 bytes per UMMA_K step inside the swizzle row
                        umma_ss<TF32>(da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), tmem_d, (it | k) != 0, IDESC);
                    umma_commit(&empty_bar[stage]);                       // frees the smem stage when the MMAs retire
                    if (it == kiters - 1) umma_commit(&tfull_bar[acc]);   // accumulator complete -> epilogue
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == TC_NACC) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // epilogue warps 2..5: TMEM lane quarter = warp % 4
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const int ry = row / p.TW, rx = row % p.TW;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int b = tile;
            const int nt = b % p.n_tiles; b /= p.n_tiles;
            const int txi = b % p.tiles_x; b /= p.tiles_x;
            const int tyi = b % p.tiles_y;
            const int img = b / p.tiles_y;
            const int n0 = nt * BLOCK_N;
            const int gy = tyi * p.TH + ry, gx = txi * p.TW + rx;
            const bool valid = gy < p.Ho && gx < p.Wo;
            const size_t pix = ((size_t)img * p.Ho + gy) * p.Wo + gx;
            T* __restrict__ out = reinterpret_cast<T*>(p.out) + pix * p.ocs + p.oco;
            float* __restrict__ outf = reinterpret_cast<float*>(p.out) + pix * p.ocs + p.oco;   // p.out_f32
            const T* __restrict__ res = p.res ? reinterpret_cast<const T*>(p.res) + pix * p.Cout : nullptr;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * ACC_COLS);
#pragma unroll 1
            for (int cb = 0; cb < BLOCK_N; cb += 32) {
                const int co0 = n0 + cb;
                // residual as 16-byte loads issued before the accumulator read-back (see conv_tc_halo_kernel)
                constexpr int RV = 16 / (int)sizeof(T), RN = 32 / RV;
                uint4 rr[RN];
                const bool rvec = res != nullptr && valid && co0 + 32 <= p.Cout && p.Cout % RV == 0;
                if (rvec) {
#pragma unroll
                    for (int q = 0; q < RN; ++q) rr[q] = __ldg(reinterpret_cast<const uint4*>(res + co0) + q);
                }
                uint32_t r[32];
                tmem_ld32(tmem_d + (uint32_t)cb, r);
                tmem_ld_wait();
                if (cb + 32 >= BLOCK_N) {                 // last chunk is in registers: hand the accumulator back
                    tc_fence_before();
                    mbar_arrive(&tempty_bar[acc]);
                }
                if (!valid || co0 >= p.Cout) continue;
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 sc = *reinterpret_cast<const float4*>(s_scale + co0 + j);
                    const float4 sh = *reinterpret_cast<const float4*>(s_shift + co0 + j);
                    v[j] = fmaf(__uint_as_float(r[j]), sc.x, sh.x); v[j + 1] = fmaf(__uint_as_float(r[j + 1]), sc.y, sh.y);
                    v[j + 2] = fmaf(__uint_as_float(r[j + 2]), sc.z, sh.z); v[j + 3] = fmaf(__uint_as_float(r[j + 3]), sc.w, sh.w);
                }
                if (rvec) {
                    const T* rt = reinterpret_cast<const T*>(rr);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += to_f32(rt[j]);
                } else if (res) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (co0 + j < p.Cout) v[j] += to_f32(res[co0 + j]);
                }
                if (p.act == ARSEG_ACT_RELU) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                } else if (p.act == ARSEG_ACT_PRELU) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * p.slope;
                }
                if (!TF32 && p.out_f32) {
                    if (co0 + 32 <= p.Cout && ((p.ocs | p.oco) % 4 == 0)) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(outf + co0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (co0 + j < p.Cout) outf[co0 + j] = v[j];
                    }
                } else if (co0 + 32 <= p.Cout && ((p.ocs | p.oco) % (16 / (int)sizeof(T)) == 0)) {
                    if (TF32) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + co0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; j += 8)
                            *reinterpret_cast<uint4*>(reinterpret_cast<T*>(out) + co0 + j) = pack8<T>(v + j);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (co0 + j < p.Cout) out[co0 + j] = from_f32<T>(v[j]);
                }
            }
            if (++acc == TC_NACC) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// 3x3 kernel with operand reuse across the taps ("halo" kernel).  Same roles and pipelines as above, but the A
// operand of a Cin chunk is loaded ONCE per tile as a (16+2d) x 16-pixel halo box (d = dilation = padding) instead
// of nine tap-shifted 128-pixel boxes: the output tile is 16 rows x 8 pixels, so the 128 GEMM rows of tap (ky,kx)
// are the halo rows ((ty + ky d) * 16 + tx + kx d), i.e. 8-row groups at a constant stride of 16 halo rows --
// a K-major SWIZZLE_128B operand with SBO = 2048 B whose start address is simply advanced by the tap offset.
// A traffic drops from 9 x 16 KB to 36-48 KB per Cin chunk; B (weights) is streamed per tap as before.
// The 128-byte swizzle is a function of the shared-memory address bits in both the TMA write and the UMMA read,
// so a start address that is 128-byte but not 1024-byte aligned addresses the same bytes with the descriptor's
// base-offset field left 0 (verified on B200: setting it to (addr >> 7) & 7 gives wrong results).
// ---------------------------------------------------------------------------------------------
constexpr int TCH_TW = 8, TCH_TH = 16, TCH_WH = 16;      // output tile, halo box width (pixels)

struct TchParams {
    TcParams c;
    int sa, sb;             // A-halo stages, B stages
    int a_stage;            // bytes per A-halo stage
    int tps;                // taps per B stage (1 or 3): short MMAs (narrow n-tiles) need fewer barrier round trips
    int b_resident;         // 1: all 9 * kchunks weight tiles stay in shared memory for the whole kernel
};

__device__ __forceinline__ uint64_t umma_desc_sw128_sbo(uint32_t smem_addr, uint32_t sbo, uint32_t base_off) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | ((uint64_t)(base_off & 7) << 49) | (2ull << 61);
}

constexpr int TCH_EPI_WARPS = 8;                       // two warps per TMEM lane quarter, half the columns each
constexpr int TCH_THREADS = 64 + 32 * TCH_EPI_WARPS;

template <typename T, int BLOCK_N>
__global__ void __launch_bounds__(TCH_THREADS, 1) conv_tc_halo_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                      const __grid_constant__ CUtensorMap map_b, TchParams hp) {
    const TcParams& p = hp.c;
    constexpr bool TF32 = sizeof(T) == 4;
    constexpr int BK = TC_ROW_BYTES / (int)sizeof(T);
    constexpr int UMMA_K = 32 / (int)sizeof(T);
    constexpr int B_TAP = BLOCK_N * TC_ROW_BYTES;         // bytes of one tap's weight tile
    constexpr uint32_t IDESC = umma_idesc(UmmaFmt<T>::v, TC_BM, BLOCK_N);
    constexpr int ACC_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;
    constexpr int TMEM_COLS = TC_NACC * ACC_COLS;
    constexpr int MAXS = 8;
    constexpr int CHUNKS = BLOCK_N / 32;                  // 32-column epilogue chunks
    constexpr int CPW = CHUNKS >= 2 ? CHUNKS / 2 : 1;     // chunks per epilogue warp

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int b_stage = hp.tps * B_TAP;
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + hp.sa * hp.a_stage;
    uint64_t* fulla = reinterpret_cast<uint64_t*>(smem_b + hp.sb * b_stage);
    uint64_t* emptya = fulla + MAXS;
    uint64_t* fullb = emptya + MAXS;
    uint64_t* emptyb = fullb + MAXS;
    uint64_t* tfull_bar = emptyb + MAXS;
    uint64_t* tempty_bar = tfull_bar + TC_NACC;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + TC_NACC);
    float* s_scale = reinterpret_cast<float*>(tmem_slot + 4);
    float* s_shift = s_scale + p.cout_pad;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int EPI_ARRIVALS = (CHUNKS >= 2 ? TCH_EPI_WARPS : 4) * 32;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        for (int s = 0; s < MAXS; ++s) { mbar_init(&fulla[s], 1); mbar_init(&emptya[s], 1); mbar_init(&fullb[s], 1); mbar_init(&emptyb[s], 1); }
        for (int a = 0; a < TC_NACC; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], EPI_ARRIVALS); }
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < p.cout_pad; i += TCH_THREADS) {
        s_scale[i] = (p.scale && i < p.Cout) ? __ldg(p.scale + i) : 1.f;
        s_shift[i] = (p.shift && i < p.Cout) ? __ldg(p.shift + i) : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int total_tiles = p.N * p.tiles_y * p.tiles_x * p.n_tiles;
    const int d = p.dil;
    const int tgroups = 9 / hp.tps;                       // B stages per Cin chunk

    if (warp == 0) {
        if (elect_one()) {
            int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
            bool first = true;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, first = false) {
                int b = tile;
                const int nt = b % p.n_tiles; b /= p.n_tiles;
                const int txi = b % p.tiles_x; b /= p.tiles_x;
                const int tyi = b % p.tiles_y;
                const int img = b / p.tiles_y;
                const int x0 = txi * TCH_TW, y0 = tyi * TCH_TH, n0 = nt * BLOCK_N;
                for (int kc = 0; kc < p.kchunks; ++kc) {
                    mbar_wait(&emptya[sa], pa ^ 1);
                    mbar_expect_tx(&fulla[sa], (uint32_t)hp.a_stage);
                    tma_load_4d(&map_a, &fulla[sa], smem_a + sa * hp.a_stage, kc * BK, x0 - d, y0 - d, img);
                    if (++sa == hp.sa) { sa = 0; pa ^= 1; }
                    if (hp.b_resident && !first) continue;
                    // resident: stage index = kc * tgroups + tg; barrier 0 is armed once with the bytes of all stages
                    if (hp.b_resident && kc == 0) mbar_expect_tx(&fullb[0], (uint32_t)(p.kchunks * tgroups * b_stage));
                    for (int tg = 0; tg < tgroups; ++tg) {
                        uint64_t* fb = hp.b_resident ? &fullb[0] : &fullb[sb];
                        if (!hp.b_resident) {
                            mbar_wait(&emptyb[sb], pb ^ 1);
                            mbar_expect_tx(fb, (uint32_t)b_stage);
                        }
                        for (int i = 0; i < hp.tps; ++i)
                            tma_load_2d(&map_b, fb, smem_b + sb * b_stage + i * B_TAP, (tg * hp.tps + i) * p.Cin + kc * BK, n0);
                        if (++sb == hp.sb) { sb = 0; pb ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
        int acc = 0; uint32_t acc_phase = 0;
        bool first = true;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, first = false) {
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(acc * ACC_COLS);
            if (hp.b_resident) { sb = 0; if (first) { /* every weight tile has landed */ mbar_wait(&fullb[0], 0); } }
            for (int kc = 0; kc < p.kchunks; ++kc) {
                mbar_wait(&fulla[sa], pa);
                const uint32_t a_base = smem_u32(smem_a + sa * hp.a_stage);
                for (int tg = 0; tg < tgroups; ++tg) {
                    if (!hp.b_resident) mbar_wait(&fullb[sb], pb);
                    tc_fence_after();
                    if (elect_one()) {
                        for (int i = 0; i < hp.tps; ++i) {
                            const int tap = tg * hp.tps + i;
                            const int ky = tap / 3, kx = tap - 3 * ky;
                            const uint32_t a_addr = a_base + (uint32_t)(((ky * d) * TCH_WH + kx * d) * TC_ROW_BYTES);
                            const uint64_t da = umma_desc_sw128_sbo(a_addr, TCH_WH * TC_ROW_BYTES, 0u);
                            const uint64_t db = umma_desc_sw128(smem_u32(smem_b + sb * b_stage + i * B_TAP));
#pragma unroll
                            for (int k = 0; k < BK / UMMA_K; ++k)
                                umma_ss<TF32>(da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), tmem_d, (kc | tap | k) != 0, IDESC);
                        }
                        if (!hp.b_resident) umma_commit(&emptyb[sb]);
                        if (tg == tgroups - 1) {
                            umma_commit(&emptya[sa]);
                            if (kc == p.kchunks - 1) umma_commit(&tfull_bar[acc]);
                        }
                    }
                    __syncwarp();
                    if (++sb == hp.sb) { sb = 0; pb ^= 1; }
                }
                if (++sa == hp.sa) { sa = 0; pa ^= 1; }
            }
            if (++acc == TC_NACC) { acc = 0; acc_phase ^= 1; }
        }
    } else if (CHUNKS >= 2 || warp < 6) {
        // epilogue warps 2..9: TMEM lane quarter = warp % 4; warps 2..5 take the low column half, 6..9 the high half
        const int quarter = warp & 3;
        const int chalf = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        const int ry = row / TCH_TW, rx = row % TCH_TW;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int b = tile;
            const int nt = b % p.n_tiles; b /= p.n_tiles;
            const int txi = b % p.tiles_x; b /= p.tiles_x;
            const int tyi = b % p.tiles_y;
            const int img = b / p.tiles_y;
            const int n0 = nt * BLOCK_N;
            const int gy = tyi * TCH_TH + ry, gx = txi * TCH_TW + rx;
            const bool valid = gy < p.Ho && gx < p.Wo;
            const size_t pix = ((size_t)img * p.Ho + gy) * p.Wo + gx;
            T* __restrict__ out = reinterpret_cast<T*>(p.out) + pix * p.ocs + p.oco;
            float* __restrict__ outf = reinterpret_cast<float*>(p.out) + pix * p.ocs + p.oco;   // p.out_f32
            const T* __restrict__ res = p.res ? reinterpret_cast<const T*>(p.res) + pix * p.Cout : nullptr;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * ACC_COLS);
#pragma unroll 1
            for (int ci = 0; ci < CPW; ++ci) {
                const int cb = (chalf * CPW + ci) * 32;
                const int co0 = n0 + cb;
                // residual of this chunk: 16-byte loads issued BEFORE the accumulator is read back, so their latency hides
                // behind the TMEM load (32 scalar 2-byte loads after it cost BasicBlock.conv2 layers ~30 % of their time)
                constexpr int RV = 16 / (int)sizeof(T), RN = 32 / RV;
                uint4 rr[RN];
                const bool rvec = res != nullptr && valid && co0 + 32 <= p.Cout && p.Cout % RV == 0;
                if (rvec) {
#pragma unroll
                    for (int q = 0; q < RN; ++q) rr[q] = __ldg(reinterpret_cast<const uint4*>(res + co0) + q);
                }
                uint32_t r[32];
                tmem_ld32(tmem_d + (uint32_t)cb, r);
                tmem_ld_wait();
                if (ci == CPW - 1) {                      // this warp's last chunk is in registers
                    tc_fence_before();
                    mbar_arrive(&tempty_bar[acc]);
                }
                if (!valid || co0 >= p.Cout) continue;
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 sc = *reinterpret_cast<const float4*>(s_scale + co0 + j);
                    const float4 sh = *reinterpret_cast<const float4*>(s_shift + co0 + j);
                    v[j] = fmaf(__uint_as_float(r[j]), sc.x, sh.x); v[j + 1] = fmaf(__uint_as_float(r[j + 1]), sc.y, sh.y);
                    v[j + 2] = fmaf(__uint_as_float(r[j + 2]), sc.z, sh.z); v[j + 3] = fmaf(__uint_as_float(r[j + 3]), sc.w, sh.w);
                }
                const bool vec = co0 + 32 <= p.Cout && ((p.ocs | p.oco) % (16 / (int)sizeof(T)) == 0);
                if (rvec) {
                    const T* rt = reinterpret_cast<const T*>(rr);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += to_f32(rt[j]);
                } else if (res) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (co0 + j < p.Cout) v[j] += to_f32(res[co0 + j]);
                }
                if (p.act == ARSEG_ACT_RELU) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                } else if (p.act == ARSEG_ACT_PRELU) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * p.slope;
                }
                if (!TF32 && p.out_f32) {
                    if (co0 + 32 <= p.Cout && ((p.ocs | p.oco) % 4 == 0)) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(outf + co0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (co0 + j < p.Cout) outf[co0 + j] = v[j];
                    }
                } else if (vec) {
                    if (TF32) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + co0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; j += 8)
                            *reinterpret_cast<uint4*>(reinterpret_cast<T*>(out) + co0 + j) = pack8<T>(v + j);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (co0 + j < p.Cout) out[co0 + j] = from_f32<T>(v[j]);
                }
            }
            if (++acc == TC_NACC) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    });
    return fn;
}

static bool tf32_tma_round() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("ARSEG_TF32_TMA_ROUND");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

static void pick_tile(int Ho, int Wo, int& TH, int& TW) {
    long long best = -1;
    for (int tw = 8; tw <= 128; tw *= 2) {
        const int th = 128 / tw;
        const long long tiles = (long long)ceil_div(Ho, th) * ceil_div(Wo, tw);
        // fewest tiles; tie -> squarer tile (less halo traffic)
        if (best < 0 || tiles < best || (tiles == best && abs(th - tw) < abs(TH - TW))) { best = tiles; TH = th; TW = tw; }
    }
}

bool conv_tc_supported(const arseg_conv_desc* d) {
    const int es = d->engine == ARSEG_CONV_TC_TF32 ? 4 : 2;
    const int bk = 128 / es;
    if (d->engine == ARSEG_CONV_TC_TF32 && d->dtype != ARSEG_F32) return false;
    if (d->engine == ARSEG_CONV_TC_BF16 && d->dtype != ARSEG_BF16) return false;
    if (d->engine == ARSEG_CONV_TC_F16 && d->dtype != ARSEG_F16) return false;
    if (d->stride != 1 && d->stride != 2) return false;
    if (d->Cin % bk != 0) return false;
    if (d->Cout < 16) return false;
    // "same" convs (every stride-1 conv on the path) and their stride-2 versions (ResNet down-sampling blocks)
    if (2 * d->pad != d->dil * (d->KH - 1) || 2 * d->pad != d->dil * (d->KW - 1)) return false;
    if ((uintptr_t)d->in % 16 || (uintptr_t)d->w % 16) return false;
    return true;
}

template <typename T, int BLOCK_N, int STAGES>
static int launch_tc(const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& p, cudaStream_t st) {
    const size_t smem = (size_t)STAGES * (TC_A_STAGE + BLOCK_N * TC_ROW_BYTES) + (2 * STAGES + 2 * TC_NACC) * 8 + 16 + 2 * (size_t)p.cout_pad * 4 + 1024;
    ARSEG_REQUIRE(smem <= 232448, "conv_tc: Cout=%d too large for the scale/shift staging", p.Cout);
    auto kern = conv_tc_kernel<T, BLOCK_N, STAGES>;
    // per device, written once per process: an idempotent attribute, so the unsynchronised flag is a benign race between
    // nn.DataParallel worker threads (the worst case is a redundant cudaFuncSetAttribute)
    static bool configured[64] = {false};
    int dev = 0;
    ARSEG_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        ARSEG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    const long long tiles = (long long)p.N * p.tiles_y * p.tiles_x * p.n_tiles;
    ARSEG_REQUIRE(tiles > 0 && tiles < 2147483647LL, "conv_tc: too many tiles");
    const int sms = sm_count() > 0 ? sm_count() : 148;
    const unsigned blocks = (unsigned)(tiles < sms ? tiles : sms);          // persistent: one CTA per SM
    kern<<<blocks, TC_THREADS, smem, st>>>(ma, mb, p);
    ARSEG_CHECK_LAUNCH("conv_tc");
    return ARSEG_OK;
}

static int env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }

template <typename T, int BLOCK_N>
static int launch_tch(const CUtensorMap& ma, const CUtensorMap& mb, TchParams& hp, cudaStream_t st) {
    const TcParams& p = hp.c;
    const size_t fixed = (4 * 8 + 2 * TC_NACC) * 8 + 16 + 2 * (size_t)p.cout_pad * 4 + 1024;
    const size_t b_tap = (size_t)BLOCK_N * TC_ROW_BYTES;
    const long long budget = 232448 - (long long)fixed;
    hp.sa = 2;
    // weights that fit next to two A stages stay resident (64 -> 64 layers: 147 KB in tf32)
    const long long w_bytes = 9LL * p.kchunks * (long long)b_tap;
    hp.b_resident = p.n_tiles == 1 && w_bytes + 2LL * hp.a_stage <= budget && env_int("ARSEG_TC_BRES", 1) != 0;
    hp.tps = (BLOCK_N <= 64 && env_int("ARSEG_TC_TPS", 3) == 3) ? 3 : 1;
    const size_t b_stage = hp.tps * b_tap;
    if (hp.b_resident) hp.sb = (9 / hp.tps) * p.kchunks;
    else {
        long long sb = (budget - (long long)hp.sa * hp.a_stage) / (long long)b_stage;
        ARSEG_REQUIRE(sb >= 2, "conv_tc_halo: not enough shared memory (Cout=%d)", p.Cout);
        hp.sb = sb > 8 ? 8 : (int)sb;
    }
    const size_t smem = fixed + (size_t)hp.sa * hp.a_stage + (size_t)hp.sb * b_stage;
    auto kern = conv_tc_halo_kernel<T, BLOCK_N>;
    static bool configured[64] = {false};     // benign race, see launch_tc
    int dev = 0;
    ARSEG_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        ARSEG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    const long long tiles = (long long)p.N * p.tiles_y * p.tiles_x * p.n_tiles;
    ARSEG_REQUIRE(tiles > 0 && tiles < 2147483647LL, "conv_tc_halo: too many tiles");
    const int sms = sm_count() > 0 ? sm_count() : 148;
    const unsigned blocks = (unsigned)(tiles < sms ? tiles : sms);
    kern<<<blocks, TCH_THREADS, smem, st>>>(ma, mb, hp);
    ARSEG_CHECK_LAUNCH("conv_tc_halo");
    return ARSEG_OK;
}

// ARSEG_TC_HALO=0 disables the halo kernel, ARSEG_TC_N256=0 the 256-wide n-tiles, ARSEG_TC_BRES=0 resident weights,
// ARSEG_TC_TPS=1 the three-tap weight stages (A/B measurements)

int conv_tc_launch(const arseg_conv_desc* d, cudaStream_t st) {
    EncodeTiledFn encode = get_encode();
    if (!encode) { set_error("conv_tc: cuTensorMapEncodeTiled entry point not available"); return ARSEG_E_CUDA; }
    const bool tf32 = d->engine == ARSEG_CONV_TC_TF32;
    const int es = tf32 ? 4 : 2, bk = 128 / es;
    TcParams p;
    p.scale = d->scale; p.shift = d->shift; p.res = d->residual; p.out = d->out;
    p.stride = d->stride;
    p.N = d->N; p.Ho = (d->Hi - 1) / d->stride + 1; p.Wo = (d->Wi - 1) / d->stride + 1; p.Cin = d->Cin; p.Cout = d->Cout; p.KH = d->KH; p.KW = d->KW;
    p.out_f32 = (d->out_f32 && d->dtype != ARSEG_F32) ? 1 : 0;
    p.pad = d->pad; p.dil = d->dil; p.ocs = d->out_cstride; p.oco = d->out_coff; p.act = d->act; p.slope = d->prelu_slope;
    // the halo box is TCH_WH = 16 pixels wide and the tap columns reach tx + 2 * dil <= 7 + 2 * dil: dilation <= 4 (larger
    // dilations take the per-tap kernel)
    const bool halo = d->stride == 1 && d->KH == 3 && d->KW == 3 && d->pad == d->dil && d->dil >= 1 && TCH_TW + 2 * d->dil <= TCH_WH &&
                      env_int("ARSEG_TC_HALO", 1) != 0;
    if (halo) { p.TH = TCH_TH; p.TW = TCH_TW; } else pick_tile(p.Ho, p.Wo, p.TH, p.TW);
    p.tiles_x = ceil_div(p.Wo, p.TW); p.tiles_y = ceil_div(p.Ho, p.TH);
    p.kchunks = d->Cin / bk;
    const int n256 = env_int("ARSEG_TC_N256", 1);
    const int block_n = (d->Cout >= 256 && d->Cout % 256 == 0 && n256) ? 256 : (d->Cout >= 128 ? 128 : (d->Cout > 32 ? 64 : 32));
    p.n_tiles = ceil_div(d->Cout, block_n);
    p.cout_pad = p.n_tiles * block_n;
    ARSEG_REQUIRE(p.cout_pad <= TC_MAX_COUT, "conv_tc: Cout=%d > %d", d->Cout, TC_MAX_COUT);

    const bool f16 = d->engine == ARSEG_CONV_TC_F16;
    const CUtensorMapDataType dt = tf32 ? (tf32_tma_round() ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32)
                                        : (f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
    CUtensorMap ma, mb;
    {
        cuuint64_t dims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)d->Wi, (cuuint64_t)d->Hi, (cuuint64_t)d->N};
        cuuint64_t strides[3] = {(cuuint64_t)d->Cin * es, (cuuint64_t)d->Wi * d->Cin * es, (cuuint64_t)d->Hi * d->Wi * d->Cin * es};
        // stride 2: the box spans stride * T input pixels and is traversed with element stride 2 (T pixels land in smem)
        cuuint32_t box[4] = {(cuuint32_t)bk, (cuuint32_t)(halo ? TCH_WH : p.TW * d->stride), (cuuint32_t)(halo ? TCH_TH + 2 * d->dil : p.TH * d->stride), 1};
        cuuint32_t estr[4] = {1, (cuuint32_t)d->stride, (cuuint32_t)d->stride, 1};
        CUresult r = encode(&ma, dt, 4, const_cast<void*>(d->in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("conv_tc: cuTensorMapEncodeTiled(A) failed (%d)", (int)r); return ARSEG_E_CUDA; }
    }
    {
        const cuuint64_t K = (cuuint64_t)d->KH * d->KW * d->Cin;
        cuuint64_t dims[2] = {K, (cuuint64_t)d->Cout};
        cuuint64_t strides[1] = {K * es};
        cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)block_n};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&mb, dt, 2, const_cast<void*>(d->w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("conv_tc: cuTensorMapEncodeTiled(B) failed (%d)", (int)r); return ARSEG_E_CUDA; }
    }
    if (halo) {
        TchParams hp;
        hp.c = p;
        hp.a_stage = TCH_WH * (TCH_TH + 2 * d->dil) * TC_ROW_BYTES;
        if (tf32) {
            if (block_n == 256) return launch_tch<float, 256>(ma, mb, hp, st);
            if (block_n == 128) return launch_tch<float, 128>(ma, mb, hp, st);
            if (block_n == 64) return launch_tch<float, 64>(ma, mb, hp, st);
            return launch_tch<float, 32>(ma, mb, hp, st);
        }
        if (f16) {
            if (block_n == 256) return launch_tch<__half, 256>(ma, mb, hp, st);
            if (block_n == 128) return launch_tch<__half, 128>(ma, mb, hp, st);
            if (block_n == 64) return launch_tch<__half, 64>(ma, mb, hp, st);
            return launch_tch<__half, 32>(ma, mb, hp, st);
        }
        if (block_n == 256) return launch_tch<__nv_bfloat16, 256>(ma, mb, hp, st);
        if (block_n == 128) return launch_tch<__nv_bfloat16, 128>(ma, mb, hp, st);
        if (block_n == 64) return launch_tch<__nv_bfloat16, 64>(ma, mb, hp, st);
        return launch_tch<__nv_bfloat16, 32>(ma, mb, hp, st);
    }
    if (tf32) {
        if (block_n == 256) return launch_tc<float, 256, 4>(ma, mb, p, st);
        if (block_n == 128) return launch_tc<float, 128, 6>(ma, mb, p, st);
        if (block_n == 64) return launch_tc<float, 64, 8>(ma, mb, p, st);
        return launch_tc<float, 32, 8>(ma, mb, p, st);
    }
    if (f16) {
        if (block_n == 256) return launch_tc<__half, 256, 4>(ma, mb, p, st);
        if (block_n == 128) return launch_tc<__half, 128, 6>(ma, mb, p, st);
        if (block_n == 64) return launch_tc<__half, 64, 8>(ma, mb, p, st);
        return launch_tc<__half, 32, 8>(ma, mb, p, st);
    }
    if (block_n == 256) return launch_tc<__nv_bfloat16, 256, 4>(ma, mb, p, st);
    if (block_n == 128) return launch_tc<__nv_bfloat16, 128, 6>(ma, mb, p, st);
    if (block_n == 64) return launch_tc<__nv_bfloat16, 64, 8>(ma, mb, p, st);
    return launch_tc<__nv_bfloat16, 32, 8>(ma, mb, p, st);
}

}  // namespace arseg
