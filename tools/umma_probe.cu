// umma_probe.cu -- hardware probe for the operand layouts the tcgen05 CReFF engine (csrc/creff_tc.cu) relies on.
// Single CTA; every assumption is checked against a CPU computation in the same program:
//   T1  S = Q K^T : A = Q tile [128 x 64 f16] K-major SWIZZLE_128B, B = rows of a K ring ([row][24 keys][64 ch] f16,
//       128 B per key, swizzled by address bits), several key rows per MMA (N = 24 * rows), ring wrap-around.
//   T2  O = P V   : A = P (f16 pairs written to TMEM with tcgen05.st 32x32b), B = V ring rows MN-major SWIZZLE_128B.
//   T3  logits    : A = f16 activations in TMEM, B = classifier weights [16 x 64] K-major.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/_bin/umma_probe tools/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int KP = 24;          // ring pitch (keys per ring row)
constexpr int RR = 18;          // ring rows
constexpr int NKR = 14;         // key rows of a tile (k = 7)
constexpr int NK = NKR * KP;    // 336 keys
constexpr int ROWB = KP * 128;  // bytes per ring row

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (int spin = 0; spin < (1 << 20); ++spin) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(s_u32(bar)), "r"(parity), "r"(100000u) : "memory");
        if (ok) return;
    }
    printf("probe: mbarrier timeout thread %d\n", (int)threadIdx.x);
    __trap();
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_ss(uint64_t da, uint64_t db, uint32_t tmem_d, uint32_t acc, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t ta, uint64_t db, uint32_t tmem_d, uint32_t acc, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "r"(ta), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t sbo, uint32_t lbo) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}
// idesc: c_format F32 @4, a/b format F16 = 0, b_major @16, N>>3 @17, M>>4 @24
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, int b_mn) {
    return (1u << 4) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct Args {
    const __half* q;      // [128][64]
    const __half* kring;  // [RR][KP][64]
    const __half* vring;  // [RR][KP][64]
    const __half* p;      // [128][NK]  probabilities (f16)
    const __half* act;    // [128][64]
    const __half* wcls;   // [16][64]
    float* out_s;         // [128][NK]
    float* out_o;         // [128][64]
    float* out_l;         // [128][16]
    long long* clocks;    // [8]
    int r0;               // first ring row of the key patch
    int v_lbo;            // LBO (bytes) of the MN-major V descriptor
    int p_swap;           // 1: even k in the high half of the TMEM column
    int tests;            // bit 0 T1, bit 1 T2, bit 2 T3
};

// byte offset of 16-byte chunk `c` of 128-byte row `row` (row index counted from a 1024-aligned base)
__device__ __host__ inline uint32_t swz_off(int row, int c) { return (uint32_t)(row * 128 + (((c ^ row) & 7) << 4)); }

__global__ void __launch_bounds__(160, 1) probe_kernel(Args a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sQ = smem;                                  // 16 KB
    uint8_t* sK = sQ + 128 * 128;                        // RR * ROWB
    uint8_t* sV = sK + RR * ROWB;
    uint8_t* sW = sV + RR * ROWB;                        // 16 * 128 = 2 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(sW + 2048);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // fill shared memory (generic proxy), swizzled by address bits [7:10)
    for (int i = tid; i < 128 * 8; i += blockDim.x) {
        const int row = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(sQ + swz_off(row, c)) = *reinterpret_cast<const uint4*>(a.q + row * 64 + c * 8);
    }
    for (int i = tid; i < RR * KP * 8; i += blockDim.x) {
        const int row = i >> 3, c = i & 7;               // row = ring row * KP + key column; ring rows are 3 KB = 1024-aligned
        *reinterpret_cast<uint4*>(sK + swz_off(row, c)) = *reinterpret_cast<const uint4*>(a.kring + row * 64 + c * 8);
        *reinterpret_cast<uint4*>(sV + swz_off(row, c)) = *reinterpret_cast<const uint4*>(a.vring + row * 64 + c * 8);
    }
    for (int i = tid; i < 16 * 8; i += blockDim.x) {
        const int row = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(sW + swz_off(row, c)) = *reinterpret_cast<const uint4*>(a.wcls + row * 64 + c * 8);
    }
    fence_async_smem();
    if (tid == 0) {
        for (int i = 0; i < 8; ++i) mbar_init(bars + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    constexpr uint32_t COL_S = 0, COL_O = 336, COL_A = 400, COL_L = 432;

    // ------------------------------------------------------------------ T1: S = Q K^T
    if (a.tests & 1) {
        if (warp == 4) {
            if (lane == 0) {
                const long long t0 = clock64();
                // key rows r0 .. r0+13 (mod RR) in chunks of <= 8 rows that do not cross the ring end
                int done = 0;
                while (done < NKR) {
                    const int rs = (a.r0 + done) % RR;
                    int n = NKR - done;
                    if (n > 8) n = 8;
                    if (rs + n > RR) n = RR - rs;
                    const uint32_t id = idesc_f16(128, n * KP, 0);
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t da = desc_sw128(s_u32(sQ) + k * 32, 1024, 0);
                        const uint64_t db = desc_sw128(s_u32(sK) + rs * ROWB + k * 32, 1024, 0);
                        umma_ss(da, db, tmem + COL_S + done * KP, k != 0, id);
                    }
                    done += n;
                }
                umma_commit(bars + 0);
                a.clocks[0] = clock64() - t0;
            }
            __syncwarp();
        } else {
            mbar_wait(bars + 0, 0);
            tc_fence_after();
            const long long t0 = clock64();
            const uint32_t tq = tmem + ((uint32_t)(warp * 32) << 16);
            for (int c0 = 0; c0 < NK; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(tq + COL_S + c0, r);
                tmem_ld_wait();
                for (int j = 0; j < 16; ++j) a.out_s[(warp * 32 + lane) * NK + c0 + j] = __uint_as_float(r[j]);
            }
            if (tid == 0) a.clocks[1] = clock64() - t0;
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    // ------------------------------------------------------------------ T2: O = P V (A from TMEM)
    if (a.tests & 2) {
        if (warp < 4) {
            const int q = warp * 32 + lane;
            const uint32_t tq = tmem + ((uint32_t)(warp * 32) << 16);
            for (int c0 = 0; c0 < NK / 2; c0 += 8) {
                uint32_t r[8];
                for (int j = 0; j < 8; ++j) {
                    const uint32_t lo = __half_as_ushort(a.p[q * NK + 2 * (c0 + j)]), hi = __half_as_ushort(a.p[q * NK + 2 * (c0 + j) + 1]);
                    r[j] = a.p_swap ? ((lo << 16) | hi) : ((hi << 16) | lo);
                }
                tmem_st8(tq + COL_S + c0, r);
            }
            tmem_st_wait();
            tc_fence_before();
        }
        __syncthreads();
        if (warp == 4) {
            tc_fence_after();
            if (lane == 0) {
                const long long t0 = clock64();
                const uint32_t id = idesc_f16(128, 64, 1);
                // 16 keys per MMA = 2 KB of ring; key index kk (0..335) -> ring byte offset ((r0 + kk / KP) % RR) * ROWB + (kk % KP) * 128
                for (int ks = 0; ks < NK / 16; ++ks) {
                    const int kk = ks * 16;
                    const int rrow = (a.r0 + kk / KP) % RR;
                    // NOTE: a 16-key step may straddle two ring rows (24 keys per row): rows are contiguous except at the wrap;
                    // r0 is even and RR is even, so a straddling step (kk % 48 == 16) never sits on the wrap
                    const uint32_t addr = s_u32(sV) + rrow * ROWB + (kk % KP) * 128;
                    const uint64_t db = desc_sw128(addr, 1024, (uint32_t)a.v_lbo);
                    umma_ts(tmem + COL_S + ks * 8, db, tmem + COL_O, ks != 0, id);
                }
                umma_commit(bars + 1);
                a.clocks[2] = clock64() - t0;
            }
            __syncwarp();
        } else {
            mbar_wait(bars + 1, 0);
            tc_fence_after();
            const uint32_t tq = tmem + ((uint32_t)(warp * 32) << 16);
            for (int c0 = 0; c0 < 64; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(tq + COL_O + c0, r);
                tmem_ld_wait();
                for (int j = 0; j < 32; ++j) a.out_o[(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(r[j]);
            }
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    // ------------------------------------------------------------------ T3: logits = act W^T (A from TMEM, B K-major)
    if (a.tests & 4) {
        if (warp < 4) {
            const int q = warp * 32 + lane;
            const uint32_t tq = tmem + ((uint32_t)(warp * 32) << 16);
            for (int c0 = 0; c0 < 32; c0 += 8) {
                uint32_t r[8];
                for (int j = 0; j < 8; ++j) {
                    const uint32_t lo = __half_as_ushort(a.act[q * 64 + 2 * (c0 + j)]), hi = __half_as_ushort(a.act[q * 64 + 2 * (c0 + j) + 1]);
                    r[j] = a.p_swap ? ((lo << 16) | hi) : ((hi << 16) | lo);
                }
                tmem_st8(tq + COL_A + c0, r);
            }
            tmem_st_wait();
            tc_fence_before();
        }
        __syncthreads();
        if (warp == 4) {
            tc_fence_after();
            if (lane == 0) {
                const uint32_t id = idesc_f16(128, 16, 0);
                for (int k = 0; k < 4; ++k) {
                    const uint64_t db = desc_sw128(s_u32(sW) + k * 32, 1024, 0);
                    umma_ts(tmem + COL_A + k * 8, db, tmem + COL_L, k != 0, id);
                }
                umma_commit(bars + 2);
            }
            __syncwarp();
        } else {
            mbar_wait(bars + 2, 0);
            tc_fence_after();
            const uint32_t tq = tmem + ((uint32_t)(warp * 32) << 16);
            uint32_t r[16];
            tmem_ld16(tq + COL_L, r);
            tmem_ld_wait();
            for (int j = 0; j < 16; ++j) a.out_l[(warp * 32 + lane) * 16 + j] = __uint_as_float(r[j]);
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// ------------------------------------------------------------------------------------------------------------------
// timing kernel: issue / completion cost of the MMA shapes and TMEM read patterns the engine uses (one CTA, clock64)
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
}
__global__ void __launch_bounds__(160, 1) time_kernel(long long* out, float* sink) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + 128 * 128;
    uint8_t* sV = sK + RR * ROWB;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + RR * ROWB);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (128 * 128 + 2 * RR * ROWB) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x2c002c00u;   // small f16 values
    fence_async_smem();
    if (tid == 0) {
        for (int i = 0; i < 16; ++i) mbar_init(bars + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (warp == 4) {
        // whole warp runs the loops (warp-uniform operands -> uniform registers), one elected lane issues
        const uint64_t dq = desc_sw128(s_u32(sQ), 1024, 0), dk = desc_sw128(s_u32(sK), 1024, 0), dv = desc_sw128(s_u32(sV), 1024, 0);
        for (int test = 0; test < 6; ++test) {
            const long long t0 = clock64();
            if (elect_one()) {
                if (test == 0) {            // S = Q K^T: N = 192 + 144, 4 k-steps each
                    for (int k = 0; k < 4; ++k) umma_ss(dq + 2 * k, dk + 2 * k, tmem, k != 0, idesc_f16(128, 192, 0));
                    for (int k = 0; k < 4; ++k) umma_ss(dq + 2 * k, dk + 8 * (ROWB >> 4) + 2 * k, tmem + 192, k != 0, idesc_f16(128, 144, 0));
                } else if (test == 1) {     // O = P V: 21 x (N = 64, A from TMEM, B MN-major)
                    for (int ks = 0; ks < 21; ++ks) umma_ts(tmem + 8 * ks, dv + ks * 128, tmem + 336, ks != 0, idesc_f16(128, 64, 1));
                } else if (test == 2) {     // O = P V + row sums (N = 16 against a constant tile)
                    for (int ks = 0; ks < 21; ++ks) {
                        umma_ts(tmem + 8 * ks, dv + ks * 128, tmem + 336, ks != 0, idesc_f16(128, 64, 1));
                        umma_ts(tmem + 8 * ks, dk, tmem + 400, ks != 0, idesc_f16(128, 16, 0));
                    }
                } else if (test == 3) {     // O = P V with N = 80: V extended by 16 channels would give the sums in the same MMA
                    for (int ks = 0; ks < 21; ++ks) umma_ts(tmem + 8 * ks, dv + ks * 128, tmem + 336, ks != 0, idesc_f16(128, 64, 1));
                    for (int ks = 0; ks < 21; ++ks) umma_ts(tmem + 8 * ks, dk, tmem + 400, ks != 0, idesc_f16(128, 16, 0));
                } else if (test == 4) {     // P as an smem A operand (K-major), V MN-major
                    for (int ks = 0; ks < 21; ++ks) umma_ss(dq + 2 * (ks & 3), dv + ks * 128, tmem + 336, ks != 0, idesc_f16(128, 64, 1));
                } else {                    // classifier: 4 x (N = 16, A from TMEM)
                    for (int k = 0; k < 4; ++k) umma_ts(tmem + 416 + 8 * k, dk + 2 * k, tmem + 448, k != 0, idesc_f16(128, 16, 0));
                }
                umma_commit(bars + test);
            }
            __syncwarp();
            const long long t1 = clock64();
            mbar_wait(bars + test, 0);
            const long long t2 = clock64();
            if (lane == 0) { out[2 * test] = t1 - t0; out[2 * test + 1] = t2 - t0; }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp < 4) {
        const uint32_t tq = tmem + ((uint32_t)(warp * 32) << 16);
        float acc = 0.f;
        // (a) row by row: ld16 + ld8, wait, consume -- 8 rows, as pass 1 of the softmax
        long long t0 = clock64();
        for (int j = 0; j < 8; ++j) {
            uint32_t r[24];
            tmem_ld16(tq + j * 24, reinterpret_cast<uint32_t(&)[16]>(r[0]));
            tmem_ld8(tq + j * 24 + 16, r + 16);
            tmem_ld_wait();
            for (int c = 0; c < 22; ++c) acc = fmaxf(acc, __uint_as_float(r[c]));
        }
        long long t1 = clock64();
        if (tid == 0) out[16] = t1 - t0;
        // (b) the same with the next row's loads in flight while the current row is consumed
        t0 = clock64();
        {
            uint32_t ra[24], rb[24];
            tmem_ld16(tq, reinterpret_cast<uint32_t(&)[16]>(ra[0])); tmem_ld8(tq + 16, ra + 16);
            for (int j = 0; j < 8; j += 2) {
                tmem_ld_wait();
                tmem_ld16(tq + (j + 1) * 24, reinterpret_cast<uint32_t(&)[16]>(rb[0])); tmem_ld8(tq + (j + 1) * 24 + 16, rb + 16);
                for (int c = 0; c < 22; ++c) acc = fmaxf(acc, __uint_as_float(ra[c]));
                tmem_ld_wait();
                if (j + 2 < 8) { tmem_ld16(tq + (j + 2) * 24, reinterpret_cast<uint32_t(&)[16]>(ra[0])); tmem_ld8(tq + (j + 2) * 24 + 16, ra + 16); }
                for (int c = 0; c < 22; ++c) acc = fmaxf(acc, __uint_as_float(rb[c]));
            }
        }
        t1 = clock64();
        if (tid == 0) out[17] = t1 - t0;
        // (c) one 32-column load + wait, repeated: raw round trip
        t0 = clock64();
        for (int j = 0; j < 8; ++j) {
            uint32_t r[32];
            tmem_ld32(tq + j * 32, r);
            tmem_ld_wait();
            acc += __uint_as_float(r[j]);
        }
        t1 = clock64();
        if (tid == 0) out[18] = t1 - t0;
        // (d) 176 ex2 + 88 packs per lane (pass 2 arithmetic without TMEM traffic)
        t0 = clock64();
        float x = acc * 1e-30f;
        uint32_t pk = 0;
        for (int j = 0; j < 88; ++j) {
            float e0, e1;
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(x + (float)j));
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(x - (float)j));
            __half2 h = __floats2half2_rn(e0, e1);
            pk ^= *reinterpret_cast<uint32_t*>(&h);
        }
        t1 = clock64();
        if (tid == 0) out[19] = t1 - t0;
        // (e) tcgen05.st of 12 columns x 8 rows + wait
        t0 = clock64();
        for (int j = 0; j < 8; ++j) {
            uint32_t r[8] = {pk, pk, pk, pk, pk, pk, pk, pk};
            tmem_st8(tq + j * 12, r);
        }
        tmem_st_wait();
        t1 = clock64();
        if (tid == 0) out[20] = t1 - t0;
        if (acc == 123.f) sink[tid] = acc + (float)pk;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

static float frand(uint32_t& s) { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffff) / 65536.f - 0.5f; }

int main() {
    uint32_t seed = 12345;
    std::vector<__half> q(128 * 64), kr(RR * KP * 64), vr(RR * KP * 64), p(128 * NK), act(128 * 64), w(16 * 64);
    for (auto& x : q) x = __float2half(frand(seed));
    for (auto& x : kr) x = __float2half(frand(seed));
    for (auto& x : vr) x = __float2half(frand(seed));
    for (auto& x : p) x = __float2half(frand(seed) + 0.5f);
    for (auto& x : act) x = __float2half(frand(seed) * 4.f);
    for (auto& x : w) x = __float2half(frand(seed));
    __half *dq, *dk, *dv, *dp, *da, *dw;
    float *ds, *dO, *dl;
    long long* dc;
    CK(cudaMalloc(&dq, q.size() * 2)); CK(cudaMalloc(&dk, kr.size() * 2)); CK(cudaMalloc(&dv, vr.size() * 2));
    CK(cudaMalloc(&dp, p.size() * 2)); CK(cudaMalloc(&da, act.size() * 2)); CK(cudaMalloc(&dw, w.size() * 2));
    CK(cudaMalloc(&ds, 128 * NK * 4)); CK(cudaMalloc(&dO, 128 * 64 * 4)); CK(cudaMalloc(&dl, 128 * 16 * 4)); CK(cudaMalloc(&dc, 64));
    CK(cudaMemcpy(dq, q.data(), q.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dk, kr.data(), kr.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dv, vr.data(), vr.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dp, p.data(), p.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(da, act.data(), act.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dw, w.data(), w.size() * 2, cudaMemcpyHostToDevice));
    const size_t smem = 1024 + 128 * 128 + 2 * RR * ROWB + 2048 + 256;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

    struct Variant { int r0, v_lbo, p_swap; };
    const Variant vars[] = {{0, 0, 0}, {8, 0, 0}, {16, 0, 0}, {0, 1024, 0}, {0, 0, 1}, {8, 128, 0}};
    for (const Variant& v : vars) {
        CK(cudaMemset(ds, 0xff, 128 * NK * 4)); CK(cudaMemset(dO, 0xff, 128 * 64 * 4)); CK(cudaMemset(dl, 0xff, 128 * 16 * 4));
        Args a{dq, dk, dv, dp, da, dw, ds, dO, dl, dc, v.r0, v.v_lbo, v.p_swap, 7};
        probe_kernel<<<1, 160, smem>>>(a);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant r0=%d lbo=%d swap=%d: kernel failed: %s\n", v.r0, v.v_lbo, v.p_swap, cudaGetErrorString(e)); return 1; }
        std::vector<float> hs(128 * NK), ho(128 * 64), hl(128 * 16);
        long long hc[8];
        CK(cudaMemcpy(hs.data(), ds, hs.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(ho.data(), dO, ho.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hl.data(), dl, hl.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hc, dc, 64, cudaMemcpyDeviceToHost));
        double es = 0, eo = 0, el = 0;
        for (int m = 0; m < 128; ++m) {
            for (int kk = 0; kk < NK; ++kk) {
                const int rrow = (v.r0 + kk / KP) % RR, col = kk % KP;
                double acc = 0;
                for (int c = 0; c < 64; ++c) acc += (double)__half2float(q[m * 64 + c]) * (double)__half2float(kr[(rrow * KP + col) * 64 + c]);
                es = fmax(es, fabs(acc - hs[m * NK + kk]));
            }
            for (int c = 0; c < 64; ++c) {
                double acc = 0;
                for (int kk = 0; kk < NK; ++kk) {
                    const int rrow = (v.r0 + kk / KP) % RR, col = kk % KP;
                    acc += (double)__half2float(p[m * NK + kk]) * (double)__half2float(vr[(rrow * KP + col) * 64 + c]);
                }
                eo = fmax(eo, fabs(acc - ho[m * 64 + c]));
            }
            for (int j = 0; j < 16; ++j) {
                double acc = 0;
                for (int c = 0; c < 64; ++c) acc += (double)__half2float(act[m * 64 + c]) * (double)__half2float(w[j * 64 + c]);
                el = fmax(el, fabs(acc - hl[m * 16 + j]));
            }
        }
        printf("variant r0=%2d v_lbo=%4d p_swap=%d : max|err| S=%.3e  O=%.3e  L=%.3e   clocks: qk_issue=%lld s_readback=%lld pv_issue=%lld\n",
               v.r0, v.v_lbo, v.p_swap, es, eo, el, hc[0], hc[1], hc[2]);
    }
    {
        long long* dt; float* sink;
        CK(cudaMalloc(&dt, 64 * 8)); CK(cudaMalloc(&sink, 1024)); CK(cudaMemset(dt, 0, 64 * 8));
        const size_t tsm = 1024 + 128 * 128 + 2 * RR * ROWB + 512;
        CK(cudaFuncSetAttribute(time_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm));
        for (int rep = 0; rep < 2; ++rep) {
            time_kernel<<<1, 160, tsm>>>(dt, sink);
            CK(cudaDeviceSynchronize());
        }
        long long h[64];
        CK(cudaMemcpy(h, dt, 64 * 8, cudaMemcpyDeviceToHost));
        const char* names[6] = {"QK 8 MMAs (N=192,144)", "PV 21 x N=64 TS", "PV+SUM interleaved 42", "PV then SUM 42", "PV 21 x N=64 SS", "CLS 4 x N=16 TS"};
        for (int i = 0; i < 6; ++i) printf("mma timing %-26s issue %5lld  complete %5lld cycles\n", names[i], h[2 * i], h[2 * i + 1]);
        printf("tmem: 8 rows ld24+wait+max %lld | double-buffered %lld | 8 x (ld32+wait) %lld | 176 ex2 + 88 pack %lld | 8 x st8 + wait %lld cycles\n",
               h[16], h[17], h[18], h[19], h[20]);
    }
    return 0;
}
