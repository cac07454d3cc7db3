// ops_basic.cu -- layout conversion, resampling, pooling, gating, stem conv (HBM-bound glue ops).
// All kernels are written for sm_100a; NHWC activations, 16-byte vectorised along C where aligned.
#include "common.cuh"
#include <cstdlib>
#include <mutex>

namespace arseg {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// ------------------------------------------------------------------------------------------
// NCHW <-> NHWC (per image: transpose of a C x (H*W) matrix) through a 32x33 smem tile
// ------------------------------------------------------------------------------------------
// ROWS_ON_X: the row-tile index is blockIdx.x (the long axis of either direction: H*W tiles)
template <typename TI, typename TO, bool ROWS_ON_X = false>
__global__ void transpose_kernel(const TI* __restrict__ src, TO* __restrict__ dst, int rows, int cols) {
    // src: [n][rows][cols] -> dst: [n][cols][rows]
    __shared__ float tile[32][33];
    const size_t base = (size_t)blockIdx.z * rows * cols;
    const int c0 = (ROWS_ON_X ? blockIdx.y : blockIdx.x) * 32, r0 = (ROWS_ON_X ? blockIdx.x : blockIdx.y) * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int r = r0 + i, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[i][threadIdx.x] = to_f32(src[base + (size_t)r * cols + c]);
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int c = c0 + i, r = r0 + threadIdx.x;
        if (r < rows && c < cols) dst[base + (size_t)c * rows + r] = from_f32<TO>(tile[threadIdx.x][i]);
    }
}

// ------------------------------------------------------------------------------------------
// bilinear / nearest resize of fp32 planes (NCHW)
// ------------------------------------------------------------------------------------------
__global__ void resize_nchw_kernel(const float* __restrict__ src, float* __restrict__ dst, int planes,
                                   int Hi, int Wi, int Ho, int Wo, int mode, float sh, float sw) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
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
    for (int p = blockIdx.z; p < planes; p += gridDim.z) {
        const float* s = src + (size_t)p * Hi * Wi;
        float v;
        if (mode == ARSEG_RESIZE_NEAREST) v = s[(size_t)y0 * Wi + x0];
        else v = ly0 * (lx0 * s[(size_t)y0 * Wi + x0] + lx1 * s[(size_t)y0 * Wi + x1]) +
                 ly1 * (lx0 * s[(size_t)y1 * Wi + x0] + lx1 * s[(size_t)y1 * Wi + x1]);
        dst[((size_t)p * Ho + y) * Wo + x] = v;
    }
}

// ------------------------------------------------------------------------------------------
// resize NHWC -> channel slice of an NHWC destination.  One thread per (pixel, VEC channels).
// ------------------------------------------------------------------------------------------
template <typename T, int VEC>
struct alignas(sizeof(T) * VEC) Pack { T v[VEC]; };

// One CTA per output row (grid-stride over N*Ho rows): the row's vertical taps / weights are computed once and the
// per-element index math is 32-bit (64-bit div/mod per element made the previous flat version ALU-bound).
template <typename T, int VEC>
__global__ void __launch_bounds__(256) resize_nhwc_kernel(const T* __restrict__ src, T* __restrict__ dst, int N, int Hi, int Wi, int C,
                                                          int Ho, int Wo, int dcs, int dco, int mode, float sh, float sw) {
    const int cv = C / VEC;
    const int row_elems = Wo * cv;
    for (int row = blockIdx.x; row < N * Ho; row += gridDim.x) {
        const int n = row / Ho, y = row - n * Ho;
        const T* s = src + (size_t)n * Hi * Wi * C;
        T* drow = dst + ((size_t)row * Wo) * dcs + dco;
        if (mode == ARSEG_RESIZE_NEAREST) {
            const T* srow = s + (size_t)nearest_src(sh, y, Hi) * Wi * C;
            for (int i = threadIdx.x; i < row_elems; i += blockDim.x) {
                const int x = i / cv, c = (i - x * cv) * VEC;
                *reinterpret_cast<Pack<T, VEC>*>(drow + (size_t)x * dcs + c) =
                    *reinterpret_cast<const Pack<T, VEC>*>(srow + (size_t)nearest_src(sw, x, Wi) * C + c);
            }
        } else {
            int y0, y1;
            float ly0, ly1;
            bilinear_src(sh, y, Hi, mode, y0, y1, ly0, ly1);
            const T* r0 = s + (size_t)y0 * Wi * C;
            const T* r1 = s + (size_t)y1 * Wi * C;
            for (int i = threadIdx.x; i < row_elems; i += blockDim.x) {
                const int x = i / cv, c = (i - x * cv) * VEC;
                int x0, x1;
                float lx0, lx1;
                bilinear_src(sw, x, Wi, mode, x0, x1, lx0, lx1);
                const Pack<T, VEC> a = *reinterpret_cast<const Pack<T, VEC>*>(r0 + (size_t)x0 * C + c);
                const Pack<T, VEC> b = *reinterpret_cast<const Pack<T, VEC>*>(r0 + (size_t)x1 * C + c);
                const Pack<T, VEC> cc = *reinterpret_cast<const Pack<T, VEC>*>(r1 + (size_t)x0 * C + c);
                const Pack<T, VEC> d = *reinterpret_cast<const Pack<T, VEC>*>(r1 + (size_t)x1 * C + c);
                Pack<T, VEC> o;
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    float v = ly0 * (lx0 * to_f32(a.v[k]) + lx1 * to_f32(b.v[k])) +
                              ly1 * (lx0 * to_f32(cc.v[k]) + lx1 * to_f32(d.v[k]));
                    o.v[k] = from_f32<T>(v);
                }
                *reinterpret_cast<Pack<T, VEC>*>(drow + (size_t)x * dcs + c) = o;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Exact x2 bilinear up-sampling, align_corners=False (F.upsample(scale_factor=2), model/pspnet.py:38-41): one thread
// per (input pixel, VEC channels) loads the clamped 3x3 neighbourhood once and writes the 2x2 output block, 2.25 loads
// per output instead of 4.  Same arithmetic as resize_nhwc_kernel, expression for expression: for an exact factor of 2
// the ATen source coordinates are i - 0.25 / i + 0.25, i.e. weights (0.25, 0.75) / (0.75, 0.25), and (1, 0) in row /
// column 0 where the coordinate clamps to 0.
// ------------------------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void __launch_bounds__(256) upsample2x_nhwc_kernel(const T* __restrict__ src, T* __restrict__ dst, int N, int Hi, int Wi, int C,
                                                              int dcs, int dco) {
    const int cv = C / VEC;
    const long long total = (long long)N * Hi * Wi * cv;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % cv) * VEC;
        long long r = idx / cv;
        const int j = (int)(r % Wi); r /= Wi;
        const int i = (int)(r % Hi);
        const int n = (int)(r / Hi);
        const int ys[3] = {i > 0 ? i - 1 : 0, i, i < Hi - 1 ? i + 1 : i};
        const int xs[3] = {j > 0 ? j - 1 : 0, j, j < Wi - 1 ? j + 1 : j};
        const T* s = src + (size_t)n * Hi * Wi * C + c;
        float v[3][3][VEC];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                const Pack<T, VEC> pk = *reinterpret_cast<const Pack<T, VEC>*>(s + ((size_t)ys[a] * Wi + xs[b]) * C);
#pragma unroll
                for (int k = 0; k < VEC; ++k) v[a][b][k] = to_f32(pk.v[k]);
            }
        // output row 2i + dy reads input rows (dy, dy + 1) of the neighbourhood with weights wy[dy]; columns likewise
        const float wy0[2] = {i == 0 ? 1.f : 0.25f, 0.75f}, wy1[2] = {i == 0 ? 0.f : 0.75f, 0.25f};
        const float wx0[2] = {j == 0 ? 1.f : 0.25f, 0.75f}, wx1[2] = {j == 0 ? 0.f : 0.75f, 0.25f};
        T* d = dst + (((size_t)n * 2 * Hi + 2 * i) * (2 * Wi) + 2 * j) * dcs + dco + c;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                Pack<T, VEC> o;
#pragma unroll
                for (int k = 0; k < VEC; ++k)
                    o.v[k] = from_f32<T>(wy0[dy] * (wx0[dx] * v[dy][dx][k] + wx1[dx] * v[dy][dx + 1][k]) +
                                         wy1[dy] * (wx0[dx] * v[dy + 1][dx][k] + wx1[dx] * v[dy + 1][dx + 1][k]));
                *reinterpret_cast<Pack<T, VEC>*>(d + ((size_t)dy * 2 * Wi + dx) * dcs) = o;
            }
    }
}

template <typename T, int VEC>
static void upsample2x_launch(const T* s, T* d, int N, int Hi, int Wi, int C, int dcs, int dco, cudaStream_t st) {
    const long long total = (long long)N * Hi * Wi * (C / VEC);
    const long long cap = (long long)sm_count() * 16, blocks = (total + 255) / 256;
    upsample2x_nhwc_kernel<T, VEC><<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(s, d, N, Hi, Wi, C, dcs, dco);
}

// ------------------------------------------------------------------------------------------
// adaptive average pool NHWC: block = (32 channel lanes, 8 window rows); grid = (N*Ho*Wo, ceil(C/32))
// ------------------------------------------------------------------------------------------
template <typename T, bool MAXP>
__global__ void adaptive_pool_kernel(const T* __restrict__ in, void* __restrict__ out_, int N, int H, int W, int C,
                                     int Ho, int Wo, int out_f32) {
    __shared__ float red[8][33];
    int p = blockIdx.x;
    const int ox = p % Wo; p /= Wo;
    const int oy = p % Ho;
    const int n = p / Ho;
    const int c = blockIdx.y * 32 + threadIdx.x;
    const int ys = (oy * H) / Ho, ye = ((oy + 1) * H + Ho - 1) / Ho;
    const int xs = (ox * W) / Wo, xe = ((ox + 1) * W + Wo - 1) / Wo;
    const int ww = xe - xs, cnt = (ye - ys) * ww;
    float acc = MAXP ? -INFINITY : 0.f;
    if (c < C) {
        for (int i = threadIdx.y; i < cnt; i += 8) {
            const int y = ys + i / ww, x = xs + i % ww;
            const float v = to_f32(in[(((size_t)n * H + y) * W + x) * C + c]);
            acc = MAXP ? fmaxf(acc, v) : acc + v;
        }
    }
    red[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
#pragma unroll
        for (int i = 1; i < 8; ++i) acc = MAXP ? fmaxf(acc, red[i][threadIdx.x]) : acc + red[i][threadIdx.x];
        const size_t o = (((size_t)n * Ho + oy) * Wo + ox) * C + c;
        if (MAXP) reinterpret_cast<float*>(out_)[o] = acc;
        else if (out_f32) reinterpret_cast<float*>(out_)[o] = acc / (float)cnt;
        else reinterpret_cast<T*>(out_)[o] = from_f32<T>(acc / (float)cnt);
    }
}

// ------------------------------------------------------------------------------------------
// tiny fp32 linear: one warp per output element
// ------------------------------------------------------------------------------------------
__global__ void linear_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                              float* __restrict__ y, int N, int K, int M, int relu) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= N * M) return;
    const int n = warp / M, m = warp % M;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(x[(size_t)n * K + k], w[(size_t)m * K + k], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        acc += b ? b[m] : 0.f;
        y[(size_t)n * M + m] = relu ? fmaxf(acc, 0.f) : acc;
    }
}

// ------------------------------------------------------------------------------------------
// ARM / FFM gating
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void gate_kernel(const T* __restrict__ feat, const float* __restrict__ gate, const float* __restrict__ gs,
                            const float* __restrict__ gb, int add_identity, const float* __restrict__ add_chan,
                            const T* __restrict__ add_pix, T* __restrict__ out, int N, long long HW, int C) {
    const long long total = (long long)N * HW * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int n = (int)(i / (HW * C));
        const float f = to_f32(feat[i]);
        float g = gate[(size_t)n * C + c];
        g = g * (gs ? gs[c] : 1.f) + (gb ? gb[c] : 0.f);
        g = 1.f / (1.f + expf(-g));
        float v = f * g;
        if (add_identity) v += f;
        if (add_chan) v += add_chan[(size_t)n * C + c];
        if (add_pix) v += to_f32(add_pix[i]);
        out[i] = from_f32<T>(v);
    }
}

// ------------------------------------------------------------------------------------------
// maxpool 3x3 s2 p1 NHWC
// ------------------------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void maxpool_kernel(const T* __restrict__ in, T* __restrict__ out, int N, int H, int W, int C, int Ho, int Wo) {
    const int cv = C / VEC;
    const long long total = (long long)N * Ho * Wo * cv;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % cv) * VEC;
        long long p = idx / cv;
        const int x = (int)(p % Wo); p /= Wo;
        const int y = (int)(p % Ho);
        const int n = (int)(p / Ho);
        float m[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) m[i] = -INFINITY;
        for (int dy = 0; dy < 3; ++dy) {
            const int iy = y * 2 - 1 + dy;
            if (iy < 0 || iy >= H) continue;
            for (int dx = 0; dx < 3; ++dx) {
                const int ix = x * 2 - 1 + dx;
                if (ix < 0 || ix >= W) continue;
                const Pack<T, VEC> v = *reinterpret_cast<const Pack<T, VEC>*>(in + (((size_t)n * H + iy) * W + ix) * C + c);
#pragma unroll
                for (int i = 0; i < VEC; ++i) m[i] = fmaxf(m[i], to_f32(v.v[i]));
            }
        }
        Pack<T, VEC> o;
#pragma unroll
        for (int i = 0; i < VEC; ++i) o.v[i] = from_f32<T>(m[i]);
        *reinterpret_cast<Pack<T, VEC>*>(out + (((size_t)n * Ho + y) * Wo + x) * C + c) = o;
    }
}

// ------------------------------------------------------------------------------------------
// stem: conv 7x7 s2 p3 (Cin=3) + scale/shift + ReLU.  NCHW fp32 in -> NHWC out.
// Block = 16x16 output pixels; weights for a 32-channel half live in smem as [tap][32] and are read
// as broadcast float4; each thread keeps 32 accumulators for its pixel.
// ------------------------------------------------------------------------------------------
// Register-tiled direct convolution: a CTA computes a 4-row x 64-column output tile for all 64 channels; a thread owns
// 4 output pixels of one row (columns tx, tx+16, tx+32, tx+48) x 16 channels, so every weight vector read from shared
// memory (LDS.128, warp-broadcast) feeds 16 FMAs and every input value 16.  Input tile 13 x 133 x 3 fp32 and the
// [147][64] weights live in shared memory (57 KB).
constexpr int STEM_TW = 64, STEM_TH = 4, STEM_CG = 16;
constexpr int STEM_IW = STEM_TW * 2 + 5, STEM_IH = STEM_TH * 2 + 5;     // 133 x 13
constexpr int STEM_IWP = STEM_IW + 1;
constexpr int STEM_IN_FLOATS = (3 * STEM_IH * STEM_IWP + 3) / 4 * 4;                // keeps s_w 16-byte aligned
constexpr size_t STEM_SMEM = (size_t)(STEM_IN_FLOATS + 147 * 64) * 4;
template <typename T>
__global__ void __launch_bounds__(256) stem_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                   const float* __restrict__ scale, const float* __restrict__ shift,
                                                   T* __restrict__ out, int H, int W, int Ho, int Wo, int Cout) {
    extern __shared__ __align__(16) float stem_sm[];
    float* s_in = stem_sm;                                   // [3][STEM_IH][STEM_IWP]
    float* s_w = stem_sm + STEM_IN_FLOATS;                   // [147][64]
    const int n = blockIdx.z;
    const int ox0 = blockIdx.x * STEM_TW, oy0 = blockIdx.y * STEM_TH;
    const int tid = threadIdx.x;
    const int ix0 = ox0 * 2 - 3, iy0 = oy0 * 2 - 3;
    for (int i = tid; i < 3 * STEM_IH * STEM_IW; i += 256) {
        const int c = i / (STEM_IH * STEM_IW), r = (i / STEM_IW) % STEM_IH, q = i % STEM_IW;
        const int y = iy0 + r, x = ix0 + q;
        float v = 0.f;
        if (y >= 0 && y < H && x >= 0 && x < W) v = __ldg(in + (((size_t)n * 3 + c) * H + y) * W + x);
        s_in[(c * STEM_IH + r) * STEM_IWP + q] = v;
    }
    // w: [Cout][7][7][3] -> s_w[tap*3+c][co]  (Cout <= 64; missing channels zero)
    for (int i = tid; i < 147 * 64; i += 256) {
        const int co = i / 147, t = i % 147;
        s_w[t * 64 + co] = co < Cout ? __ldg(w + (size_t)co * 147 + t) : 0.f;
    }
    __syncthreads();
    const int tx = tid & 15, ty = (tid >> 4) & 3, cg = tid >> 6;       // a warp = 2 rows x 16 columns of one channel group
    float acc[4][STEM_CG];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int i = 0; i < STEM_CG; ++i) acc[p][i] = 0.f;
#pragma unroll 1
    for (int ky = 0; ky < 7; ++ky) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float* row = s_in + (c * STEM_IH + ty * 2 + ky) * STEM_IWP + tx * 2;
#pragma unroll
            for (int kx = 0; kx < 7; ++kx) {
                const float4* wp = reinterpret_cast<const float4*>(s_w + ((ky * 7 + kx) * 3 + c) * 64 + cg * STEM_CG);
                const float4 w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3];
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const float v = row[p * 32 + kx];
                    acc[p][0] = fmaf(v, w0.x, acc[p][0]); acc[p][1] = fmaf(v, w0.y, acc[p][1]);
                    acc[p][2] = fmaf(v, w0.z, acc[p][2]); acc[p][3] = fmaf(v, w0.w, acc[p][3]);
                    acc[p][4] = fmaf(v, w1.x, acc[p][4]); acc[p][5] = fmaf(v, w1.y, acc[p][5]);
                    acc[p][6] = fmaf(v, w1.z, acc[p][6]); acc[p][7] = fmaf(v, w1.w, acc[p][7]);
                    acc[p][8] = fmaf(v, w2.x, acc[p][8]); acc[p][9] = fmaf(v, w2.y, acc[p][9]);
                    acc[p][10] = fmaf(v, w2.z, acc[p][10]); acc[p][11] = fmaf(v, w2.w, acc[p][11]);
                    acc[p][12] = fmaf(v, w3.x, acc[p][12]); acc[p][13] = fmaf(v, w3.y, acc[p][13]);
                    acc[p][14] = fmaf(v, w3.z, acc[p][14]); acc[p][15] = fmaf(v, w3.w, acc[p][15]);
                }
            }
        }
    }
    const int oy = oy0 + ty, co0 = cg * STEM_CG;
    if (oy < Ho && co0 < Cout) {
        float sc[STEM_CG], sh[STEM_CG];
#pragma unroll
        for (int i = 0; i < STEM_CG; ++i) {
            sc[i] = co0 + i < Cout ? __ldg(scale + co0 + i) : 0.f;
            sh[i] = co0 + i < Cout ? __ldg(shift + co0 + i) : 0.f;
        }
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int ox = ox0 + tx + 16 * p;
            if (ox >= Wo) continue;
            T* o = out + (((size_t)n * Ho + oy) * Wo + ox) * Cout + co0;
            if (co0 + STEM_CG <= Cout && Cout % 8 == 0) {
                alignas(16) T v[STEM_CG];
#pragma unroll
                for (int i = 0; i < STEM_CG; ++i) v[i] = from_f32<T>(fmaxf(fmaf(acc[p][i], sc[i], sh[i]), 0.f));
                constexpr int NV = (int)(STEM_CG * sizeof(T) / 16);
#pragma unroll
                for (int q = 0; q < NV; ++q) reinterpret_cast<uint4*>(o)[q] = reinterpret_cast<const uint4*>(v)[q];
            } else {
#pragma unroll
                for (int i = 0; i < STEM_CG; ++i)
                    if (co0 + i < Cout) o[i] = from_f32<T>(fmaxf(fmaf(acc[p][i], sc[i], sh[i]), 0.f));
            }
        }
    }
}

static inline int grid_1d(long long total, int block) {
    long long g = ceil_div_ll(total, block);
    const long long cap = (long long)sm_count() * 16;
    return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace arseg

using namespace arseg;

namespace arseg {
int stem_mma_launch(const float* in, const float* w, const float* scale, const float* shift, void* out, int out_dtype, int N, int H,
                    int W, int Ho, int Wo, cudaStream_t st);   // stem_mma.cu
}

extern "C" {

int arseg_abi_version(void) { return ARSEG_ABI_VERSION; }
const char* arseg_last_error(void) { return g_err; }

int arseg_nchw_to_nhwc(const float* src, void* dst, int dst_dtype, int N, int C, int H, int W, arseg_stream_t stream) {
    ARSEG_REQUIRE(src && dst && N > 0 && C > 0 && H > 0 && W > 0, "nchw_to_nhwc: bad args");
    dim3 grid(ceil_div(H * W, 32), ceil_div(C, 32), N), block(32, 8);
    ARSEG_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "nchw_to_nhwc: dims too large");
    if (dst_dtype == ARSEG_F32)
        transpose_kernel<float, float><<<grid, block, 0, as_stream(stream)>>>(src, (float*)dst, C, H * W);
    else if (dst_dtype == ARSEG_BF16)
        transpose_kernel<float, __nv_bfloat16><<<grid, block, 0, as_stream(stream)>>>(src, (__nv_bfloat16*)dst, C, H * W);
    else if (dst_dtype == ARSEG_F16)
        transpose_kernel<float, __half><<<grid, block, 0, as_stream(stream)>>>(src, (__half*)dst, C, H * W);
    else ARSEG_UNSUPPORTED("nchw_to_nhwc: dtype %d", dst_dtype);
    ARSEG_CHECK_LAUNCH("nchw_to_nhwc");
    return ARSEG_OK;
}

int arseg_nhwc_to_nchw(const void* src, int src_dtype, float* dst, int N, int C, int H, int W, arseg_stream_t stream) {
    ARSEG_REQUIRE(src && dst && N > 0 && C > 0 && H > 0 && W > 0, "nhwc_to_nchw: bad args");
    // the pixel-tile index goes on grid.x (2^31 - 1 blocks): 1024 x 2048 is 65536 tiles, one more than grid.y allows
    dim3 grid(ceil_div(H * W, 32), ceil_div(C, 32), N), block(32, 8);
    ARSEG_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "nhwc_to_nchw: dims too large");
    if (src_dtype == ARSEG_F32)
        transpose_kernel<float, float, true><<<grid, block, 0, as_stream(stream)>>>((const float*)src, dst, H * W, C);
    else if (src_dtype == ARSEG_BF16)
        transpose_kernel<__nv_bfloat16, float, true><<<grid, block, 0, as_stream(stream)>>>((const __nv_bfloat16*)src, dst, H * W, C);
    else if (src_dtype == ARSEG_F16)
        transpose_kernel<__half, float, true><<<grid, block, 0, as_stream(stream)>>>((const __half*)src, dst, H * W, C);
    else ARSEG_UNSUPPORTED("nhwc_to_nchw: dtype %d", src_dtype);
    ARSEG_CHECK_LAUNCH("nhwc_to_nchw");
    return ARSEG_OK;
}

int arseg_resize_nchw_f32(const float* src, float* dst, int planes, int Hi, int Wi, int Ho, int Wo, int mode,
                          arseg_stream_t stream) {
    ARSEG_REQUIRE(src && dst && planes > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, "resize_nchw: bad args");
    ARSEG_REQUIRE(mode >= 0 && mode <= 2 && Ho <= 65535, "resize_nchw: bad mode/size");
    const float sh = resize_scale(Hi, Ho, mode), sw = resize_scale(Wi, Wo, mode);
    dim3 block(128), grid(ceil_div(Wo, 128), Ho, planes < 64 ? planes : 64);
    resize_nchw_kernel<<<grid, block, 0, as_stream(stream)>>>(src, dst, planes, Hi, Wi, Ho, Wo, mode, sh, sw);
    ARSEG_CHECK_LAUNCH("resize_nchw");
    return ARSEG_OK;
}

static inline int resize_grid(int N, int Ho) { const long long r = (long long)N * Ho, cap = (long long)sm_count() * 32; return (int)(r < cap ? r : cap); }

int arseg_resize_nhwc(const void* src, void* dst, int dtype, int N, int Hi, int Wi, int C, int Ho, int Wo,
                      int dcs, int dco, int mode, arseg_stream_t stream) {
    ARSEG_REQUIRE(src && dst && N > 0 && Hi > 0 && Wi > 0 && C > 0 && Ho > 0 && Wo > 0, "resize_nhwc: bad args");
    ARSEG_REQUIRE(mode >= 0 && mode <= 2 && dcs >= C && dco >= 0 && dco + C <= dcs, "resize_nhwc: bad mode/slice");
    const float sh = resize_scale(Hi, Ho, mode), sw = resize_scale(Wi, Wo, mode);
    cudaStream_t st = as_stream(stream);
    const bool x2 = mode == ARSEG_RESIZE_BILINEAR && Ho == 2 * Hi && Wo == 2 * Wi && ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0);
    if (x2 && dtype == ARSEG_F32 && C % 4 == 0 && dcs % 4 == 0 && dco % 4 == 0) {
        upsample2x_launch<float, 4>((const float*)src, (float*)dst, N, Hi, Wi, C, dcs, dco, st);
    } else if (x2 && dtype == ARSEG_BF16 && C % 8 == 0 && dcs % 8 == 0 && dco % 8 == 0) {
        upsample2x_launch<__nv_bfloat16, 8>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, N, Hi, Wi, C, dcs, dco, st);
    } else if (x2 && dtype == ARSEG_F16 && C % 8 == 0 && dcs % 8 == 0 && dco % 8 == 0) {
        upsample2x_launch<__half, 8>((const __half*)src, (__half*)dst, N, Hi, Wi, C, dcs, dco, st);
    } else if (dtype == ARSEG_F32) {
        const float* s = (const float*)src; float* d = (float*)dst;
        if (C % 4 == 0 && dcs % 4 == 0 && dco % 4 == 0 && ((uintptr_t)s % 16 == 0) && ((uintptr_t)d % 16 == 0)) {
            resize_nhwc_kernel<float, 4><<<resize_grid(N, Ho), 256, 0, st>>>(s, d, N, Hi, Wi, C, Ho, Wo, dcs, dco, mode, sh, sw);
        } else {
            resize_nhwc_kernel<float, 1><<<resize_grid(N, Ho), 256, 0, st>>>(s, d, N, Hi, Wi, C, Ho, Wo, dcs, dco, mode, sh, sw);
        }
    } else if (dtype == ARSEG_BF16) {
        const __nv_bfloat16* s = (const __nv_bfloat16*)src; __nv_bfloat16* d = (__nv_bfloat16*)dst;
        if (C % 8 == 0 && dcs % 8 == 0 && dco % 8 == 0 && ((uintptr_t)s % 16 == 0) && ((uintptr_t)d % 16 == 0)) {
            resize_nhwc_kernel<__nv_bfloat16, 8><<<resize_grid(N, Ho), 256, 0, st>>>(s, d, N, Hi, Wi, C, Ho, Wo, dcs, dco, mode, sh, sw);
        } else {
            resize_nhwc_kernel<__nv_bfloat16, 1><<<resize_grid(N, Ho), 256, 0, st>>>(s, d, N, Hi, Wi, C, Ho, Wo, dcs, dco, mode, sh, sw);
        }
    } else if (dtype == ARSEG_F16) {
        const __half* s = (const __half*)src; __half* d = (__half*)dst;
        if (C % 8 == 0 && dcs % 8 == 0 && dco % 8 == 0 && ((uintptr_t)s % 16 == 0) && ((uintptr_t)d % 16 == 0)) {
            resize_nhwc_kernel<__half, 8><<<resize_grid(N, Ho), 256, 0, st>>>(s, d, N, Hi, Wi, C, Ho, Wo, dcs, dco, mode, sh, sw);
        } else {
            resize_nhwc_kernel<__half, 1><<<resize_grid(N, Ho), 256, 0, st>>>(s, d, N, Hi, Wi, C, Ho, Wo, dcs, dco, mode, sh, sw);
        }
    } else ARSEG_UNSUPPORTED("resize_nhwc: dtype %d", dtype);
    ARSEG_CHECK_LAUNCH("resize_nhwc");
    return ARSEG_OK;
}

int arseg_adaptive_avgpool_nhwc(const void* in, void* out, int dtype, int out_dtype, int N, int H, int W, int C, int Ho,
                                int Wo, arseg_stream_t stream) {
    ARSEG_REQUIRE(out_dtype == dtype || out_dtype == ARSEG_F32, "adaptive_avgpool: out dtype must be the input dtype or fp32");
    const int of32 = (out_dtype == ARSEG_F32) ? 1 : 0;
    ARSEG_REQUIRE(in && out && N > 0 && H > 0 && W > 0 && C > 0 && Ho > 0 && Wo > 0, "adaptive_avgpool: bad args");
    dim3 grid(N * Ho * Wo, ceil_div(C, 32)), block(32, 8);
    if (dtype == ARSEG_F32)
        adaptive_pool_kernel<float, false><<<grid, block, 0, as_stream(stream)>>>((const float*)in, out, N, H, W, C, Ho, Wo, of32);
    else if (dtype == ARSEG_BF16)
        adaptive_pool_kernel<__nv_bfloat16, false><<<grid, block, 0, as_stream(stream)>>>((const __nv_bfloat16*)in, out, N, H, W, C, Ho, Wo, of32);
    else if (dtype == ARSEG_F16)
        adaptive_pool_kernel<__half, false><<<grid, block, 0, as_stream(stream)>>>((const __half*)in, out, N, H, W, C, Ho, Wo, of32);
    else ARSEG_UNSUPPORTED("adaptive_avgpool: dtype %d", dtype);
    ARSEG_CHECK_LAUNCH("adaptive_avgpool");
    return ARSEG_OK;
}

int arseg_global_maxpool_nhwc(const void* in, float* out, int dtype, int N, int H, int W, int C, arseg_stream_t stream) {
    ARSEG_REQUIRE(in && out && N > 0 && H > 0 && W > 0 && C > 0, "global_maxpool: bad args");
    dim3 grid(N, ceil_div(C, 32)), block(32, 8);
    if (dtype == ARSEG_F32)
        adaptive_pool_kernel<float, true><<<grid, block, 0, as_stream(stream)>>>((const float*)in, out, N, H, W, C, 1, 1, 1);
    else if (dtype == ARSEG_BF16)
        adaptive_pool_kernel<__nv_bfloat16, true><<<grid, block, 0, as_stream(stream)>>>((const __nv_bfloat16*)in, out, N, H, W, C, 1, 1, 1);
    else if (dtype == ARSEG_F16)
        adaptive_pool_kernel<__half, true><<<grid, block, 0, as_stream(stream)>>>((const __half*)in, out, N, H, W, C, 1, 1, 1);
    else ARSEG_UNSUPPORTED("global_maxpool: dtype %d", dtype);
    ARSEG_CHECK_LAUNCH("global_maxpool");
    return ARSEG_OK;
}

int arseg_linear_f32(const float* x, const float* w, const float* b, float* y, int N, int K, int M, int relu,
                     arseg_stream_t stream) {
    ARSEG_REQUIRE(x && w && y && N > 0 && K > 0 && M > 0, "linear: bad args");
    const long long threads = (long long)N * M * 32;
    linear_kernel<<<(int)ceil_div_ll(threads, 256), 256, 0, as_stream(stream)>>>(x, w, b, y, N, K, M, relu);
    ARSEG_CHECK_LAUNCH("linear");
    return ARSEG_OK;
}

int arseg_gate_nhwc(const void* feat, const float* gate, const float* gs, const float* gb, int add_identity,
                    const float* add_chan, const void* add_pix, void* out, int dtype, int N, int H, int W, int C,
                    arseg_stream_t stream) {
    ARSEG_REQUIRE(feat && gate && out && N > 0 && H > 0 && W > 0 && C > 0, "gate: bad args");
    const long long total = (long long)N * H * W * C;
    if (dtype == ARSEG_F32)
        gate_kernel<float><<<grid_1d(total, 256), 256, 0, as_stream(stream)>>>(
            (const float*)feat, gate, gs, gb, add_identity, add_chan, (const float*)add_pix, (float*)out, N, (long long)H * W, C);
    else if (dtype == ARSEG_BF16)
        gate_kernel<__nv_bfloat16><<<grid_1d(total, 256), 256, 0, as_stream(stream)>>>(
            (const __nv_bfloat16*)feat, gate, gs, gb, add_identity, add_chan, (const __nv_bfloat16*)add_pix,
            (__nv_bfloat16*)out, N, (long long)H * W, C);
    else if (dtype == ARSEG_F16)
        gate_kernel<__half><<<grid_1d(total, 256), 256, 0, as_stream(stream)>>>(
            (const __half*)feat, gate, gs, gb, add_identity, add_chan, (const __half*)add_pix,
            (__half*)out, N, (long long)H * W, C);
    else ARSEG_UNSUPPORTED("gate: dtype %d", dtype);
    ARSEG_CHECK_LAUNCH("gate");
    return ARSEG_OK;
}

int arseg_maxpool3x3s2_nhwc(const void* in, void* out, int dtype, int N, int H, int W, int C, arseg_stream_t stream) {
    ARSEG_REQUIRE(in && out && N > 0 && H > 0 && W > 0 && C > 0, "maxpool: bad args");
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    cudaStream_t st = as_stream(stream);
    if (dtype == ARSEG_F32) {
        if (C % 4 == 0) {
            long long total = (long long)N * Ho * Wo * (C / 4);
            maxpool_kernel<float, 4><<<grid_1d(total, 256), 256, 0, st>>>((const float*)in, (float*)out, N, H, W, C, Ho, Wo);
        } else {
            long long total = (long long)N * Ho * Wo * C;
            maxpool_kernel<float, 1><<<grid_1d(total, 256), 256, 0, st>>>((const float*)in, (float*)out, N, H, W, C, Ho, Wo);
        }
    } else if (dtype == ARSEG_BF16) {
        if (C % 8 == 0) {
            long long total = (long long)N * Ho * Wo * (C / 8);
            maxpool_kernel<__nv_bfloat16, 8><<<grid_1d(total, 256), 256, 0, st>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, N, H, W, C, Ho, Wo);
        } else {
            long long total = (long long)N * Ho * Wo * C;
            maxpool_kernel<__nv_bfloat16, 1><<<grid_1d(total, 256), 256, 0, st>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, N, H, W, C, Ho, Wo);
        }
    } else if (dtype == ARSEG_F16) {
        if (C % 8 == 0) {
            long long total = (long long)N * Ho * Wo * (C / 8);
            maxpool_kernel<__half, 8><<<grid_1d(total, 256), 256, 0, st>>>((const __half*)in, (__half*)out, N, H, W, C, Ho, Wo);
        } else {
            long long total = (long long)N * Ho * Wo * C;
            maxpool_kernel<__half, 1><<<grid_1d(total, 256), 256, 0, st>>>((const __half*)in, (__half*)out, N, H, W, C, Ho, Wo);
        }
    } else ARSEG_UNSUPPORTED("maxpool: dtype %d", dtype);
    ARSEG_CHECK_LAUNCH("maxpool");
    return ARSEG_OK;
}

int arseg_conv_stem7x7s2(const float* in, const float* w, const float* scale, const float* shift, void* out,
                         int out_dtype, int N, int H, int W, int Cout, arseg_stream_t stream) {
    ARSEG_REQUIRE(in && w && scale && shift && out && N > 0 && H > 0 && W > 0 && Cout > 0, "stem: bad args");
    ARSEG_REQUIRE(Cout <= 64, "stem: Cout=%d > 64", Cout);
    const int Ho = (H + 6 - 7) / 2 + 1, Wo = (W + 6 - 7) / 2 + 1;
    // 16-bit plans: tensor-core stem (stem_mma.cu); ARSEG_STEM_SIMT=1 keeps the CUDA-core kernel for A/B runs
    if ((out_dtype == ARSEG_F16 || out_dtype == ARSEG_BF16) && Cout == 64) {
        const char* e = getenv("ARSEG_STEM_SIMT");
        if (!(e && atoi(e) != 0)) return stem_mma_launch(in, w, scale, shift, out, out_dtype, N, H, W, Ho, Wo, as_stream(stream));
    }
    dim3 grid(ceil_div(Wo, STEM_TW), ceil_div(Ho, STEM_TH), N);
    ARSEG_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "stem: dims too large");
    static bool configured[64] = {false};
    int dev = 0;
    ARSEG_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        ARSEG_CUDA(cudaFuncSetAttribute(stem_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STEM_SMEM));
        ARSEG_CUDA(cudaFuncSetAttribute(stem_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STEM_SMEM));
        ARSEG_CUDA(cudaFuncSetAttribute(stem_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STEM_SMEM));
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    if (out_dtype == ARSEG_F32)
        stem_kernel<float><<<grid, 256, STEM_SMEM, as_stream(stream)>>>(in, w, scale, shift, (float*)out, H, W, Ho, Wo, Cout);
    else if (out_dtype == ARSEG_BF16)
        stem_kernel<__nv_bfloat16><<<grid, 256, STEM_SMEM, as_stream(stream)>>>(in, w, scale, shift, (__nv_bfloat16*)out, H, W, Ho, Wo, Cout);
    else if (out_dtype == ARSEG_F16)
        stem_kernel<__half><<<grid, 256, STEM_SMEM, as_stream(stream)>>>(in, w, scale, shift, (__half*)out, H, W, Ho, Wo, Cout);
    else ARSEG_UNSUPPORTED("stem: dtype %d", out_dtype);
    ARSEG_CHECK_LAUNCH("stem");
    return ARSEG_OK;
}

}  // extern "C"
