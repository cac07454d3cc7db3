// creff_mma_common.cuh -- device helpers shared by the tensor-core CReFF engines (creff_mma.cu: tile engine,
// creff_march.cu: column-marching engine): MMA/ldmatrix wrappers, MV-field arithmetic (evaluation.py:177-180),
// bilinear gather records for warpFeature (evaluation.py:61-87) and lr_up (model/attention.py:191).
#pragma once
#include "common.cuh"
#include <cuda_fp16.h>

namespace arseg {

constexpr int MC = 64;            // channels

struct CreffMmaParams {
    const float* hr; int hr_shared;
    const void* flow; int flow_dtype, Hm, Wm;
    const void* lr; int h, w;
    const float *wq, *bq, *wk, *bk, *wv, *bv, *wcls, *bcls;
    int ncls, log_softmax;
    float* out_p; float* out_logits; uint8_t* out_argmax;
    int N, C, H, W;
    int tiles_x, tiles_y;
    int seg_rows, ncols, nseg;   // column-marching engine: rows per segment, column strips, segments
    int dbg;                     // ARSEG_XTRACE builds only: bit 0 = D skips its convolutions, bit 1 = C skips its math, bit 2 = G skips loads
};

// ---------------------------------------------------------------------------------------------
// small PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float clamp_h(float v) { return fminf(fmaxf(v, -60000.f), 60000.f); }
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
// byte offset of channel `ch` of row `pos` in a swizzled [pos][64 x f16] tile
__device__ __forceinline__ uint32_t swz(int pos, int ch) {
    return (uint32_t)(pos * 128 + ((((ch >> 3) ^ pos) & 7) << 4) + (ch & 7) * 2);
}
__device__ __forceinline__ uint32_t swz_chunk(int pos, int chunk) { return (uint32_t)(pos * 128 + (((chunk ^ pos) & 7) << 4)); }

__device__ __forceinline__ double mflow_raw(const void* flow, int dtype, size_t idx) {
    if (dtype == ARSEG_I16) return (double)reinterpret_cast<const int16_t*>(flow)[idx] / 4.0;  // dataset/camvid.py:625
    if (dtype == ARSEG_F64) return reinterpret_cast<const double*>(flow)[idx];
    return (double)reinterpret_cast<const float*>(flow)[idx];
}
// evaluation.py:177-180 at one feature pixel (f64): flow * Hf/Hm, bilinear align_corners=True resize
__device__ __forceinline__ void mflow_at(const CreffMmaParams& p, int n, int fy, int fx, double& u, double& v) {
    const size_t base = (size_t)n * p.Hm * p.Wm;
    const double sc = (double)p.H / (double)p.Hm;
    if (p.Hm == p.H && p.Wm == p.W) {
        const size_t i = (base + (size_t)fy * p.Wm + fx) * 2;
        u = mflow_raw(p.flow, p.flow_dtype, i) * sc;
        v = mflow_raw(p.flow, p.flow_dtype, i + 1) * sc;
        return;
    }
    const double sh = p.H > 1 ? (double)(p.Hm - 1) / (double)(p.H - 1) : 0.0;
    const double sw = p.W > 1 ? (double)(p.Wm - 1) / (double)(p.W - 1) : 0.0;
    const double ry = sh * fy, rx = sw * fx;
    int ya = min((int)ry, p.Hm - 1), xa = min((int)rx, p.Wm - 1);
    const int yb = ya + (ya < p.Hm - 1 ? 1 : 0), xb = xa + (xa < p.Wm - 1 ? 1 : 0);
    const double ly1 = ry - ya, ly0 = 1.0 - ly1, lx1 = rx - xa, lx0 = 1.0 - lx1;
    double r[2];
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
        const double a = mflow_raw(p.flow, p.flow_dtype, (base + (size_t)ya * p.Wm + xa) * 2 + ch) * sc;
        const double b = mflow_raw(p.flow, p.flow_dtype, (base + (size_t)ya * p.Wm + xb) * 2 + ch) * sc;
        const double c = mflow_raw(p.flow, p.flow_dtype, (base + (size_t)yb * p.Wm + xa) * 2 + ch) * sc;
        const double d = mflow_raw(p.flow, p.flow_dtype, (base + (size_t)yb * p.Wm + xb) * 2 + ch) * sc;
        r[ch] = ly0 * (lx0 * a + lx1 * b) + ly1 * (lx0 * c + lx1 * d);
    }
    u = r[0]; v = r[1];
}

// Per-position gather record: 4 tap weights (invalid taps already zeroed) + packed source address:
// info = (pixel index of the NW tap, clamped into the image) << 2 | dx << 1 | dy, or -1 = "all zero".
struct PosRec { float4 w; int info; int cx, cy; };   // (cx, cy) = the clamped NW tap (info >> 2 = cy * width + cx)

// warped-hr sample (evaluation.py:61-87) at feature pixel (fy,fx); zero outside the image (depthwise padding)
// rcp_w / rcp_h > 0: 2/(W-1), 2/(H-1) in f64 -- the grid normalisation (evaluation.py:80-81) is then a multiply
// instead of an f64 divide (differs from the divide by <= 1 ulp of f64 BEFORE the cast to fp32, evaluation.py:83)
// mv_raw != nullptr: the position's MV was loaded by the caller (int16 quarter-pel pair at feature resolution).
// RCP (compile time): rcp_w / rcp_h are known to be > 0 -- the exact-divide path (two f64 divisions, ~300 instructions) is
// then not even instantiated (instruction-cache footprint of the warp-specialised engines).
template <bool RCP = false>
__device__ __forceinline__ PosRec pos_hr(const CreffMmaParams& p, int n, int fy, int fx, double rcp_w = 0.0, double rcp_h = 0.0,
                                         const int* mv_raw = nullptr) {
    PosRec r; r.w = make_float4(0.f, 0.f, 0.f, 0.f); r.info = -1; r.cx = r.cy = 0;
    if (fy < 0 || fy >= p.H || fx < 0 || fx >= p.W) return r;
    float ix = (float)fx, iy = (float)fy;
    if (p.flow) {
        double u, v;
        if (mv_raw) {   // dataset/camvid.py:625: int16 / 4.0; evaluation.py:177: * Hf/Hm (= 1 here)
            u = (double)(short)(*mv_raw & 0xffff) / 4.0;
            v = (double)(short)(*mv_raw >> 16) / 4.0;
        } else {
            mflow_at(p, n, fy, fx, u, v);
        }
        if (RCP || rcp_w > 0.0) {
            const float gx = (float)(((double)(float)fx + u) * rcp_w - 1.0), gy = (float)(((double)(float)fy + v) * rcp_h - 1.0);
            ix = ((gx + 1.f) * (float)p.W - 1.f) / 2.f;
            iy = ((gy + 1.f) * (float)p.H - 1.f) / 2.f;
        } else {
            warp_source_pos(fx, fy, u, v, p.W, p.H, ix, iy);
        }
    }
    if (!(ix > -1.5f && ix < (float)p.W + 0.5f && iy > -1.5f && iy < (float)p.H + 0.5f)) return r;  // all taps outside
    const float fxn = floorf(ix), fyn = floorf(iy);
    const int xa = (int)fxn, ya = (int)fyn, xb = xa + 1, yb = ya + 1;
    const bool vxa = xa >= 0 && xa < p.W, vxb = xb >= 0 && xb < p.W, vya = ya >= 0 && ya < p.H, vyb = yb >= 0 && yb < p.H;
    const float wxa = (fxn + 1.f) - ix, wxb = ix - fxn, wya = (fyn + 1.f) - iy, wyb = iy - fyn;   // grid_sample weights
    r.w.x = (vya && vxa) ? wxa * wya : 0.f;
    r.w.y = (vya && vxb) ? wxb * wya : 0.f;
    r.w.z = (vyb && vxa) ? wxa * wyb : 0.f;
    r.w.w = (vyb && vxb) ? wxb * wyb : 0.f;
    const int cxa = min(max(xa, 0), p.W - 1), cxb = min(max(xb, 0), p.W - 1);
    const int cya = min(max(ya, 0), p.H - 1), cyb = min(max(yb, 0), p.H - 1);
    r.info = ((cya * p.W + cxa) << 2) | ((cxb - cxa) << 1) | (cyb - cya);
    r.cx = cxa; r.cy = cya;
    return r;
}
// lr_up sample (model/attention.py:191, bilinear align_corners=True); zero outside the image
__device__ __forceinline__ PosRec pos_lr(const CreffMmaParams& p, float lsh, float lsw, int fy, int fx) {
    PosRec r; r.w = make_float4(0.f, 0.f, 0.f, 0.f); r.info = -1; r.cx = r.cy = 0;
    if (fy < 0 || fy >= p.H || fx < 0 || fx >= p.W) return r;
    int ya, yb, xa, xb; float lya, lyb, lxa, lxb;
    bilinear_src(lsh, fy, p.h, ARSEG_RESIZE_BILINEAR_AC, ya, yb, lya, lyb);
    bilinear_src(lsw, fx, p.w, ARSEG_RESIZE_BILINEAR_AC, xa, xb, lxa, lxb);
    r.w = make_float4(lya * lxa, lya * lxb, lyb * lxa, lyb * lxb);
    r.info = ((ya * p.w + xa) << 2) | ((xb - xa) << 1) | (yb - ya);
    r.cx = xa; r.cy = ya;
    return r;
}

// A gather record names a 2x2 source block that lies INSIDE the image (top-left pixel clamped to [0, W-2] x [0, H-2]) plus
// the four weights of its pixels, so the taps are base, base + one pixel, base + one row, base + one row + one pixel.
// PosRec names the taps by a clamped NW tap and dx / dy flags instead; where a flag is 0 both taps of that direction read
// the same pixel (image border), and their weights are merged onto whichever block column / row holds it.
__device__ __forceinline__ void rec_block_of(const PosRec& r, int Wimg, int Himg, float4& w, int& bx, int& by) {
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

// two consecutive channels of a source pixel as float2
__device__ __forceinline__ float2 ld2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
__device__ __forceinline__ float2 ld2(const __nv_bfloat16* p) {
    const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(p));
    return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}

constexpr int BAR_FULL = 1, BAR_EMPTY = 3, BAR_GATHER = 5;   // named barrier ids (0 = __syncthreads)
__device__ __forceinline__ void nbar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void nbar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// four consecutive channels of a source pixel as float4
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
    return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                       __uint_as_float(u.y & 0xffff0000u));
}
__device__ __forceinline__ float4 ld4(const __half* p) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float2 ld2(const __half* p) {
    const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(p));
    return __half22float2(*reinterpret_cast<const __half2*>(&u));
}
// (lo, hi) -> packed f16x2 with saturation to +-65504 (one F2FP.SATFINITE)
__device__ __forceinline__ uint32_t pack_h2_sat(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

}  // namespace arseg
