// local_attention.cu -- the literal `localAttention` operator boundary (model/attention.py:7-11) and
// warpFeature (evaluation.py:61-87) on the reference's own layouts (NCHW fp32, logits [N,H,W,kH*kW]).
// These are the un-fused drop-ins; the throughput path is the fused kernel in creff.cu.
#include "common.cuh"

namespace arseg {

constexpr int LA_MAXK = 15;  // max kW / kH handled by the register arrays below

// S[n,y,x,i*kW+j] = sum_c Q[n,c,y,x] * K[n,c,y+i-rh,x+j-rw]   (OOB taps -> exactly 0)
// block (32 x-lanes, kH tap rows); each thread keeps kW accumulators.
__global__ void similar_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k, float* __restrict__ out,
                                   int C, int H, int W, int kH, int kW) {
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y, n = blockIdx.z, i = threadIdx.y;
    if (x >= W) return;
    const int rh = kH / 2, rw = kW / 2;
    const int yy = y + i - rh;
    float acc[LA_MAXK];
#pragma unroll
    for (int j = 0; j < LA_MAXK; ++j) acc[j] = 0.f;
    if (yy >= 0 && yy < H) {
        const size_t plane = (size_t)H * W;
        const float* qp = q + (size_t)n * C * plane + (size_t)y * W + x;
        const float* kp = k + (size_t)n * C * plane + (size_t)yy * W;
        for (int c = 0; c < C; ++c) {
            const float qv = qp[c * plane];
#pragma unroll
            for (int j = 0; j < LA_MAXK; ++j) {
                const int xx = x + j - rw;
                if (j < kW && xx >= 0 && xx < W) acc[j] = fmaf(qv, kp[c * plane + xx], acc[j]);
            }
        }
    }
    float* o = out + (((size_t)n * H + y) * W + x) * (kH * kW) + i * kW;
#pragma unroll
    for (int j = 0; j < LA_MAXK; ++j)
        if (j < kW) o[j] = acc[j];
}

// O[n,c,y,x] = sum_ij V[n,c,y+i-rh,x+j-rw] * A[n,y,x,i*kW+j]
// block (32 x-lanes, 8 channel groups); the attention row of each pixel is staged in smem.
__global__ void weighting_fwd_kernel(const float* __restrict__ v, const float* __restrict__ a, float* __restrict__ out,
                                     int C, int H, int W, int kH, int kW) {
    extern __shared__ float s_a[];  // [32][T+1]
    const int T = kH * kW;
    const int x0 = blockIdx.x * 32, y = blockIdx.y, n = blockIdx.z;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int npx = min(32, W - x0);
    const float* ap = a + (((size_t)n * H + y) * W + x0) * T;
    for (int i = tid; i < npx * T; i += 256) s_a[(i / T) * (T + 1) + i % T] = ap[i];
    __syncthreads();
    const int x = x0 + threadIdx.x;
    if (x >= W) return;
    const int rh = kH / 2, rw = kW / 2;
    const float* sa = s_a + threadIdx.x * (T + 1);
    const size_t plane = (size_t)H * W;
    for (int c = threadIdx.y; c < C; c += 8) {
        const float* vp = v + ((size_t)n * C + c) * plane;
        float acc = 0.f;
        for (int i = 0; i < kH; ++i) {
            const int yy = y + i - rh;
            if (yy < 0 || yy >= H) continue;
            for (int j = 0; j < kW; ++j) {
                const int xx = x + j - rw;
                if (xx < 0 || xx >= W) continue;
                acc = fmaf(vp[(size_t)yy * W + xx], sa[i * kW + j], acc);
            }
        }
        out[((size_t)n * C + c) * plane + (size_t)y * W + x] = acc;
    }
}

// out[n,c,y',x'] = sum_ij Wt[n,p,i*kW+j] * X[n,c,p],  p = (y'-i+rh, x'-j+rw) inside the image.
// (similar_backward(is_ori=False): X = x_ori, Wt = grad_out;  weighting_backward_ori: X = grad_out, Wt = x_weight)
__global__ void transposed_weighting_kernel(const float* __restrict__ xin, const float* __restrict__ wt,
                                            float* __restrict__ out, int C, int H, int W, int kH, int kW) {
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y, n = blockIdx.z;
    if (x >= W) return;
    const int T = kH * kW, rh = kH / 2, rw = kW / 2;
    const size_t plane = (size_t)H * W;
    for (int c = threadIdx.y; c < C; c += 8) {
        const float* xp = xin + ((size_t)n * C + c) * plane;
        float acc = 0.f;
        for (int i = 0; i < kH; ++i) {
            const int py = y - i + rh;
            if (py < 0 || py >= H) continue;
            for (int j = 0; j < kW; ++j) {
                const int px = x - j + rw;
                if (px < 0 || px >= W) continue;
                acc = fmaf(wt[(((size_t)n * H + py) * W + px) * T + i * kW + j], xp[(size_t)py * W + px], acc);
            }
        }
        out[((size_t)n * C + c) * plane + (size_t)y * W + x] = acc;
    }
}

// ---- warpFeature (evaluation.py:61-87); warp_source_pos lives in common.cuh ------------------
template <typename TF>
__global__ void warp_feature_kernel(const float* __restrict__ feat, const TF* __restrict__ flow, float* __restrict__ out,
                                    int C, int H, int W) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, n = blockIdx.z;
    if (x >= W) return;
    const TF* f = flow + (((size_t)n * H + y) * W + x) * 2;
    float ix, iy;
    if (sizeof(TF) == 8) {
        warp_source_pos(x, y, (double)f[0], (double)f[1], W, H, ix, iy);
    } else {
        // float32 flow: the reference arithmetic then runs in fp32 (grid + flow stays fp32)
        const float vx = (float)x + (float)f[0], vy = (float)y + (float)f[1];
        const float gx = 2.0f * vx / (float)max(W - 1, 1) - 1.0f;
        const float gy = 2.0f * vy / (float)max(H - 1, 1) - 1.0f;
        ix = ((gx + 1.f) * (float)W - 1.f) / 2.f;
        iy = ((gy + 1.f) * (float)H - 1.f) / 2.f;
    }
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
    const float wnw = ((fx + 1.f) - ix) * ((fy + 1.f) - iy), wne = (ix - fx) * ((fy + 1.f) - iy);
    const float wsw = ((fx + 1.f) - ix) * (iy - fy), wse = (ix - fx) * (iy - fy);
    const bool vx0 = x0 >= 0 && x0 < W, vx1 = x1 >= 0 && x1 < W, vy0 = y0 >= 0 && y0 < H, vy1 = y1 >= 0 && y1 < H;
    const size_t plane = (size_t)H * W;
    for (int c = 0; c < C; ++c) {
        const float* p = feat + ((size_t)n * C + c) * plane;
        float acc = 0.f;
        if (vy0 && vx0) acc += p[(size_t)y0 * W + x0] * wnw;
        if (vy0 && vx1) acc += p[(size_t)y0 * W + x1] * wne;
        if (vy1 && vx0) acc += p[(size_t)y1 * W + x0] * wsw;
        if (vy1 && vx1) acc += p[(size_t)y1 * W + x1] * wse;
        out[((size_t)n * C + c) * plane + (size_t)y * W + x] = acc;
    }
}

static int check_la(const char* name, const void* a, const void* b, const void* c, int N, int C, int H, int W, int kH, int kW) {
    ARSEG_REQUIRE(a && b && c, "%s: null pointer", name);
    ARSEG_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0, "%s: bad shape", name);
    ARSEG_REQUIRE(kH > 0 && kW > 0 && (kH & 1) && (kW & 1) && kH <= LA_MAXK && kW <= LA_MAXK,
                  "%s: window %dx%d unsupported (odd, <= %d)", name, kH, kW, LA_MAXK);
    ARSEG_REQUIRE(H <= 65535 && N <= 65535, "%s: H or N too large", name);
    return ARSEG_OK;
}

}  // namespace arseg

using namespace arseg;

extern "C" {

int arseg_local_similar_fwd(const float* x_ori, const float* x_loc, float* out, int N, int C, int H, int W, int kH,
                            int kW, arseg_stream_t stream) {
    int rc = check_la("similar_forward", x_ori, x_loc, out, N, C, H, W, kH, kW);
    if (rc) return rc;
    dim3 grid(ceil_div(W, 32), H, N), block(32, kH);
    similar_fwd_kernel<<<grid, block, 0, as_stream(stream)>>>(x_ori, x_loc, out, C, H, W, kH, kW);
    ARSEG_CHECK_LAUNCH("similar_forward");
    return ARSEG_OK;
}

int arseg_local_weighting_fwd(const float* x_ori, const float* x_weight, float* out, int N, int C, int H, int W,
                              int kH, int kW, arseg_stream_t stream) {
    int rc = check_la("weighting_forward", x_ori, x_weight, out, N, C, H, W, kH, kW);
    if (rc) return rc;
    dim3 grid(ceil_div(W, 32), H, N), block(32, 8);
    const size_t smem = (size_t)32 * (kH * kW + 1) * sizeof(float);
    weighting_fwd_kernel<<<grid, block, smem, as_stream(stream)>>>(x_ori, x_weight, out, C, H, W, kH, kW);
    ARSEG_CHECK_LAUNCH("weighting_forward");
    return ARSEG_OK;
}

int arseg_local_similar_bwd(const float* x, const float* grad_out, float* grad_in, int N, int C, int H, int W, int kH,
                            int kW, int is_ori, arseg_stream_t stream) {
    int rc = check_la("similar_backward", x, grad_out, grad_in, N, C, H, W, kH, kW);
    if (rc) return rc;
    if (is_ori) {
        // dL/dx_ori[n,c,y,x] = sum_ij g[n,y,x,ij] * x_loc[n,c,y+i-r,x+j-r]  == weighting_forward(x_loc, g)
        return arseg_local_weighting_fwd(x, grad_out, grad_in, N, C, H, W, kH, kW, stream);
    }
    dim3 grid(ceil_div(W, 32), H, N), block(32, 8);
    transposed_weighting_kernel<<<grid, block, 0, as_stream(stream)>>>(x, grad_out, grad_in, C, H, W, kH, kW);
    ARSEG_CHECK_LAUNCH("similar_backward");
    return ARSEG_OK;
}

int arseg_local_weighting_bwd_ori(const float* x_weight, const float* grad_out, float* grad_ori, int N, int C, int H,
                                  int W, int kH, int kW, arseg_stream_t stream) {
    int rc = check_la("weighting_backward_ori", x_weight, grad_out, grad_ori, N, C, H, W, kH, kW);
    if (rc) return rc;
    dim3 grid(ceil_div(W, 32), H, N), block(32, 8);
    transposed_weighting_kernel<<<grid, block, 0, as_stream(stream)>>>(grad_out, x_weight, grad_ori, C, H, W, kH, kW);
    ARSEG_CHECK_LAUNCH("weighting_backward_ori");
    return ARSEG_OK;
}

int arseg_local_weighting_bwd_weight(const float* x_ori, const float* grad_out, float* grad_weight, int N, int C,
                                     int H, int W, int kH, int kW, arseg_stream_t stream) {
    // dL/dA[n,y,x,ij] = sum_c g[n,c,y,x] * V[n,c,y+i-r,x+j-r]  == similar_forward(g, V)
    return arseg_local_similar_fwd(grad_out, x_ori, grad_weight, N, C, H, W, kH, kW, stream);
}

int arseg_warp_feature_nchw(const float* feature, const void* flow, int flow_dtype, float* out, int B, int C, int H,
                            int W, arseg_stream_t stream) {
    ARSEG_REQUIRE(feature && flow && out && B > 0 && C > 0 && H > 0 && W > 0, "warpFeature: bad args");
    ARSEG_REQUIRE(H <= 65535 && B <= 65535, "warpFeature: H or B too large");
    dim3 grid(ceil_div(W, 128), H, B), block(128);
    if (flow_dtype == ARSEG_F64)
        warp_feature_kernel<double><<<grid, block, 0, as_stream(stream)>>>(feature, (const double*)flow, out, C, H, W);
    else if (flow_dtype == ARSEG_F32)
        warp_feature_kernel<float><<<grid, block, 0, as_stream(stream)>>>(feature, (const float*)flow, out, C, H, W);
    else ARSEG_UNSUPPORTED("warpFeature: flow dtype %d (want f32/f64)", flow_dtype);
    ARSEG_CHECK_LAUNCH("warpFeature");
    return ARSEG_OK;
}

}  // extern "C"
