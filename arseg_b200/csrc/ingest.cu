// ingest.cu -- the data formats on either side of the non-keyframe path, on the device (sm_100a):
//   * frame ingest: uint8 HWC frames (what PIL / cv2 decode to) -> transforms.ToTensor + Normalize
//     (dataset/camvid.py:182-185) -> the bilinear (align_corners=True) LR down-scale of evaluation.py:186-188, one kernel,
//     fp32 NCHW out.  A step then moves 1 byte per sample over PCIe instead of 4.
//   * mergeMotion (pre-process/generate_compressed_dataset_camvid.py:6-56): chains the per-frame HEVC MV maps dumped by
//     the patched dec265 (short[H][W][3] = mvx, mvy quarter-pel, refIdx) back to the keyframe, frame by frame; every frame is
//     one pixel-parallel launch (its parents are already resolved, so one hop suffices), output = the int16 [H,W,2]
//     quarter-pel wire format the dataset reads (dataset/camvid.py:624-626).
#include "common.cuh"

namespace arseg {

// ToTensor (u8 -> f32, / 255) then Normalize ((x - mean) / std), as torchvision does it: fp32 divide, subtract, divide
__device__ __forceinline__ float norm_u8(uint8_t v, float mean, float stdv) { return ((float)v / 255.f - mean) / stdv; }

__global__ void __launch_bounds__(128) frame_ingest_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, int N, int Hi, int Wi,
                                                           int Ho, int Wo, int mode, float sh, float sw, float m0, float m1, float m2,
                                                           float s0, float s1, float s2) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, n = blockIdx.z;
    if (x >= Wo) return;
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    bilinear_src(sh, y, Hi, mode, y0, y1, ly0, ly1);
    bilinear_src(sw, x, Wi, mode, x0, x1, lx0, lx1);
    const uint8_t* s = src + (size_t)n * Hi * Wi * 3;
    const uint8_t* a = s + ((size_t)y0 * Wi + x0) * 3;
    const uint8_t* b = s + ((size_t)y0 * Wi + x1) * 3;
    const uint8_t* c = s + ((size_t)y1 * Wi + x0) * 3;
    const uint8_t* d = s + ((size_t)y1 * Wi + x1) * 3;
    const float mean[3] = {m0, m1, m2}, stdv[3] = {s0, s1, s2};
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        // the reference normalises the full-resolution frame first and interpolates the normalised values (ATen's bilinear formula)
        const float va = norm_u8(a[ch], mean[ch], stdv[ch]), vb = norm_u8(b[ch], mean[ch], stdv[ch]);
        const float vc = norm_u8(c[ch], mean[ch], stdv[ch]), vd = norm_u8(d[ch], mean[ch], stdv[ch]);
        dst[(((size_t)n * 3 + ch) * Ho + y) * Wo + x] = ly0 * (lx0 * va + lx1 * vb) + ly1 * (lx0 * vc + lx1 * vd);
    }
}

// np.round(v / 4) for an integer v: round half to even (generate_compressed_dataset_camvid.py:26-27)
__device__ __forceinline__ int round_quarter(int v) {
    const int base = v >> 2, r = v & 3;            // floor division / remainder 0..3
    return r < 2 ? base : (r > 2 ? base + 1 : base + (base & 1));
}

// frame f (1-based).  dp[f'][y][x] = x | y << 12 | frame << 24: the ancestor of pixel (x, y) of frame f' -- always in frame 0
// once f' >= 1 has been processed; frame 0 itself is "unresolved" (the reference's -1 rows), so a parent in frame 0 links to
// the parent position itself (:45) and a parent in a later frame links to that parent's ancestor (:44).
__global__ void __launch_bounds__(256) merge_motion_kernel(const int16_t* __restrict__ map, uint32_t* __restrict__ dp, int16_t* __restrict__ out,
                                                           int f, int H, int W) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= H * W) return;
    const int j1 = i / W, k1 = i - j1 * W;
    int mvx = map[(size_t)i * 3], mvy = map[(size_t)i * 3 + 1], ref = map[(size_t)i * 3 + 2];
    if (ref < 0 || ref >= 3) { mvx = 0; mvy = 0; ref = 0; }          // intra blocks (:20-22; max_ref_num = 3)
    int j2 = j1 + round_quarter(mvy), k2 = k1 + round_quarter(mvx);
    const int f2 = max(0, f - ref - 1);
    j2 = min(max(j2, 0), H - 1); k2 = min(max(k2, 0), W - 1);
    uint32_t link = (uint32_t)k2 | ((uint32_t)j2 << 12) | ((uint32_t)f2 << 24);
    if (f2 >= 1) link = dp[(size_t)f2 * H * W + (size_t)j2 * W + k2];
    dp[(size_t)f * H * W + i] = link;
    const int ax = (int)(link & 0xfffu), ay = (int)((link >> 12) & 0xfffu);
    out[(size_t)i * 2] = (int16_t)((ax - k1) * 4);                    // :53-54, saved as np.short (:277)
    out[(size_t)i * 2 + 1] = (int16_t)((ay - j1) * 4);
}

}  // namespace arseg

using namespace arseg;

extern "C" {

int arseg_frame_ingest_u8(const uint8_t* src, const float* mean, const float* stdv, float* dst, int N, int Hi, int Wi, int Ho, int Wo,
                          int mode, arseg_stream_t stream) {
    ARSEG_REQUIRE(src && mean && stdv && dst && N > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, "frame_ingest: bad args");
    ARSEG_REQUIRE((mode == ARSEG_RESIZE_BILINEAR || mode == ARSEG_RESIZE_BILINEAR_AC) && Ho <= 65535 && N <= 65535, "frame_ingest: bad mode/size");
    const float sh = resize_scale(Hi, Ho, mode), sw = resize_scale(Wi, Wo, mode);
    dim3 block(128), grid(ceil_div(Wo, 128), Ho, N);
    frame_ingest_kernel<<<grid, block, 0, as_stream(stream)>>>(src, dst, N, Hi, Wi, Ho, Wo, mode, sh, sw, mean[0], mean[1], mean[2],
                                                              stdv[0], stdv[1], stdv[2]);
    ARSEG_CHECK_LAUNCH("frame_ingest");
    return ARSEG_OK;
}

size_t arseg_merge_motion_workspace_bytes(int F, int H, int W) { return (size_t)(F + 1) * H * W * sizeof(uint32_t); }

int arseg_merge_motion(const int16_t* maps, void* workspace, size_t workspace_bytes, int16_t* out, int F, int H, int W, arseg_stream_t stream) {
    ARSEG_REQUIRE(maps && workspace && out && F >= 1 && H > 0 && W > 0, "merge_motion: bad args");
    ARSEG_REQUIRE(H <= 4096 && W <= 4096 && F <= 127, "merge_motion: %dx%d x %d frames exceeds the packed link format (4096 x 4096 x 127)", H, W, F);
    ARSEG_REQUIRE(workspace_bytes >= arseg_merge_motion_workspace_bytes(F, H, W), "merge_motion: workspace too small");
    uint32_t* dp = reinterpret_cast<uint32_t*>(workspace);
    const size_t plane = (size_t)H * W;
    for (int f = 1; f <= F; ++f) {       // frame f reads the links of frames < f: one launch per frame, in stream order
        merge_motion_kernel<<<ceil_div(H * W, 256), 256, 0, as_stream(stream)>>>(maps + (size_t)(f - 1) * plane * 3, dp, out + (size_t)(f - 1) * plane * 2, f, H, W);
        ARSEG_CHECK_LAUNCH("merge_motion");
    }
    return ARSEG_OK;
}

}  // extern "C"
