/* arseg.h -- C ABI of libarseg_sm100a.so: the B200-native (sm_100a) kernels behind AR-Seg's
 * per-non-keyframe inference path.
 *
 * The reference (THU-LYJ-Lab/AR-Seg) is pure Python/PyTorch; its only native boundary on this path is
 * the un-vendored `localAttention` CUDA extension (5 functions, imported at model/attention.py:7-11).
 * Everything else it reaches through ATen/cuDNN.  Each entry point below names the reference interface
 * it replaces (file:line relative to /root/reference).  INTEGRATION.md shows the reference-side binding.
 *
 * Conventions
 *   - All pointers are DEVICE pointers owned by the caller (PyTorch); the library never allocates,
 *     frees or retains them.  Work is enqueued on `stream` (a cudaStream_t passed as void*); no entry
 *     point synchronises.  The device is the one current in the calling thread.
 *   - Return value: 0 = ARSEG_OK, negative = error; arseg_last_error() returns a thread-local message.
 *     Nothing throws across the boundary.
 *   - "NCHW" tensors are the reference's API layout (fp32, contiguous); "NHWC" tensors are the internal
 *     activation layout (`dtype` = ARSEG_F32, ARSEG_F16 or ARSEG_BF16).
 */
#ifndef ARSEG_H_
#define ARSEG_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ARSEG_ABI_VERSION 6

enum { ARSEG_OK = 0, ARSEG_E_BADARG = -1, ARSEG_E_UNSUPPORTED = -2, ARSEG_E_CUDA = -3 };
enum { ARSEG_F32 = 0, ARSEG_BF16 = 1, ARSEG_F64 = 2, ARSEG_I16 = 3, ARSEG_F16 = 4 };
enum { ARSEG_ACT_NONE = 0, ARSEG_ACT_RELU = 1, ARSEG_ACT_PRELU = 2 };
/* resize modes: bilinear align_corners=False / True, legacy nearest */
enum { ARSEG_RESIZE_BILINEAR = 0, ARSEG_RESIZE_BILINEAR_AC = 1, ARSEG_RESIZE_NEAREST = 2 };
enum { ARSEG_NCHW = 0, ARSEG_NHWC = 1 };
/* convolution engines: SIMT fp32 (exact-order fp32 FMA), tcgen05 kind::tf32, tcgen05 kind::f16 (bf16) */
enum { ARSEG_CONV_SIMT_F32 = 1, ARSEG_CONV_TC_TF32 = 2, ARSEG_CONV_TC_BF16 = 3, ARSEG_CONV_TC_F16 = 4 };
/* CReFF engines: exact fp32 SIMT (any C, NCHW hr), or tensor-core window attention (f16 operands with fp32 accumulate
 * -- TF32-class error; NHWC hr and lr; C = 64: tcgen05 / TMEM for fp16 hr + lr, mma.sync for fp32 hr; C = 128..1024: mma.sync) */
enum { ARSEG_CREFF_EXACT_F32 = 0, ARSEG_CREFF_MMA_F16 = 1, ARSEG_CREFF_TCGEN05 = 2 };
/* arseg_creff_args.phase: the tcgen05 engine is a workspace pre-pass (MV warp of the keyframe feature, which depends only on
 * the keyframe feature and the MV fields) followed by the attention kernel; a caller may issue the two separately so that the
 * pre-pass overlaps the LR-branch CNN on another stream.  The other engines do nothing in ARSEG_CREFF_PHASE_PREPASS. */
enum { ARSEG_CREFF_PHASE_ALL = 0, ARSEG_CREFF_PHASE_PREPASS = 1, ARSEG_CREFF_PHASE_MAIN = 2 };

typedef void* arseg_stream_t;

int arseg_abi_version(void);
const char* arseg_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * 1. `localAttention` operator boundary (model/attention.py:7-11).  NCHW fp32, logits [N,H,W,kH*kW].
 * ---------------------------------------------------------------------------------------------- */
/* localAttention.similar_forward  (call site model/attention.py:18) */
int arseg_local_similar_fwd(const float* x_ori, const float* x_loc, float* out,
                            int N, int C, int H, int W, int kH, int kW, arseg_stream_t stream);
/* localAttention.weighting_forward (call site model/attention.py:38) */
int arseg_local_weighting_fwd(const float* x_ori, const float* x_weight, float* out,
                              int N, int C, int H, int W, int kH, int kW, arseg_stream_t stream);
/* localAttention.similar_backward (call sites model/attention.py:27-28); `x` is the OTHER operand */
int arseg_local_similar_bwd(const float* x, const float* grad_out, float* grad_in,
                            int N, int C, int H, int W, int kH, int kW, int is_ori, arseg_stream_t stream);
/* localAttention.weighting_backward_ori (model/attention.py:47) */
int arseg_local_weighting_bwd_ori(const float* x_weight, const float* grad_out, float* grad_ori,
                                  int N, int C, int H, int W, int kH, int kW, arseg_stream_t stream);
/* localAttention.weighting_backward_weight (model/attention.py:48) */
int arseg_local_weighting_bwd_weight(const float* x_ori, const float* grad_out, float* grad_weight,
                                     int N, int C, int H, int W, int kH, int kW, arseg_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * 2. evaluation.py helpers
 * ---------------------------------------------------------------------------------------------- */
/* warpFeature(feature, flow) (evaluation.py:61-87): feature [B,C,H,W] fp32, flow [B,H,W,2] f32|f64
 * (pixels, ch0=x ch1=y) at FEATURE resolution.  Grid normalised with the align_corners=True formula in
 * the flow's dtype, cast to fp32, sampled bilinear / zeros / align_corners=False. */
int arseg_warp_feature_nchw(const float* feature, const void* flow, int flow_dtype, float* out,
                            int B, int C, int H, int W, arseg_stream_t stream);
/* F.interpolate(x, [Ho,Wo], mode='bilinear', align_corners=ac) on `planes` fp32 planes
 * (evaluation.py:186-188 frame down-scale; :201-202 logits up-sampling). mode = ARSEG_RESIZE_* */
int arseg_resize_nchw_f32(const float* src, float* dst, int planes, int Hi, int Wi, int Ho, int Wo,
                          int mode, arseg_stream_t stream);
/* logits [N,ncls,Hi,Wi] fp32 NCHW -> bilinear resize to [Ho,Wo] (mode) -> argmax over classes
 * (evaluation.py:201-204; softmax is monotone and skipped).  out_logits may be NULL. */
int arseg_resize_argmax_nchw(const float* logits, float* out_logits, uint8_t* out_argmax,
                             int N, int ncls, int Hi, int Wi, int Ho, int Wo, int mode, arseg_stream_t stream);
/* nn.LogSoftmax over dim 1 of NCHW fp32 logits (model/pspnet.py:122,229; ctor default dim -> 1 for 4-D) */
int arseg_log_softmax_nchw(const float* in, float* out, int N, int ncls, int H, int W, arseg_stream_t stream);
/* hist[label*ncls+pred] += 1 over pixels with label != ignore (evaluation.py:205-209); hist int64[ncls*ncls] */
int arseg_confusion_hist(const uint8_t* pred, const int64_t* label, long long* hist,
                         long long npix, int ncls, int ignore_label, arseg_stream_t stream);

/* Frame ingest: uint8 HWC frames [N,Hi,Wi,3] (decoded PNG / video) -> transforms.ToTensor + Normalize(mean, std)
 * (dataset/camvid.py:182-185; dataset/cityscapes.py:88-93) -> bilinear resize to [Ho,Wo] (evaluation.py:186-188 uses
 * align_corners=True = ARSEG_RESIZE_BILINEAR_AC), fp32 NCHW [N,3,Ho,Wo].  mean / std: HOST float[3], read during the call. */
int arseg_frame_ingest_u8(const uint8_t* src, const float* mean, const float* stdv, float* dst, int N, int Hi, int Wi,
                          int Ho, int Wo, int mode, arseg_stream_t stream);
/* mergeMotion(workspace_dir, 0, F) (pre-process/generate_compressed_dataset_camvid.py:6-56): chains the per-frame MV maps
 * of the patched HEVC decoder (maps int16 [F,H,W,3] = mvx, mvy quarter-pel, refIdx of frames 1..F; frame 0 = the keyframe)
 * back to the keyframe.  out int16 [F,H,W,2] = the merged quarter-pel MV field of every frame 1..F (frame F's plane is the
 * `.bin` the dataset loads, dataset/camvid.py:624-626).  workspace: >= arseg_merge_motion_workspace_bytes(F,H,W) bytes. */
size_t arseg_merge_motion_workspace_bytes(int F, int H, int W);
int arseg_merge_motion(const int16_t* maps, void* workspace, size_t workspace_bytes, int16_t* out, int F, int H, int W,
                       arseg_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * 3. Layout conversion at the API edge
 * ---------------------------------------------------------------------------------------------- */
int arseg_nchw_to_nhwc(const float* src, void* dst, int dst_dtype, int N, int C, int H, int W, arseg_stream_t stream);
int arseg_nhwc_to_nchw(const void* src, int src_dtype, float* dst, int N, int C, int H, int W, arseg_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * 4. CNN operators on NHWC activations (replace the ATen/cuDNN calls made by model/extractors.py,
 *    model/pspnet.py, model/bisenet.py, model/pspnet_semseg.py)
 * ---------------------------------------------------------------------------------------------- */
/* conv1 7x7 s2 p3 (no bias) + BN(eval, folded) + ReLU  (model/extractors.py:112-114,148-150;
 * model/bisenet.py:72-74,85-87; SpatialPath.conv1 :329).  in: NCHW fp32 [N,3,H,W];
 * w: [64][7][7][3] fp32; out: NHWC [N,Ho,Wo,64] of out_dtype. */
int arseg_conv_stem7x7s2(const float* in, const float* w, const float* scale, const float* shift,
                         void* out, int out_dtype, int N, int H, int W, int Cout, arseg_stream_t stream);
/* nn.MaxPool2d(3, 2, 1) (model/extractors.py:115,151) NHWC */
int arseg_maxpool3x3s2_nhwc(const void* in, void* out, int dtype, int N, int H, int W, int C, arseg_stream_t stream);

typedef struct arseg_conv_desc {
    const void* in;        /* NHWC [N,Hi,Wi,Cin], dtype `dtype` */
    const void* w;         /* [Cout][KH][KW][Cin], same dtype as `in` (fp32 for SIMT/TF32, fp16 / bf16 for F16 / BF16) */
    const float* scale;    /* [Cout] folded BN scale, or NULL (=1) */
    const float* shift;    /* [Cout] folded BN shift + bias, or NULL (=0) */
    const void* residual;  /* NHWC [N,Ho,Wo,Cout] (dense), added before the activation, or NULL */
    void* out;             /* NHWC [N,Ho,Wo,out_cstride] written at channel offset out_coff */
    int dtype;             /* ARSEG_F32 | ARSEG_F16 | ARSEG_BF16 (in, w, residual, out) */
    int N, Hi, Wi, Cin, Cout, KH, KW, stride, pad, dil;
    int out_cstride, out_coff;
    int act;               /* ARSEG_ACT_* */
    float prelu_slope;
    int engine;            /* ARSEG_CONV_* */
    int out_f32;           /* 1: `out` is fp32 although dtype is a 16-bit type (the LR feature p handed to the CReFF
                              kernel stays fp32: it is the residual of model/attention.py:210); tcgen05 engines only */
} arseg_conv_desc;
/* nn.Conv2d (+BatchNorm2d eval) (+residual) (+ReLU/PReLU) as one implicit-GEMM kernel */
int arseg_conv2d_nhwc(const arseg_conv_desc* d, arseg_stream_t stream);

/* F.interpolate / F.upsample / nn.Upsample on NHWC; writes channels [dst_coff, dst_coff+C) of a
 * destination with channel stride dst_cstride (fuses torch.cat, model/pspnet.py:29-30) */
int arseg_resize_nhwc(const void* src, void* dst, int dtype, int N, int Hi, int Wi, int C,
                      int Ho, int Wo, int dst_cstride, int dst_coff, int mode, arseg_stream_t stream);
/* nn.AdaptiveAvgPool2d((Ho,Wo)) (model/pspnet.py:23; model/pspnet_semseg.py:18) and torch.mean(dim=(2,3))
 * (model/bisenet.py:254,292,390) on NHWC -> NHWC [N,Ho,Wo,C] (fp32 accumulate); out_dtype = dtype or ARSEG_F32 */
int arseg_adaptive_avgpool_nhwc(const void* in, void* out, int dtype, int out_dtype, int N, int H, int W, int C,
                                int Ho, int Wo, arseg_stream_t stream);
/* F.adaptive_max_pool2d(x,(1,1)) (model/pspnet.py:215) NHWC -> fp32 [N,C] */
int arseg_global_maxpool_nhwc(const void* in, float* out, int dtype, int N, int H, int W, int C, arseg_stream_t stream);
/* nn.Linear (+ReLU) on fp32 [N,K] x [M,K]^T (model/pspnet.py:124-128) */
int arseg_linear_f32(const float* x, const float* w, const float* b, float* y, int N, int K, int M, int relu,
                     arseg_stream_t stream);
/* ARM / FFM gating (model/bisenet.py:258-260, 396-398) and the branch sums (:295,:301):
 *   out[n,y,x,c] = feat * sigmoid(gate_scale[c]*gate[n,c] + gate_shift[c]) * 1 (+ feat if add_identity)
 *                  (+ add_chan[n,c]) (+ add_pix[n,y,x,c])
 * gate = output of the 1x1 conv on the pooled vector (fp32 [N,C]); gate_scale/shift = folded bn_atten. */
int arseg_gate_nhwc(const void* feat, const float* gate, const float* gate_scale, const float* gate_shift,
                    int add_identity, const float* add_chan, const void* add_pix, void* out, int dtype,
                    int N, int H, int W, int C, arseg_stream_t stream);

/* Pyramid pooling module in three launches (PSPModule model/pspnet.py:14-31; PPM model/pspnet_semseg.py:12-30).
 * bins[nlev] = pooled sizes (1,2,3,6); B = sum bins^2.  Host array, read during the call.
 *   pool:   NHWC `in` -> fp32 [N][B][C], level-major, every level's bins row-major (nn.AdaptiveAvgPool2d windows)
 *   conv:   per-level 1x1 conv w [nlev][Cout][C] fp32 (+ folded BN scale/shift [nlev][Cout] or NULL, + ReLU)
 *           on the pooled map -> fp32 [N][B][Cout]
 *   upsample_concat: out[..., stage_coff + l*Cout + c] = bilinear(level l, mode)[c]; out[..., feats_coff + c] = feats;
 *           out NHWC [N,H,W,nlev*Cout+Cf] of `dtype` (the two slices must tile the channel axis) */
int arseg_pyramid_pool_nhwc(const void* in, float* out, int dtype, int N, int H, int W, int C, const int* bins, int nlev,
                            arseg_stream_t stream);
int arseg_pyramid_conv1x1(const float* pooled, const float* w, const float* scale, const float* shift, int relu, float* out,
                          int N, int C, int Cout, const int* bins, int nlev, arseg_stream_t stream);
int arseg_pyramid_upsample_concat(const float* stage, const void* feats, void* out, int dtype, int N, int H, int W, int Cout,
                                  int Cf, int stage_coff, int feats_coff, int mode, const int* bins, int nlev,
                                  arseg_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * 5. Fused MV-warp + CReFF + classifier (evaluation.py:177-183 + model/attention.py:184-213 +
 *    final_conv/log-softmax of model/pspnet.py:226-229, + argmax of evaluation.py:204)
 * ---------------------------------------------------------------------------------------------- */
typedef struct arseg_creff_args {
    const void* hr;         /* keyframe feature p_HR, NCHW [Nhr,C,H,W] or NHWC [Nhr,H,W,C] (hr_layout), fp32 or (NHWC, C = 64,
                               tcgen05 engine only) fp16 (hr_dtype); Nhr = 1 (shared by all N frames) or N */
    int hr_shared;          /* 1: hr has batch 1 and is broadcast over the N frames */
    int hr_layout;          /* ARSEG_NCHW (required by ARSEG_CREFF_EXACT_F32) | ARSEG_NHWC (required by the tensor-core engines) */
    int engine;             /* ARSEG_CREFF_*; an engine that does not support the arguments returns ARSEG_E_UNSUPPORTED */
    const void* flow;       /* NULL: hr is already warped (MyAttention.forward semantics).
                               else MV field [N,Hm,Wm,2] at FRAME resolution: ARSEG_I16 quarter-pel
                               (dataset/camvid.py:624-626 on-disk format), or ARSEG_F32/ARSEG_F64 pixels */
    int flow_dtype, Hm, Wm;
    const void* lr;         /* LR feature p [N,C,h,w]: ARSEG_NCHW fp32, or ARSEG_NHWC of lr_dtype */
    int lr_layout, lr_dtype, h, w;
    const float *wq, *bq, *wk, *bk, *wv, *bv;   /* depthwise 3x3 [C,3,3] + bias [C] (model/attention.py:161-164) */
    const float *wcls, *bcls;                   /* final_conv [ncls,C], [ncls]; NULL = no classifier */
    int ncls, log_softmax;
    float* out_p;           /* fused p NCHW fp32 [N,C,H,W] or NULL */
    float* out_logits;      /* NCHW fp32 [N,ncls,H,W] or NULL */
    uint8_t* out_argmax;    /* [N,H,W] argmax over classes at feature resolution, or NULL */
    int N, C, H, W, k;
    void* workspace;        /* device scratch of >= arseg_creff_workspace_bytes(a) bytes, owned by the caller; may be NULL
                               when that function returns 0 (C = 64 and the exact engine need none) */
    size_t workspace_bytes;
    int hr_dtype;           /* ARSEG_F32 (0) | ARSEG_F16 (tcgen05 engine only; ARSEG_CREFF_MMA_F16 with an fp16 hr is routed to it) */
    int phase;              /* ARSEG_CREFF_PHASE_* (0 = the whole operator in one call) */
} arseg_creff_args;
/* Engines: ARSEG_CREFF_EXACT_F32 -- fp32 SIMT, any C multiple of 16, k in {3,5,7,9}.
 *          ARSEG_CREFF_MMA_F16   -- mma.sync column-marching engine (C = 64) / two-launch wide engine (C = 128 .. 1024).
 *          ARSEG_CREFF_TCGEN05   -- tcgen05 / TMEM engine: C = 64, k in {3,5,7}, NHWC fp32 or fp16 hr, NHWC fp16 lr, ncls <= 32;
 *                                   needs the workspace (MV-warped keyframe rows, ~ N x H x W x 128 bytes). */
int arseg_creff_fused_fwd(const arseg_creff_args* a, arseg_stream_t stream);
/* Scratch the call above needs for these arguments (ARSEG_CREFF_MMA_F16 with C = 128, 192, ... 1024: Q, K, V in fp16
 * and the lr_up residual in fp32 pass through it between the two launches of that engine; ARSEG_CREFF_TCGEN05: the
 * MV-warped keyframe rows in fp16 + the lr_up gather records); 0 otherwise. */
size_t arseg_creff_workspace_bytes(const arseg_creff_args* a);

#ifdef __cplusplus
}
#endif
#endif /* ARSEG_H_ */
