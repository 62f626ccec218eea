/* rerevst_b200.h -- C ABI of the B200-native ReReVST stylization hot path.
 *
 * One shared library (rerevst-code_b200/csrc/librerevst_b200.so, sm_100a only).  Every entry
 * point takes plain device pointers, sizes and a CUDA stream (as void*); none takes a torch
 * type.  The reference (daooshee/ReReVST-Code @ b7f39f2) is pure Python/PyTorch and has no FFI
 * of its own, so each function below cites the reference Python code it replaces; the Python
 * host side (rerevst-code_b200/*.py) binds them with ctypes and mirrors the reference's module
 * API (TransformerNet / Stylization / warp / TemporalLoss).  INTEGRATION.md shows the binding.
 *
 * All functions return 0 on success, non-zero on error; rrv_last_error() gives the message.
 * Nothing here synchronises the host with the device unless stated.
 *
 * Activation layout ("planes"): NHWC, 16-bit, C a multiple of 8.  An fp32 activation v is
 * carried as hi = bf16(v) and lo = T(v - hi) in two separate tensors (T = bf16 or fp16, see
 * rrv_set_lo_format); the tensor-core path multiplies hi*Whi + hi*Wlo + lo*Whi ("x3" mode,
 * fp32-accurate).  With lo == NULL only hi is read/written ("bf16" mode, BASELINE config 3).
 */
#ifndef REREVST_B200_H
#define REREVST_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RRV_ABI_VERSION 3

/* ---- library ------------------------------------------------------------------------ */
int         rrv_abi_version(void);
const char* rrv_last_error(void);
/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
uint64_t    rrv_launch_count(void);
/* 0: lo plane is bf16 (default); 1: lo plane is fp16 (3 more mantissa bits, same cost). */
int         rrv_set_lo_format(int fmt);
int         rrv_get_lo_format(void);

/* ---- fused epilogue ------------------------------------------------------------------ */
/* Applied to every conv accumulator / pointwise input, in this order (each stage optional):
 *   v = acc + bias[c]                       nn.Conv2d bias
 *   v = act(v)                              1: ReLU (VGG, style_network_global.py:275-281)
 *                                           2: LeakyReLU(0.2) (ResidualBlock :106, KernelFilter :192)
 *   v = clamp((v - mean1[c]) * rstd1[c])    saved-stat InstanceNorm.forward, :43-57
 *   v = v + residual[n, y>>s, x>>s, c]      ResidualBlock shortcut :122 / KernelFilter :217
 *   v = clamp((v - mean2[c]) * rstd2[c])    Decoder.norm[i], :363
 *   v = v * scale[c] + shift[c]             AdaIN, :364
 * norm tables are float[4][C] = {mean, rstd, lo, hi}; affine is float[2][C] = {std, mean}. */
typedef struct rrv_epilogue {
    const float* bias;      /* [C] or NULL */
    int32_t      act;       /* 0 none, 1 relu, 2 leaky relu 0.2 */
    const float* norm1;     /* [4][C] or NULL */
    const void*  res_hi;    /* residual planes [rN][rH][rW][C] or NULL */
    const void*  res_lo;    /* NULL in bf16 mode */
    int32_t      res_shift; /* 0, or 1 when the residual lives at half resolution */
    int32_t      res_H, res_W;
    int64_t      res_batch_stride; /* elements; 0 broadcasts one residual over the batch */
    const float* norm2;     /* [4][C] or NULL */
    const float* affine;    /* [2][C] or NULL */
    int32_t      res_f32;   /* 1: res_hi points to an fp32 NHWC tensor (res_lo NULL): the 1x1 shortcut of a ResidualBlock is
                             * kept in fp32 -- the same bytes as two 16-bit planes, one add instead of unpack + two */
} rrv_epilogue;

/* RRV_OUT_BGR_F32 / RRV_OUT_BGR_U8 (the RGB head, Cout == 3): transform_back_image + tensor2numpy (test/framework.py:39-49)
 * and the crop of generate_real_video.py:167 run in the convolution's epilogue; the output is the HWC BGR frame
 * [N][crop_h][crop_w][3] in [0,255], fp32 (what Stylization.transfer returns) or uint8 (rint, what cv2.imwrite makes of it). */
enum { RRV_OUT_PLANES = 0, RRV_OUT_F32_NHWC = 1, RRV_OUT_F32_NCHW = 2, RRV_OUT_BGR_F32 = 3, RRV_OUT_BGR_U8 = 4 };
/* Operand terms of the fp32-accurate split (in_lo != NULL).  FULL: hi*Whi + hi*Wlo + lo*Whi (three MMAs per k-slice);
 * NO_WLO drops hi*Wlo (weights enter at bf16 precision), NO_ALO drops lo*Whi (activations enter at bf16 precision): two MMAs.
 * oracle/precision_sweep.py measures what each layer tolerates. */
enum { RRV_TERMS_FULL = 0, RRV_TERMS_NO_WLO = 1, RRV_TERMS_NO_ALO = 2 };
enum { RRV_IMPL_FFMA = 0, RRV_IMPL_TCGEN05 = 1 };

/* ---- convolution ---------------------------------------------------------------------- */
/* 3x3 (zero pad 1) or 1x1 stride-1 convolution over planes, with the nearest x2 upsample of
 * ResidualBlock.forward (style_network_global.py:112-113) optionally folded into the gather.
 * Replaces every nn.Conv2d / F.conv2d on the per-frame path except conv1_1:
 *   Encoder.slice[2..19] (:275-281), KernelFilter.down_sample/upsample/apply_filter (:181-217),
 *   ResidualBlock.conv1/conv2/conv_shortcut (:103-122), Decoder.slice1 (:341, :450). */
typedef struct rrv_conv {
    int32_t N, H, W;        /* output batch and spatial size */
    int32_t Cin, Cout;      /* Cin % 16 == 0; Cout % 8 == 0 for planes / NHWC outputs */
    int32_t ksize;          /* 1 or 3 */
    int32_t ups;            /* 1: input is [N][H/2][W/2][Cin] and is read as its nearest x2 upsample */
    const void* in_hi;
    const void* in_lo;      /* NULL: bf16 mode */
    const float* w_f32;     /* FFMA path: [ksize*ksize][Cin][Cout] fp32 */
    const void* w_tc;       /* tcgen05 path: blob produced by rrv_pack_weights_tc */
    rrv_epilogue ep;
    int32_t out_mode;       /* RRV_OUT_* */
    void* out_hi;           /* planes [N][H][W][Cout] */
    void* out_lo;           /* NULL: bf16 mode */
    float* out_f32;         /* NHWC [N][H][W][Cout], or NCHW [N][out_C][H][W] */
    int32_t out_C;          /* channels kept for RRV_OUT_F32_NCHW (3 for the RGB head) */
    int32_t Cin_used;       /* 0, or the number of leading input channels that meet non-zero weights (a tensor carried zero-padded
                             * to the next 64 channels); the rest is not multiplied.  The tcgen05 path takes Cin == 32 as it is
                             * (64-byte operand rows): KernelFilter's inner tensor needs no padding */
    int32_t pool;           /* 1: nn.MaxPool2d(2, 2) (vgg19.features[4|9|18], floor) fused behind bias + activation; the
                             * planes output is [N][H/2][W/2][Cout].  tcgen05 path, ups == 0, Cout % 32 == 0, no norm /
                             * residual / affine stage. */
    int32_t terms;          /* RRV_TERMS_* (tcgen05 path with lo planes) */
    void*   out_img;        /* RRV_OUT_BGR_*: [N][crop_h][crop_w][3] fp32 or uint8 */
    int32_t crop_y0, crop_x0, crop_h, crop_w;   /* RRV_OUT_BGR_*: window of the H x W result that is kept */
    double* stats;          /* optional double[5][Cout] = {count, sum, sum of squares, min, max} of the values this convolution
                             * WRITES, accumulated by its epilogue (atomics; initialise with rrv_stats_init, convert with
                             * rrv_stats_sums_to_m2): the statistics of the InstanceNorm that follows (style_network_frame.py:39-43,
                             * style_network_global.py:59-77) without another pass over the tensor.  tcgen05 path, epilogue of
                             * bias + activation only, not with pool. */
    int32_t stats_minmax;   /* 0: rows 3, 4 are left alone (frame mode has no clamp); 1: min / max as well (pre-pass) */
} rrv_conv;

int rrv_conv2d(const rrv_conv* p, int impl, void* stream);

/* Size in bytes of, and packing into, the tensor-core weight blob for a layer.
 * w_oihw: fp32 [Cout][Cin][k][k] (PyTorch layout) on the device. */
int64_t rrv_tc_weight_bytes(int Cin, int Cout, int ksize, int ups);
int rrv_pack_weights_tc(const float* w_oihw, int Cin, int Cout, int ksize, int ups, void* blob, void* stream);
/* Tuning knobs of the tensor-core kernel, for kernel development and the variant tests.  Process-global; the setters and the
 * convolution calls synchronise on one mutex and every rrv_conv2d call works from ONE snapshot of the knobs, so a call never
 * sees half an update -- but a knob flipped by one thread does change what other threads' later calls do.  Defaults 256, 2:
 * widest Cout tile; 128-pixel M tiles per weight tile in the row-reuse main loop (1 | 2). */
int rrv_tc_tune(int max_bn, int mt);
/* CTA pairs (tcgen05 cta_group::2, clusters of 2): enabled by default for Cout tiles >= min_bn (64). */
int rrv_tc_tune_pair(int enable, int min_bn);
/* Merge the three dx taps of a 3x3 convolution (and the column phases of a nearest-x2 one) into one MMA along N when
 * 3 (4) Cout <= 256: the 64-channel layers, whose cost is the A-operand fetch.  Enabled by default. */
int rrv_tc_tune_merge(int enable);
/* Programmatic dependent launch between consecutive convolutions of a stream (enabled by default): the next layer's CTAs are
 * scheduled while the last persistent round of this one still runs, do their prologue and wait.  A caller that keeps several
 * independent frames in flight on several streams turns it off while it captures their graphs: the waiting CTAs would hold exactly
 * the SMs the other frame's kernels could use (1080p, 2 frames in flight: 173 frames/s with, 180 without). */
int rrv_tc_tune_pdl(int enable);
/* Measurement only (tools/timeline_frame.py): while a device buffer of nslots x 4 uint64 is set, every tensor-core convolution that
 * is launched -- or captured into a CUDA graph -- takes the next slot and records, in globaltimer nanoseconds, {first CTA start,
 * first CTA past its griddepcontrol.wait, first CTA end, last CTA end} with atomicMin / atomicMax (initialise a slot to
 * {~0, ~0, ~0, 0}).  nslots = 0 turns it off.  Process-global like the tuning knobs. */
int rrv_tc_timeline(void* buf, int nslots);
/* KernelFilter fold (apply_filter, style_network_global.py:194-217; per frame in test/style_network_frame.py:97-105): the two
 * predicted 32x32 matrices wf1, wf2 ([out][in], fp32) are multiplied into the filter's down_sample (512 -> 32) and upsample
 * (32 -> 512) 3x3 weights (PyTorch OIHW fp32) and written as tensor-core blobs: down_blob = rrv_tc_weight_bytes(512, 32, 3, 0)
 * bytes for the 512 -> 32 convolution with down_bias[32] = wf1 . down_b; up_blob = rrv_tc_weight_bytes(32, 512, 3, 0) bytes for
 * the 32 -> 512 convolution. */
int rrv_fold_filter(const float* wf1, const float* wf2, const float* down_w, const float* down_b, const float* up_w,
                    void* down_blob, float* down_bias, void* up_blob, void* stream);
/* fp32 [Cout][Cin][k][k] -> [k*k][Cin_pad][Cout_pad] fp32 (zero padded) for the FFMA path. */
int rrv_pack_weights_f32(const float* w_oihw, int Cin, int Cout, int ksize, int Cin_pad, int Cout_pad,
                         float* out, void* stream);

/* ---- first layer ---------------------------------------------------------------------- */
/* conv1_1 (3 -> 64, 3x3, ReLU) straight from the network input, with the conversions the
 * reference does before it fused in:
 *   src_kind 0: fp32 NCHW normalised RGB [N][3][H][W]  (TransformerNet.forward input, :499-501)
 *   src_kind 1: uint8 HWC BGR [N][H][W][3]             (numpy2tensor + transform_image,
 *                                                       test/framework.py:26-35)
 * gray != 0 applies TransformerNet.RGB2Gray (:487-497) first (Encoder path); gray == 0 is the
 * EncoderStyle / Vgg19 path.  w: fp32 [64][3][3][3] (PyTorch layout), bias [64].
 * Output: planes [N][H][W][64] and/or fp32 NHWC (out_f32 may be NULL). */
int rrv_first_layer(const void* src, int src_kind, int gray, int N, int H, int W,
                    const float* w, const float* bias,
                    void* out_hi, void* out_lo, float* out_f32, void* stream);

/* ---- pointwise -------------------------------------------------------------------------- */
/* nn.MaxPool2d(2, 2) of vgg19.features[4|9|18] on planes: [N][H][W][C] -> [N][H/2][W/2][C]. */
int rrv_maxpool2x2(const void* in_hi, const void* in_lo, int N, int H, int W, int C,
                   void* out_hi, void* out_lo, void* stream);
/* The epilogue chain alone over an fp32 NHWC tensor (pre-pass: normalise after the global
 * statistics are known).  C / 4 must divide 256 (C = 32, 64, 128, 256, 512, 1024: a thread keeps one channel group).  in_batch_stride == 0 broadcasts one input over the batch (quirk Q1,
 * KernelFilter.compute :223-230).  Output per out_mode (planes or fp32 NHWC). */
int rrv_pointwise(const float* in, int64_t in_batch_stride, int N, int H, int W, int C,
                  const rrv_epilogue* ep, int out_mode, void* out_hi, void* out_lo, float* out_f32,
                  void* stream);
/* Same, and the per-channel {sum, sum of squares[, min, max]} of the values written go into `stats` (see rrv_stats_init):
 * the statistics of the NEXT normalisation come out of the pass that applies the previous one. */
int rrv_pointwise_stats(const float* in, int64_t in_batch_stride, int N, int H, int W, int C,
                        const rrv_epilogue* ep, int out_mode, void* out_hi, void* out_lo, float* out_f32,
                        double* stats, int stats_minmax, void* stream);
/* planes -> fp32 NCHW (debug / feature export) and fp32 NCHW -> planes. */
int rrv_planes_to_nchw(const void* hi, const void* lo, int N, int H, int W, int C, float* out, void* stream);
int rrv_nchw_to_planes(const float* in, int N, int H, int W, int C, void* hi, void* lo, void* stream);
/* ReshapeTool.process (test/generate_real_video.py:66-83): cv2.copyMakeBorder(..., BORDER_REFLECT) of uint8 HWC frames on
 * the device, [N][H][W][3] -> [N][PH][PW][3] with the image at (top, left); the edge pixel is repeated. */
int rrv_reflect_pad_u8(const void* src, int N, int H, int W, int top, int left, int PH, int PW, void* dst, void* stream);
/* transform_back_image + tensor2numpy (test/framework.py:39-49): fp32 NCHW [N][3][H][W] ->
 * fp32 HWC BGR in [0,255], cropped to rows [y0, y0+h) and cols [x0, x0+w)
 * (generate_real_video.py:167). */
int rrv_postprocess_bgr(const float* in, int N, int H, int W, int y0, int x0, int h, int w,
                        float* out, void* stream);
/* Same, rounded to uint8 (rint of the clamped value: what cv2.imwrite's float32 -> uint8 conversion stores,
 * generate_real_video.py:170); a quarter of the bytes to download. */
int rrv_postprocess_bgr_u8(const float* in, int N, int H, int W, int y0, int x0, int h, int w,
                           uint8_t* out, void* stream);

/* ---- backward of the frozen Vgg19 loss network -------------------------------------------- */
/* train/train.py:376-414: Loss.backward() runs first through Vgg19 (train/style_networks.py:284-314, requires_grad False), which
 * needs data gradients only.  The data gradient of a stride-1 3x3 convolution is rrv_conv2d itself on weights with the channel
 * axes swapped and the taps rotated by 180 degrees (repacked once by the host side); the two passes below are the rest.
 * rrv_relu_backward: out planes [n] = (y > 0) ? g + g2 : 0 -- ReLU backward fused with the conversion to operand planes
 *   (g, y fp32 NHWC; g2 NULL or a second gradient arriving at the same tensor, e.g. a loss tap; out_lo NULL in bf16 mode).
 * rrv_maxpool2x2_backward: gx [N][H][W][C] from g [N][H/2][W/2][C] and the pool's input y: the first maximum of each window
 *   in row-major order receives the gradient (ATen's choice); a dropped odd last row / column gets zero. */
int rrv_relu_backward(const float* g, const float* g2, const float* y, int64_t n, void* out_hi, void* out_lo, void* stream);
int rrv_maxpool2x2_backward(const float* g, const float* y, int N, int H, int W, int C, float* gx, void* stream);

/* ---- statistics ------------------------------------------------------------------------ */
/* Per-channel partial statistics of an fp32 NHWC tensor over (N,H,W):
 * part = double[5][C] = {count, sum, M2 about the local mean, min, max}.  Two passes over the
 * data like InstanceNorm.compute (style_network_global.py:59-77).  Partials from several ranks
 * are merged on the device with rrv_stats_merge (Chan's parallel formula, fixed order). */
int rrv_channel_stats(const float* x, int64_t npix, int C, double* part, void* stream);
/* One-pass variant: a producing kernel (rrv_conv.stats, rrv_pointwise's stats argument) accumulates {sum, sum of squares,
 * min, max} into a partial initialised by rrv_stats_init (count = number of pixels the producer will write); afterwards
 * rrv_stats_sums_to_m2 turns row 2 into M2 about the mean, which makes it a partial like rrv_channel_stats' (mergeable,
 * finalizable). */
int rrv_stats_init(double* part, int C, double count, void* stream);
int rrv_stats_sums_to_m2(double* part, int C, void* stream);
int rrv_stats_merge(const double* parts, int nparts, int C, double* merged, void* stream);
/* kind 0: saved-stat table float[4][C] = {mean, rsqrt(M2/n + eps), (min-mean)*rstd, (max-mean)*rstd}
 *         (InstanceNorm.compute, biased variance, eps 1e-8);
 * kind 1: float[2][C] = {sqrt(M2/(n-1) + eps), mean} = AdaIN {scale, shift}
 *         (EncoderStyle.cal_mean_std :304-315, unbiased variance, eps 1e-5);
 * kind 2: float[C] = mean (FilterPredictor spatial/batch mean :163-167);
 * kind 3: float[4][C] = {mean, rstd, -inf, +inf}: frame-mode InstanceNorm (no clamp,
 *         style_network_frame.py:39-43).
 * kind | 16: row 2 of `part` still holds sum(x^2) (straight from a one-pass producer, no rrv_stats_sums_to_m2 in between). */
int rrv_stats_finalize(const double* part, int C, int kind, float eps, float* out, void* stream);
/* FilterPredictor's content / style term (style_network_global.py:150-172; per frame in style_network_frame.py:53-62) is
 * mean_{n,y,x}(conv3x3(x) + b): only the MEAN of a 512 -> 32 convolution.  The convolution is linear, so that mean follows from
 * nine per-channel sums of x (total, first / last row and column, corners: each tap reads everything but one border row and
 * column) and one dot product per output -- one read of x instead of the convolution.  x: planes [N][H][W][Cin]; w: fp32 OIHW
 * [Cout][Cin][3][3]; scratch: double[9][Cin]; part: double[5][Cout] = {N H W, sum over all output pixels, 0, 0, 0}, a partial
 * that rrv_stats_merge / rrv_stats_finalize(kind 2) accept. */
int rrv_conv3x3_output_sum(const void* in_hi, const void* in_lo, int N, int H, int W, int Cin, const float* w_oihw,
                           const float* bias, int Cout, double* scratch, double* part, void* stream);
/* FilterPredictor FC (:157,169): out[1024] = W[1024][64] . concat(c[32], s[32]) + b. */
int rrv_filter_fc(const float* w, const float* b, const float* c_mean, const float* s_mean,
                  float* out, void* stream);

/* ---- optical-flow warp ------------------------------------------------------------------ */
/* warp(x, flo, padding_mode='border') of train/loss_networks.py:20-38: nearest-neighbour
 * grid_sample with border padding, align_corners=False.  x, out: fp32 [B][C][H][W];
 * flo: fp32 [B][2][H][W].  src_index (optional, int32 [B][H][W][2] = {iy, ix}) receives the
 * integer source coordinates (bit-exact contract). */
int rrv_warp_nearest_border(const float* x, const float* flo, int B, int C, int H, int W,
                            float* out, int32_t* src_index, void* stream);
/* TemporalLoss.forward (:106-111): warped = warp(first, flo); *loss = mean|warped - second|.
 * loss_accum: device double[1] scratch; loss: device float[1]. */
int rrv_temporal_loss(const float* first, const float* second, const float* flo,
                      int B, int C, int H, int W, float* warped, double* loss_accum, float* loss,
                      void* stream);
/* Backward of warp w.r.t. x (nearest => zero gradient w.r.t. the grid): grad_x[src] += grad_out. */
int rrv_warp_backward(const float* grad_out, const float* flo, int B, int C, int H, int W,
                      float* grad_x, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* REREVST_B200_H */
