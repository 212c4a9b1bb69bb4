/* adb200.h — C ABI of libadb200.so: the B200 (sm_100a) kernels under the ADAM-Dehaze hot path.
 *
 * The reference (talha-alam/ADAM-Dehaze) has no FFI layer: its hot path is a set of torch.nn.Module.forward
 * bodies.  Each entry point below replaces the library calls one of those bodies makes; the
 * reference file:line it stands in for is cited at each declaration.  The Python modules in
 * adam_dehaze_b200/models/ (same names/signatures as the reference modules) bind these with ctypes — see
 * INTEGRATION.md for the stub a maintainer of the reference would add.
 *
 * Conventions (all entry points):
 *   - plain C types only: raw DEVICE pointers, ints, an explicit cudaStream_t passed as void*;
 *   - return 0 on success, a negative adb_status otherwise; adb_last_error() gives the thread-local message;
 *   - never allocate device memory, never synchronise the stream, never take ownership of a buffer;
 *   - activations are NHWC bf16 ("feature maps"), images are NCHW fp32 in [0,1] (the reference's tensor contract,
 *     data/dataset.py:97-99);
 *   - `n_dev` (nullable) is a device int holding the live image count of a routed bucket; when non-NULL a launch
 *     processes images [n_start, min(n_start + n, *n_dev)) and is a no-op beyond it, so routing needs no host
 *     round-trip (replaces the torch.any()/nonzero() syncs of models/routing.py:55-61).
 */
#ifndef ADB200_H
#define ADB200_H

#include <stdint.h>

#if defined(__GNUC__)
#define ADB_API __attribute__((visibility("default")))
#else
#define ADB_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  ADB_OK = 0,
  ADB_ERR_INVALID = -1,   /* bad argument / unsupported shape */
  ADB_ERR_CUDA = -2,      /* CUDA runtime / driver error */
  ADB_ERR_NO_DEVICE = -3, /* no sm_100 device or driver entry point missing */
  ADB_ERR_KERNEL = -4     /* in-kernel protocol time-out flag was raised */
} adb_status;

/* activation applied by the conv epilogue */
enum { ADB_ACT_NONE = 0, ADB_ACT_RELU = 1, ADB_ACT_TANH = 2, ADB_ACT_SIGMOID = 3,
       ADB_ACT_SIGMOID2 = 4 /* 2*sigmoid(v) - 1 (LowIntensityDehazeModel's (out - 0.5)*2, low_intensity.py:113); adb_img_head_* only */ };

/* convolution kinds (the geometry the implicit-GEMM producer walks) */
enum {
  ADB_CONV_S1 = 0,     /* kh x kw, stride 1, zero padding `pad` (nn.Conv2d in ConvBlock, base_model.py:11-13) */
  ADB_CONV_S2 = 1,     /* kh x kw, stride 2, zero padding `pad`; H and W even (encoder downsamples, medium:25,35; high:26,36) */
  ADB_CONVT_4X4S2 = 2, /* nn.ConvTranspose2d(k=4, s=2, p=1) as four 2x2 sub-pixel phases (medium:53,63; high:57,68) */
  ADB_CONV_K4_S2D = 3  /* 4x4 taps at offsets -2..+1, stride 1, same-size output: a 7x7 stride-2 pad-3 stem (torchvision resnet /
                          densenet conv1 / conv0 called from classifier.py:24-36) over the SPACE-TO-DEPTH image
                          [n, H/2, W/2, 16] (channel (py*2+px)*3 + c, 12 real): out(y,x) = sum W[u][v] X[2y+u-3][2x+v-3] with
                          u = 2R + py - 1.  32 bytes of operand per output pixel instead of the 320 of a full im2col. */
};

/* epilogue kinds */
enum {
  ADB_EPI_FEATURE = 0, /* y = act(acc*scale + shift (+ residual)) -> NHWC bf16 feature map (ConvBlock/ResidualBlock, base_model.py:23,36-41) */
  ADB_EPI_DOT = 1,     /* g = sigmoid(dot(act(acc*scale+shift), dot_w) + dot_b) -> fp32 [n,h,w]; cout_pad 16 or 32, dot_w fp32[cout_pad]
                          (detail_branch tail, high:84-89; transmission_branch tail, high:184-189) */
  ADB_EPI_IMAGE = 2    /* final 3-channel head fused with the output arithmetic, NCHW fp32 out (low:45; medium:117; high:135-138) */
};

/* image-epilogue arithmetic (ADB_EPI_IMAGE); v = act(acc*scale+shift) per colour channel */
enum {
  ADB_IMG_BLEND = 0,      /* out = (1-alpha)*x + alpha*v          (LightweightDehazeModel, low_intensity.py:45) */
  ADB_IMG_RESIDUAL = 1,   /* out = clamp(x + v, 0, 1)             (MediumIntensityDehazeModel, medium_intensity.py:117) */
  ADB_IMG_GUIDED = 2      /* out = clamp(x + v*guidance, 0, 1)    (HighIntensityDehazeModel, high_intensity.py:135-138) */
};

typedef struct adb_conv_desc {
  /* --- inputs: one or two NHWC bf16 feature maps read as a channel concatenation [src0 | src1]
         (torch.cat([up, skip], 1) never materialised: medium:100,111; high:117,129) */
  const void* src0; int32_t c0; int32_t c0_pitch;  /* channels used / channel pitch of the buffer */
  const void* src1; int32_t c1; int32_t c1_pitch;  /* src1 == NULL, c1 == 0 for a single source */
  int32_t n, h_in, w_in;                            /* images in this launch, input height/width */
  /* --- geometry */
  int32_t kind, kh, kw, pad;
  /* --- weights, packed by adb_pack order (see below), bf16 [groups][cout_pad][ktot]; epilogue affine fp32 [cout_pad] */
  const void* w_packed; const float* scale; const float* shift;
  int32_t cout, cout_pad;                           /* cout_pad: multiple of 16, >= 16, <= 512 */
  int32_t act;
  /* --- epilogue */
  int32_t epi;
  const void* residual; int32_t res_pitch;          /* FEATURE: optional NHWC bf16 residual, same n/h/w as dst */
  void* dst; int32_t dst_pitch; int32_t dst_c_off;  /* FEATURE: NHWC bf16 dst buffer, channel pitch, first channel written */
  const float* dot_w; float dot_b; float* dot_out;  /* DOT */
  int32_t img_mode;                                 /* IMAGE */
  const float* img_x;       /* NCHW fp32 hazy input batch the bucket was gathered from */
  float* img_out;           /* NCHW fp32 output batch (scatter target) */
  const int32_t* img_index; /* nullable: image i of this launch is batch row img_index[n_start+i] of img_x/img_out */
  const float* img_guidance;/* GUIDED: fp32 [n,h,w] of this launch's images */
  const float* img_alpha;   /* BLEND: device scalar (skip_alpha parameter, low_intensity.py:31) */
  /* --- routed-bucket dynamic batch */
  const int32_t* n_dev; int32_t n_start;
  /* --- tuning (0 = choose automatically) */
  int32_t tune_mt, tune_stages, tune_acc_stages;
  int32_t tune_flags;   /* bit 0: one TMA box per tap (no halo re-use); bit 1: descriptor base-offset experiment;
                           bit 9 (512): never take the rolling-row kernel (w_fold ignored);
                           bit 10 (1024): halve the N tile (twice the N tiles); bit 11 (2048): never split 256 channels in two */
  /* --- optional pre-activation fused into the input operand (1x1 stride-1 FEATURE convs): the conv sees
         relu(x[..., c]*pre_scale[c] + pre_shift[c]) over the c0+c1 input channels.  DenseNet's norm1/relu1 ahead of conv1
         and the transition norm/relu (torchvision densenet121 called as the north_star's HDEN) — every dense layer applies
         a different affine to the same concatenated map, so it cannot ride in the producer's epilogue. */
  const float* pre_scale; const float* pre_shift;
  /* --- optional row-folded packing of a 3x3 stride-1 filter (NULL = not provided).  When it is given and the launch is
         eligible (FEATURE epilogue, cout_pad a multiple of 32 with 3*cout_pad <= 256, w_in >= 128, filter fits shared
         memory) adb_conv2d runs the rolling-row kernel (csrc/conv_roll.cu): the three filter rows ride in the MMA's N
         dimension and accumulate into adjacent output rows held in TMEM.
         w_fold[s][(2 - r)*cout_pad + co][c] = W[co][c][r][s], bf16 [3][3*cout_pad][c0+c1]. */
  const void* w_fold;
  /* --- optional channel partials of the STORED output (FEATURE epilogue, plain / stride-2 kinds): for every group of 32 tile rows
         the epilogue also writes, per output channel, two partials of the bf16 values it stores.  stat_mode 0/1: (sum, sum of
         squares) — BatchNorm2d(train) statistics without a pass over the conv output (base_model.py:15-16 under model.train()),
         folded by adb_bn_finalize_stats.  stat_mode 2: (sum, max) — AttentionBlock's AdaptiveAvgPool2d(1) / AdaptiveMaxPool2d(1)
         of the block input (base_model.py:64-66) without a pass over it, folded by adb_attn_pool_from_stats.  Layout fp32
         [slots][2][cout_pad], slots = adb_conv2d_stat_slots(), image-major (slots / n consecutive slots per image); rows outside
         the image count nothing; every slot of a live image is written; a launch with stat_out takes conv_igemm_kernel. */
  float* stat_out; int32_t stat_mode;
} adb_conv_desc;

/* Weight packing order expected in w_packed (done on the host side by adam_dehaze_b200/engine.py):
 *   ADB_CONV_S1 / ADB_CONV_S2:  w_packed[co][(r*kw + s)*(c0+c1) + c] = W[co][c][r][s]
 *   ADB_CONVT_4X4S2:            phase g = a*2 + b (a = oh&1, b = ow&1); tap (i,j), i,j in {0,1};
 *                               r = a ? 2*i : 1 + 2*i   (input row  q + (a ? 1-i : -i));  s likewise from b, j
 *                               w_packed[g][co][(i*2 + j)*cin + ci] = Wt[ci][co][r][s]
 *   ADB_CONV_K4_S2D:            w_packed[co][(R*4 + S)*16 + (py*2 + px)*3 + c] = W[co][c][2R+py-1][2S+px-1] (0 outside the 7x7)
 *   c0 and c1 are multiples of 16; the kernel walks K in 64-channel chunks (ragged last chunk per source).
 */

/* Library / device */
ADB_API const char* adb_last_error(void);
ADB_API int adb_version(void);
ADB_API int adb_device_check(void);                 /* 0 when the current device is sm_100 and the TMA encoder resolved */

/* Implicit-GEMM convolution on tcgen05/TMEM tiles fed by TMA (bf16 x bf16 -> fp32).
 * Replaces nn.Conv2d / nn.ConvTranspose2d + BatchNorm2d(eval) + activation (+ residual add) of
 * models/dehazing/base_model.py:4-41 and the encoder/decoder/head convs of medium_intensity.py:16-76,
 * high_intensity.py:17-90, low_intensity.py:16-28. */
ADB_API int adb_conv2d(const adb_conv_desc* desc, void* stream);
/* FLOPs (2*MAC) the descriptor's launch performs for n images — the figure bench.py's roofline uses. */
ADB_API double adb_conv2d_flops(const adb_conv_desc* desc);
/* number of partial-statistics slots a launch with stat_out writes; 0 = this launch does not produce them (it takes the
 * rolling-row kernel, or is not a plain / stride-2 FEATURE conv): leave stat_out null and use adb_bn_train_stats; < 0 = bad desc */
ADB_API int64_t adb_conv2d_stat_slots(const adb_conv_desc* desc);

/* Image -> stem operand (bf16, kp channels, zero padded):
 *   out[i,ho,wo,(r*kw+s)*3+c] = x[idx(i), c, ho*sh + r - ph, wo*stride + s - pad]   (0 outside the image)
 * kh == 1: only the horizontal taps are unrolled (sh = 1, ph = 0, ho = h), which turns a kh x kw x 3 stem into a
 *          kh x 1 conv over kp channels (dehazing stems: low:16, medium:16, high:17,85);
 * kh  > 1: full im2col with the stride on both axes; the stem becomes a 1x1 conv (stride-2 HDEN stem, classifier.py:24). */
ADB_API int adb_stem_pack(const float* x, const int32_t* index, const int32_t* n_dev, int32_t n_start, int32_t n,
                  int32_t h, int32_t w, int32_t kh, int32_t kw, int32_t pad, int32_t stride, int32_t kp,
                  void* out, void* stream);

/* Layout converters (tests, classifier features): NCHW fp32 <-> NHWC bf16 */
ADB_API int adb_nchw_to_nhwc_bf16(const float* x, int32_t n, int32_t c, int32_t h, int32_t w, int32_t c_pitch, void* out, void* stream);
ADB_API int adb_nhwc_bf16_to_nchw(const void* x, int32_t n, int32_t c, int32_t h, int32_t w, int32_t c_pitch, float* out, void* stream);

/* AttentionBlock (base_model.py:43-78) as three HBM-bound passes over an NHWC bf16 map x[n,h,w,c]:
 *  1) pool:   sum_c, max_c over h*w per (image, channel)                        (avg_pool/max_pool, :64-66)
 *  2) gate:   gate = sigmoid(fc(avg)+fc(max)) (fc = w2*relu(w1*.)), then per pixel the channel mean and max of
 *             x*gate -> stats[n,h,w,2] fp32                                      (:66-73)
 *  3) apply:  spatial = sigmoid(conv7x7(stats)) (smem-tiled stencil, fp32 [n,h,w] scratch), y = x*gate*spatial (:74-78)
 * pool_buf: caller scratch of adb_pool_scratch_floats(n,h,w,c) floats; its first n*2*c floats end up holding
 * [n][2][c] (sum, max).  The reduction is deterministic (per-block partials folded in block order by the last block). */
ADB_API int64_t adb_pool_scratch_floats(int32_t n, int32_t h, int32_t w, int32_t c);
ADB_API int adb_attn_pool(const void* x, int32_t n, int32_t h, int32_t w, int32_t c, const int32_t* n_dev, int32_t n_start,
                  float* pool_buf, void* stream);
/* adb_attn_pool's result from the producing conv's stat_mode-2 partials (`slots_per_image` x [2][cpitch] per image) instead of a pass
 * over x; writes the first n*2*c floats of pool_buf (what adb_attn_gate_stats reads).  scratch: adb_attn_pool_stat_scratch_floats. */
ADB_API int64_t adb_attn_pool_stat_scratch_floats(int32_t n, int32_t slots_per_image, int32_t c);
ADB_API int adb_attn_pool_from_stats(const float* stat, int32_t n, int32_t slots_per_image, int32_t cpitch, int32_t c,
                                     const int32_t* n_dev, int32_t n_start, float* scratch, float* pool_buf, void* stream);
ADB_API int adb_attn_gate_stats(const void* x, int32_t n, int32_t h, int32_t w, int32_t c, const int32_t* n_dev, int32_t n_start,
                        const float* pool_buf, const float* w1 /*[c/r][c]*/, const float* w2 /*[c][c/r]*/, int32_t c_red,
                        float* gate /*[n][c]*/, float* stats /*[n,h,w,2]*/, void* stream);
ADB_API int adb_attn_apply(const void* x, int32_t n, int32_t h, int32_t w, int32_t c, const int32_t* n_dev, int32_t n_start,
                   const float* gate, const float* stats, const float* w_spatial /*[2][7][7]*/, float* spatial /*[n,h,w]*/,
                   void* y, void* stream);

/* Pooling for the HDEN backbones (torchvision resnet/densenet called from models/classifier.py:24-36,91). */
ADB_API int adb_maxpool3x3s2(const void* x, int32_t n, int32_t h, int32_t w, int32_t c, void* y, int32_t pitch_out, void* stream);
ADB_API int adb_global_avgpool(const void* x, int32_t n, int32_t h, int32_t w, int32_t c, float* scratch /*adb_pool_scratch_floats*/,
                       float* y /*[n][c] fp32*/, void* stream);
/* DenseNet121 HDEN arm (north_star; torchvision densenet121 — the reference has no DenseNet):
 * pre-activation y = relu(x*scale + shift) on the first c channels of an NHWC bf16 map (norm1/relu1 ahead of conv1),
 * and the 2x2/2 average pool of the transitions. */
ADB_API int adb_affine_relu(const void* x, int64_t pixels, int32_t c, int32_t pitch_in, const float* scale, const float* shift,
                    void* y, int32_t pitch_out, void* stream);
ADB_API int adb_avgpool2x2(const void* x, int32_t n, int32_t h, int32_t w, int32_t c, int32_t pitch_in, void* y,
                   int32_t pitch_out, void* stream);
/* Non-default branch variants: nn.MaxPool2d(k, k) for k in {2, 4} (medium_intensity.py:144,149;
 * high_intensity.py:164,167) and nn.UpsamplingBilinear2d(scale_factor) == bilinear with align_corners=True
 * (medium_intensity.py:146,151; high_intensity.py:171,173) on NHWC bf16 maps.  The upsampler writes channels
 * [c_off, c_off+c) of a wider buffer so COrunInspiredModel's three-scale concat is never materialised. */
ADB_API int adb_maxpool_kxk(const void* x, int32_t n, int32_t h, int32_t w, int32_t c, int32_t k, const int32_t* n_dev,
                            int32_t n_start, void* y, void* stream);
ADB_API int adb_upsample_bilinear(const void* x, int32_t n, int32_t h, int32_t w, int32_t c, int32_t scale,
                                  const int32_t* n_dev, int32_t n_start, void* y, int32_t pitch_out, int32_t c_off, void* stream);
/* Classifier head, fp32: logits = W2*relu(W1*f + b1) + b2 (models/classifier.py:72-78, eval mode: dropout = identity). */
ADB_API int adb_head_mlp(const float* feat, int32_t n, int32_t f, const float* w1, const float* b1, int32_t hidden,
                 const float* w2, const float* b2, int32_t classes, float* logits, void* stream);

/* y[n][fout] = act(W x + b), fp32 (first layer of the GatedRouter gate MLP, routing.py:155-163). */
ADB_API int adb_linear(const float* x, int32_t n, int32_t fin, const float* w, const float* b /*nullable*/, int32_t fout,
               int32_t relu, float* y, void* stream);

/* Routing (models/routing.py:40-61): intensity = argmax(logits,1) (first max wins, NaN counts as max, like
 * torch.argmax), masks, and a stable 3-way compaction: bucket k lists the ascending batch rows with intensity == k.
 * intensity_in (nullable) overrides the argmax (HardRouter.forward(x, intensity=...), routing.py:23,40).
 * Outputs: intensity int64[b]; masks uint8[3][b]; bucket_index int32[3][b]; bucket_count int32[3]. */
ADB_API int adb_route(const float* logits, const int64_t* intensity_in, int32_t b, int32_t classes,
              int64_t* intensity, uint8_t* masks, int32_t* bucket_index, int32_t* bucket_count, void* stream);

/* HardRouter's output buffer (routing.py:31 `torch.zeros_like(x)`): rows whose class id is outside {0,1,2} are written by
 * no branch and must read as zeros; every other row is overwritten in full by its branch.  Clears only the former
 * (out: NCHW fp32 [b][row_elems], intensity int64[b] as written by adb_route) instead of memset-ing the whole batch. */
ADB_API int adb_zero_unrouted(float* out, const int64_t* intensity, int32_t b, int64_t row_elems, void* stream);

/* Soft/gated blend (routing.py:111-127, 215-221): w = softmax(logits/T) (or given weights when temperature <= 0),
 * out = sum_k w[:,k] * y_k, NCHW fp32. */
ADB_API int adb_blend3(const float* y0, const float* y1, const float* y2, const float* logits_or_weights, float temperature,
               int32_t b, int64_t chw, float* weights_out /*[b][3]*/, float* out, void* stream);

/* Loss reductions (training/loss.py:121,81,177) forward + backward w.r.t. pred/logits.
 * l1: mean|p-t|, mse: mean (p-t)^2 — one pass, warp-shuffle + one atomic per block; grad_scale multiplies dL/dp. */
ADB_API int adb_l1_mse_fwd(const float* pred, const float* target, int64_t numel, float* out2 /*[l1, mse]*/, void* stream);
ADB_API int adb_l1_bwd(const float* pred, const float* target, int64_t numel, float grad_scale, float* grad, void* stream);
ADB_API int adb_mse_bwd(const float* pred, const float* target, int64_t numel, float grad_scale, float* grad, void* stream);
ADB_API int adb_ce_fwd_bwd(const float* logits, const int64_t* labels, int32_t b, int32_t classes, float grad_scale,
                   float* loss /*[1]*/, float* grad_logits /*nullable [b][classes]*/, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Training step (training/train_dehazing.py:71-96, training/train_joint.py:129-154: forward in train() mode,
 * loss.backward(), optimizer.step()).  Data gradients (dgrad) re-use adb_conv2d with transformed weights
 * (adam_dehaze_b200/training/autograd.py); the entry points below are the rest of backward.
 * ------------------------------------------------------------------------------------------------------------------ */

/* Weight gradient of one convolution as a pixel-reduction implicit GEMM on tcgen05 (MN-major operands straight from the
 * NHWC TMA boxes), fp32 accumulate:
 *     dw[m][c][r][s] (+)= sum_{n,y,x} small[n,y,x,m] * large[n, y*stride + r - pad, x*stride + s - pad, c]
 *  - nn.Conv2d:             small = dL/d(conv output) (cg channels), large = the conv input ([act0 | act1] concat);
 *                           dw is conv.weight.grad, [cout][cin][kh][kw]
 *  - nn.ConvTranspose2d(4,2,1): small = the layer INPUT, large = dL/d(output), kind = ADB_CONV_S2, kh = kw = 4, pad = 1;
 *                           dw is weight.grad, [cin][cout][4][4]
 * layout ADB_WG_STEM: `act0` is an adb_stem_pack operand (kh x 1 conv over c0 = kp channels holding stem_kw x 3 values);
 *                           dw is the 3-channel stem's weight.grad [cout][3][kh][stem_kw].
 * workspace: adb_wgrad_workspace_bytes(desc) bytes of fp32 scratch (per-split partial slabs, folded in a fixed order:
 * the result is deterministic).  A smaller workspace (>= one slab) is accepted and lowers the split count. */
enum { ADB_WG_OIHW = 0, ADB_WG_STEM = 1 };
typedef struct adb_wgrad_desc {
  const void* grad; int32_t cg; int32_t cg_pitch; int32_t cg_true;   /* small map: channels (multiple of 16) / pitch / rows of dw written (0 = cg) */
  const void* act0; int32_t c0; int32_t c0_pitch;                    /* large map source(s), NHWC bf16 */
  const void* act1; int32_t c1; int32_t c1_pitch;
  int32_t n, h_in, w_in;                                             /* of the large map */
  int32_t kind, kh, kw, pad;                                         /* ADB_CONV_S1 ('same') or ADB_CONV_S2 (halving) */
  float* workspace; int64_t workspace_bytes;
  float* dw; int32_t layout; int32_t stem_kw; int32_t accumulate;    /* accumulate != 0: dw += result */
  int32_t mode;   /* 0 = choose (tap-packed kernel when cg <= 64), 1 = channel-major kernel, 2 = tap-packed kernel */
} adb_wgrad_desc;
ADB_API int64_t adb_wgrad_workspace_bytes(const adb_wgrad_desc* desc);
ADB_API int adb_wgrad(const adb_wgrad_desc* desc, void* stream);
ADB_API double adb_wgrad_flops(const adb_wgrad_desc* desc);

/* BatchNorm2d in train() mode (base_model.py:15-16 under model.train(), train_dehazing.py:66) on a raw conv output z
 * (NHWC bf16, `pixels` = n*h*w rows of `c` channels, channel pitch `pitch`):
 *   mean/var over pixels (biased variance normalises, the unbiased one updates running_var with `momentum`),
 *   mean[c], rstd[c] are kept for backward, scale = gamma*rstd and shift = beta - mean*scale feed adb_affine_act.
 * scratch: adb_bn_scratch_floats(pixels, c) floats.  running_* / num_batches_tracked are nullable. */
ADB_API int64_t adb_bn_scratch_floats(int64_t pixels, int32_t c);
/* The same result as adb_bn_train_stats from the conv epilogue's partials (adb_conv_desc.stat_out, `slots` x [2][cpitch]) instead of
 * a pass over z.  scratch: adb_bn_stat_scratch_floats(slots, c) floats. */
ADB_API int64_t adb_bn_stat_scratch_floats(int64_t slots, int32_t c);
ADB_API int adb_bn_finalize_stats(const float* stat, int64_t slots, int32_t cpitch, int64_t pixels, int32_t c, const float* gamma,
                                  const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                                  int64_t* num_batches_tracked, float* scratch, float* mean, float* rstd, float* scale,
                                  float* shift, void* stream);
ADB_API int adb_bn_train_stats(const void* z, int64_t pixels, int32_t c, int32_t pitch, const float* gamma, const float* beta,
                               float eps, float momentum, float* running_mean, float* running_var,
                               int64_t* num_batches_tracked, float* scratch, float* mean, float* rstd, float* scale,
                               float* shift, void* stream);
/* y = act(z*scale + shift (+ residual)) — the normalise/activate pass of ConvBlock / ResidualBlock in train() mode. */
ADB_API int adb_affine_act(const void* z, int32_t pitch_z, int64_t pixels, int32_t c, const float* scale, const float* shift,
                           const void* residual, int32_t pitch_r, int32_t act, void* y, int32_t pitch_y, void* stream);
/* Backward of y = act(BN(z) (+ residual)):  g = dy * act'(y) is written to g_out (it is also the gradient of the
 * residual input; g_out may alias dy), dgamma = sum g*xhat, dbeta = sum g, dz = gamma*rstd*(g - mean(g) - xhat*mean(g*xhat)).
 * gamma == NULL (and z == NULL): bias-only layer — g_out is dz and dbeta is the conv bias gradient.  y == NULL: no activation. */
ADB_API int adb_bn_bwd(const void* dy, int32_t pitch_dy, const void* y, int32_t pitch_y, const void* z, int32_t pitch_z,
                       int64_t pixels, int32_t c, int32_t act, const float* gamma, const float* mean, const float* rstd,
                       float* scratch, void* g_out, int32_t pitch_g, void* dz, int32_t pitch_dz, float* dgamma, float* dbeta,
                       int32_t accumulate, void* stream);
/* The same backward for y = relu(BN(z)) with no residual input, without touching y or materialising g: the ReLU mask is
 * recomputed as fmaf(z, scale, shift) > 0 from the (scale, shift) adb_bn_train_stats returned for the forward (the exact
 * expression adb_affine_act evaluated), so the pass structure is read(dy, z) + read(dy, z)/write(dz) instead of
 * read(dy, y, z)/write(g) + read(g, z)/write(dz).  dz may alias dy; dz_accumulate != 0: dz += result (the gradient prefix
 * of a DenseNet block buffer, replacing a separate adb_add_bf16 pass).  base_model.py:15-24, torchvision densenet norm/relu. */
ADB_API int adb_bn_relu_bwd(const void* dy, int32_t pitch_dy, const void* z, int32_t pitch_z, int64_t pixels, int32_t c,
                            const float* scale, const float* shift, const float* gamma, const float* mean, const float* rstd,
                            float* scratch, void* dz, int32_t pitch_dz, int32_t dz_accumulate, float* dgamma, float* dbeta,
                            int32_t accumulate, void* stream);
/* a[..., :c] += b[..., :c] (gradient accumulation at a fan-out: skip connections, concat sources). */
ADB_API int adb_add_bf16(void* a, int32_t pitch_a, const void* b, int32_t pitch_b, int64_t pixels, int32_t c, void* stream);

/* The image heads un-fused for training: z = the 3-channel head conv's raw output (NHWC bf16, first 3 channels of `pitch`);
 * forward = the ADB_EPI_IMAGE arithmetic (img_mode, act); backward turns dL/d(out) (NCHW fp32) into dz (bf16, padding
 * channels zeroed), dguidance (fp32 [n,h,w], GUIDED) and red4 = {dbias[0..2], dalpha} (low:45, medium:117, high:135-138). */
ADB_API int adb_img_head_fwd(const void* z, int32_t pitch, const float* x, const float* guidance, const float* alpha,
                             int32_t mode, int32_t act, int32_t n, int32_t h, int32_t w, float* out, void* stream);
ADB_API int adb_img_head_bwd(const float* dout, const void* z, int32_t pitch, const float* x, const float* guidance,
                             const float* alpha, int32_t mode, int32_t act, int32_t n, int32_t h, int32_t w, void* dz,
                             float* dguidance, float* red4, void* stream);
/* 1x1 conv to one channel + sigmoid (detail_branch tail, high_intensity.py:87-89) forward / backward;
 * red = {dw[0..c), dbias}. */
ADB_API int adb_dot_head_fwd(const void* y, int32_t pitch, int32_t c, const float* w, const float* b, int64_t pixels, float* g,
                             void* stream);
ADB_API int adb_dot_head_bwd(const float* dg, const float* g, const void* y, int32_t pitch, int32_t c, const float* w,
                             int64_t pixels, void* dy, int32_t pitch_dy, float* red, void* stream);

/* AttentionBlock backward (base_model.py:64-78) from the tensors its forward kept (pool_buf [n][2][c], gate [n][c],
 * stats [n,h,w,2], spatial [n,h,w]): dx (NHWC bf16) and the gradients of fc.0 / fc.2 / conv_spatial (fp32, overwritten).
 * scratch: adb_attn_bwd_scratch_floats(n,h,w,c) floats. */
ADB_API int64_t adb_attn_bwd_scratch_floats(int32_t n, int32_t h, int32_t w, int32_t c);
ADB_API int adb_attn_bwd(const void* dy, const void* x, int32_t n, int32_t h, int32_t w, int32_t c, const float* pool,
                         const float* gate, const float* stats, const float* spatial, const float* w1, const float* w2,
                         int32_t c_red, const float* w_spatial, float* scratch, void* dx, float* dw1, float* dw2,
                         float* dw_spatial, void* stream);

/* Backward of adb_blend3 (SoftRouter / GatedRouter under train_joint.py:141-150): dy_k = w[:,k] * dout,
 * dweights[b][k] = sum dout * y_k, and (dlogits non-NULL) the softmax(logits/T) backward. */
ADB_API int adb_blend3_bwd(const float* dout, const float* y0, const float* y1, const float* y2, const float* weights,
                           float temperature, int32_t b, int64_t chw, float* dy0, float* dy1, float* dy2, float* dweights,
                           float* dlogits, void* stream);

/* Perceptual loss terms (training/loss.py:47-108: VGG16 content loss, LPIPS-alex), forward and d/d(pred).  The frozen
 * feature networks run on adb_conv2d (forward: bias + ReLU fused; backward: the data-gradient form); these are the
 * pieces around them.  Host pointers are marked _host. */
ADB_API int adb_image_affine(const float* x, int32_t n, int32_t h, int32_t w, const float* scale3_host, const float* shift3_host,
                             float* y, void* stream);           /* y = x*scale[c] + shift[c], NCHW fp32 (loss.py:63-67,104-105) */
ADB_API int adb_maxpool_fwd(const void* x, int32_t n, int32_t h, int32_t w, int32_t c, int32_t k, int32_t stride, int32_t pad,
                            void* y, void* stream);             /* nn.MaxPool2d(k, stride, pad), NHWC bf16, k in {2,3,4} */
ADB_API int adb_maxpool_bwd(const void* dy, const void* x, const void* y, int32_t n, int32_t h, int32_t w, int32_t c, int32_t k,
                            int32_t stride, int32_t pad, void* dx, void* stream);
/* out[0] += mean((a-b)^2) (F.mse_loss, loss.py:81); da (nullable, bf16) = grad_scale * 2(a-b)/numel */
ADB_API int adb_mse_feat(const void* a, const void* b, int64_t numel, float grad_scale, float* out, void* da, void* stream);
/* LPIPS tap: channel-unit-normalise both maps, lin-weighted squared difference, spatial mean: out[n] += value;
 * dfa (nullable, bf16) = grad_scale * d value / d fa */
ADB_API int adb_lpips_tap(const void* fa, const void* fb, int32_t n, int32_t h, int32_t w, int32_t c, const float* lin_w,
                          float grad_scale, float* out, void* dfa, void* stream);
/* Transpose of adb_stem_pack (gradient of a 3-channel stem w.r.t. the image): dx (NCHW fp32) (+)= scale[c] * col2im(dcols) */
ADB_API int adb_stem_unpack(const void* dcols, int32_t n, int32_t h, int32_t w, int32_t kh, int32_t kw, int32_t pad,
                            int32_t stride, int32_t kp, const float* scale3_host, int32_t accumulate, float* dx, void* stream);

/* HDEN in train() mode (train_joint.py:129-150 trains the classifier through the router): backward of the global
 * average pool, element-wise products for the head's dropout masks / ReLU gate, and nn.Linear backward (classifier.py:72-78). */
ADB_API int adb_broadcast_hw(const float* dfeat, int32_t n, int32_t h, int32_t w, int32_t c, float scale, void* dx, void* stream);
ADB_API int adb_mul_f32(const float* a, const float* b /*nullable*/, const float* gate /*nullable: 0 where gate <= 0*/, int64_t n,
                        float* out, void* stream);
ADB_API int adb_linear_bwd(const float* x, const float* w, const float* dy, int32_t n, int32_t fin, int32_t fout, float* dx,
                           float* dw, float* db, void* stream);

/* backward of adb_avgpool2x2 (DenseNet transition pool): dx[n, h, w, :c] = dy[n, h/2, w/2, :c] / 4; h, w of dx */
ADB_API int adb_avgpool2x2_bwd(const void* dy, int32_t pitch_dy, int32_t n, int32_t h, int32_t w, int32_t c, void* dx,
                               int32_t pitch_dx, void* stream);

/* Weight re-packing after an optimizer step: out[i] = bf16(src[idx[i]]), 0 where idx[i] < 0 (idx = the packing's
 * permutation of the fp32 parameter, derived once on the host side). */
ADB_API int adb_gather_cast(const float* src, const int32_t* idx, int64_t n, void* out, void* stream);
/* Many re-packs in one launch.  jobs_dev: device array of njobs records {const float* src; const int32_t* idx; void* out;
 * int64_t n; int64_t first_block} (40 bytes each); job j owns blocks [first_block_j, first_block_j + ceil(n_j / 2048)) of the
 * total_blocks-block grid. */
ADB_API int adb_gather_cast_multi(const void* jobs_dev, int32_t njobs, int64_t total_blocks, void* stream);

/* backward of adb_upsample_bilinear (align_corners=True): dx[n,h,w,c] from dy[n, h*scale, w*scale, c_off : c_off+c] */
ADB_API int adb_upsample_bilinear_bwd(const void* dy, int32_t pitch_dy, int32_t c_off, int32_t n, int32_t h, int32_t w,
                                      int32_t c, int32_t scale, void* dx, void* stream);

/* One Adam step on a flat fp32 tensor with torch.optim.Adam semantics (train_dehazing.py:33-37: weight_decay is L2
 * added to the gradient); grad_scale multiplies the gradient first (1/world_size after a sum all-reduce). */
ADB_API int adb_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t numel, float lr,
                          float beta1, float beta2, float eps, float weight_decay, int32_t step, float grad_scale, void* stream);

/* The same step with torch.optim.Adam's per-parameter skipping: the flat tensor is cut into `n_seg` segments
 * (seg_offsets int64[n_seg], ascending, multiples of 4, seg_offsets[0] == 0); a segment with seg_live[s] <= 0 (no rank
 * produced a gradient for that parameter: p.grad is None everywhere) is left untouched — no moment decay, no weight
 * decay, no step count (torch/optim/adam.py skips such parameters; train_joint.py:149-150 under HardRouter).  seg_step
 * (int32[n_seg], device) is the per-parameter step count, advanced here; seg_bc1 / seg_bc2s are fp32[n_seg] scratch. */
ADB_API int adb_adam_step_segments(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t numel, float lr,
                                   float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                                   const int64_t* seg_offsets, int32_t n_seg, const float* seg_live, int32_t* seg_step,
                                   float* seg_bc1, float* seg_bc2s, void* stream);

/* Image-quality metrics on the device (evaluation/metrics.py:13-36): per image PSNR
 * (data_range 1) and SSIM with skimage's defaults on the channel-mean grayscale (7x7 uniform window, sample
 * covariance, K1 = 0.01, K2 = 0.03, mean over the interior).  pred/target: NCHW fp32 [n,3,h,w]; scratch: 2n doubles. */
ADB_API int adb_image_metrics(const float* pred, const float* target, int32_t n, int32_t h, int32_t w, double* scratch,
                              float* psnr, float* ssim, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Route guard (csrc/guard_fp32.cu): fp32 re-evaluation of HDEN for the batch rows whose bf16 top-2 logit gap is below
 * eps, so that HardRouter's argmax (routing.py:41-43) equals the fp32 reference's even for near-ties.  No host round
 * trip: the rows are listed on the device (flag_index / flag_count), every kernel of a pass works on list positions
 * [*cursor, *cursor + cap) and exits when they are not live, and adb_guard_graph_* wraps one pass in a CUDA-graph WHILE
 * node that repeats it while rows remain.  Maps are NHWC fp32; arithmetic is fp32 FMA on the CUDA cores in the
 * reference's op order (conv -> BatchNorm affine -> ReLU / residual / pool), classifier.py:80-97. */
ADB_API int adb_guard_flags(const float* logits, int32_t b, int32_t classes, float eps, int32_t* flag_index /*[b]*/,
                            int32_t* flag_count, int32_t* cursor /*reset to 0*/, void* stream);
/* slots[0] = the NCHW fp32 image batch, slots[1] = the fp32 logits [b][classes] to patch: read through the slots by the
 * stem conv and the scatter, so a captured graph serves every batch of the same shape */
ADB_API int adb_guard_set_slots(const void** slots, const void* images, void* logits, void* stream);
typedef struct adb_f32_conv_desc {
  const int32_t* flag_index; const int32_t* flag_count; const int32_t* cursor; int32_t cap;   /* the live rows of this pass */
  const float* x;               /* NHWC fp32 [cap][h_in][w_in][in_pitch] (in_nchw == 0) */
  const float* const* x_slot;   /* in_nchw != 0: *x_slot = NCHW fp32 [B][cin][h_in][w_in]; row = flag_index[*cursor + i] */
  int32_t in_nchw;
  int32_t h_in, w_in, cin, in_pitch;
  int32_t kh, kw, stride, pad;
  const float* w; int32_t cout; /* fp32 [kh*kw][cin][cout] */
  const float* pre_scale; const float* pre_shift;    /* nullable: input seen as relu(x*pre_scale[c] + pre_shift[c]), zero padding after it */
  const float* post_scale; const float* post_shift;  /* nullable: y = acc*post_scale[co] + post_shift[co] */
  int32_t post_relu;
  const float* residual; int32_t res_pitch;          /* nullable NHWC fp32, added before the ReLU */
  float* y; int32_t out_pitch; int32_t out_c_off;    /* NHWC fp32 [cap][h_out][w_out][out_pitch], first channel written */
} adb_f32_conv_desc;
ADB_API int adb_f32_conv2d(const adb_f32_conv_desc* desc, void* stream);
/* mode 0: 3x3 stride-2 pad-1 max pool; mode 1: 2x2 stride-2 average pool */
ADB_API int adb_f32_pool(const int32_t* flag_index, const int32_t* flag_count, const int32_t* cursor, int32_t cap, const float* x,
                         int32_t h, int32_t w, int32_t c, int32_t in_pitch, int32_t mode, float* y, int32_t out_pitch, void* stream);
/* feats[i][c] = mean_hw relu(x*scale + shift) (scale/shift nullable = plain mean) */
ADB_API int adb_f32_global_avgpool(const int32_t* flag_index, const int32_t* flag_count, const int32_t* cursor, int32_t cap,
                                   const float* x, int32_t hw, int32_t c, int32_t pitch, const float* scale, const float* shift,
                                   float* feats, void* stream);
/* (*logits_slot)[flag_index[*cursor + i]][:] = logits_c[i][:] for the live rows */
ADB_API int adb_guard_scatter(const int32_t* flag_index, const int32_t* flag_count, const int32_t* cursor, int32_t cap,
                              const float* logits_c, int32_t classes, float* const* logits_slot, void* stream);
ADB_API int adb_guard_advance(const int32_t* flag_count, int32_t* cursor, int32_t cap, void* stream);   /* *cursor += cap */
/* begin: starts capturing `stream` into the body of a WHILE node; the caller then issues one pass on that stream;
 * end: appends "cursor += cap; repeat while cursor < count", instantiates; launch: cursor = 0, run while rows remain. */
ADB_API int adb_guard_graph_begin(const int32_t* flag_count, int32_t* cursor, int32_t cap, void* stream, void** ctx_out);
ADB_API int adb_guard_graph_end(void* ctx);
ADB_API int adb_guard_graph_launch(void* ctx, void* stream);
ADB_API int adb_guard_graph_destroy(void* ctx);

/* Input transform of the reference loader (data/dataset.py:73-99) on the device: decoded uint8 HWC images (BGR as cv2.imread
 * returns them when bgr != 0) -> NCHW fp32 in [0,1]: channel swap (cv2.cvtColor BGR2RGB), cv2.resize(INTER_LINEAR) for 8-bit
 * images reproduced bit for bit when (hs, ws) != (hd, wd), transforms.ToTensor() (/255), and the training split's horizontal
 * (bit 0) / vertical (bit 1) flips per image (flips nullable).  src: n images, src_image_stride bytes apart. */
ADB_API int adb_image_u8_to_f32(const uint8_t* src, int32_t n, int32_t hs, int32_t ws, int64_t src_image_stride, int32_t bgr,
                                const uint8_t* flips, float* dst, int32_t hd, int32_t wd, void* stream);

/* Developer aid: copy the clock64() timeline CTA 0 recorded during the last adb_conv2d launched with tune_flags bit 2
 * ([6 roles][256 events]: A producer, B producer, MMA ready, MMA issued, epilogue start, epilogue end). Synchronises. */
ADB_API int adb_debug_timeline(int64_t* host_out, int32_t count);

/* Read and clear the device-side kernel error flag (non-zero => a bounded mbarrier wait expired). Synchronises. */
ADB_API int adb_kernel_error_flag(void);

#ifdef __cplusplus
}
#endif
#endif /* ADB200_H */
