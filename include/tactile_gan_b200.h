/*
 * libtactile_gan_b200.so -- C-ABI of the B200 (sm_100a) kernels behind the tactile-gan G+D step.
 *
 * The reference (mmheydari97/tactile-gan) has no FFI: its hot path is a chain of ATen/cuDNN library
 * calls issued by PyTorch eager.  Each entry point below replaces the library calls made at the
 * cited reference lines; the Python host (tactile_gan_b200/*.py) binds them with ctypes.
 *
 * Conventions: plain pointers and sizes only; device pointers unless stated; activations are NHWC
 * bf16 with the channel count padded to a multiple of 64; `stream` is a cudaStream_t passed as
 * void*; every call returns 0 on success or a negative code (tg_last_error() gives the text),
 * never throws, never allocates device memory and never synchronises.
 */
#ifndef TACTILE_GAN_B200_H
#define TACTILE_GAN_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define TG_MAX_SRC 6
#define TG_MAX_TAPS 16

enum { TG_ACT_NONE = 0, TG_ACT_LRELU = 1, TG_ACT_SIGMOID = 2, TG_ACT_RELU = 3, TG_ACT_TANH = 4 };
enum { TG_GAN_LS = 0, TG_GAN_CE = 1, TG_GAN_W = 2, TG_GAN_HINGE = 3 };

int tg_version(void);
const char* tg_last_error(void);
int tg_device_sm_count(void);

/* NHWC bf16 tensor view: c contiguous, strides in elements. */
typedef struct tg_view {
  void* ptr;
  int n, h, w, c;
  long long sn, sh, sw;
} tg_view;

/* One operand of the (virtual-concat) K loop: activation view + its slice of the packed weights
 * bf16 [wgt_taps][wgt_rows][wgt_k]; this source multiplies columns [k_off, k_off + act.c). */
typedef struct tg_conv_src {
  tg_view act;
  const void* wgt;
  int wgt_taps, wgt_rows, wgt_k, k_off, row_off;
} tg_conv_src;

/* Implicit-GEMM convolution (tcgen05 / TMEM / TMA).
 *   out[n,y,x,co] = act( bias[co] + sum_src sum_tap sum_ci act_src[n, y*stride+dy[tap], x*stride+dx[tap], ci]
 *                                                         * wgt_src[tap_w[tap]][row_off+co][k_off+ci] )
 * Replaces nn.Conv2d / nn.ConvTranspose2d forward and their input-gradient (dgrad):
 *   generators/UNet_plusplus.py:22,26  generators/UNet.py:21,25,40,44  generators/BCDUNet.py:122-137
 *   discriminators/PatchDiscriminator.py:14,27 and torch.cat / nn.Upsample operands at
 *   generators/UNet_plusplus.py:72-84 (the concat is the multi-source K loop, never materialised).
 * stats_partial (optional): [n][stats_tiles_total][c][2] fp32 (sum, sum of squares) per output tile for
 * the InstanceNorm that follows (nn.InstanceNorm2d, UNet_plusplus.py:23,27). */
typedef struct tg_conv_desc {
  int num_src;
  tg_conv_src src[TG_MAX_SRC];
  tg_view out;
  int taps, stride;
  signed char tap_dy[TG_MAX_TAPS], tap_dx[TG_MAX_TAPS], tap_w[TG_MAX_TAPS];
  const float* bias;      /* fp32 [bias_len]; channels >= bias_len get no bias */
  int bias_len;
  float* stats_partial;
  int stats_tiles_total;  /* tiles per image in the partial buffer (0: this launch's tiles_per_img) */
  int stats_tile_off;     /* first tile slot this launch writes (phase-decomposed outputs share a buffer) */
  int act;
  float slope;
  int pool_out;           /* 1: `out` is the half-resolution tensor and the epilogue stores the 2x2 SUM of the
                           * (2*out.h x 2*out.w) result -- the input gradient of a conv that read a nearest-
                           * upsampled tensor (nn.Upsample, UNet_plusplus.py:40), folded back without materialising it */
  float* splitk_ws;       /* optional fp32 workspace (caller-owned, shared by the launches of one stream): convolutions
                           * with too few output tiles to fill the GPU (the 2x2 .. 8x8 levels of UNet, any tiny batch)
                           * split their K loop over several CTAs, each writing its partial tile to its own slice, and
                           * a second kernel sums the slices in a fixed order (deterministic) and applies bias /
                           * activation. NULL: never split. */
  long long splitk_ws_bytes;
} tg_conv_desc;

/* Weight gradient (split-K implicit GEMM over pixels, fp32 accumulation into dw with red.add):
 *   dw[tap_w[tap]][qc][pc] += sum_{n,y,x} P[n, y*stride+dy, x*stride+dx, pc] * Q[n,y,x,qc]
 * Replaces the autograd wgrad of the convolutions above (loss.backward(), train.py:134,167). */
typedef struct tg_wgrad_desc {
  int num_src;
  tg_view p[TG_MAX_SRC];
  tg_view q;
  int taps, stride;
  signed char tap_dy[TG_MAX_TAPS], tap_dx[TG_MAX_TAPS], tap_w[TG_MAX_TAPS];
  float* dw;
  int dw_rows, dw_cols;
} tg_wgrad_desc;

typedef struct tg_plan tg_plan;
int tg_conv_query_tiles(int n, int ho, int wo, int want_stats, int* out4 /* th, tw, tn, tiles_per_img */);
int tg_conv_plan_create(const tg_conv_desc* desc, tg_plan** plan);
int tg_wgrad_plan_create(const tg_wgrad_desc* desc, tg_plan** plan);
int tg_plan_run(tg_plan* plan, void* stream);
void tg_plan_destroy(tg_plan* plan);
/* device int the kernels set (non-zero) when a pipeline wait times out; host-readable after sync */
int* tg_error_flag_device_ptr(void);
/* Programmatic dependent launch of the GEMM / InstanceNorm kernels (each waits with griddepcontrol.wait after its
 * set-up): 0 off, 1 eager launches, 2 also inside stream capture (programmatic graph edges). Returns the previous value;
 * a negative argument only queries. Default 0 (env TG_PDL): neutral on the training step, +4..13 % on batch-1 inference,
 * where the engines switch it on themselves (engine.GraphEngine.forward_graphed). */
int tg_pdl_policy(int policy);
int tg_error_flag_read(void); /* synchronising D2H read of that flag (diagnostics / tests only) */

/* ---- layout packs (train.py:101 `.to(device)` + torch.cat at PatchDiscriminator.py:36, and the
 *      GP interpolation alpha*real + (1-alpha)*fake at util.py:79-83) */
int tg_pack_nchw(const float* A, const float* B, const float* wa, const float* wb, void* out, int N,
                 int HW, int cj, int C, int c_off, void* stream);
int tg_unpack_nhwc(const void* in, float* out, int N, int HW, int C, int c_off, int cj, float scale,
                   void* stream);

/* ---- thin-channel discriminator layers restated as dense 1x1 GEMMs (PatchDiscriminator.py:14,27,36):
 *      tg_im2col_pack writes cat(A, wa*B + wb*B2) as im2col rows (k*k*(ca+cb) <= 64 channels, no padding) so the
 *      first conv is a 1-tap GEMM; tg_col2im_grad folds its input gradient back to an fp32 NCHW image;
 *      tg_head_gather sums the 9 tap channels of the head's 1x1 GEMM (+bias, sigmoid); tg_head_scatter is
 *      its transpose; tg_gp_normsq_img is the penalty norm (util.py:92) on the image-space gradient */
int tg_im2col_pack(const float* A, const float* B, const float* B2, const float* wa, const float* wb, void* cols,
                   int N, int ca, int cb, int H, int W, int k, int stride, void* stream);
int tg_col2im_grad(const void* dcols, float* g, int N, int ct, int c_off, int cj, int H, int W, int k, int stride,
                   float scale, void* stream);
int tg_head_gather(const void* hc, const float* bias, void* out, int N, int Hi, int Wi, int k, int C, int act,
                   float slope, void* stream);
int tg_head_scatter(const void* dz, void* dzc, int N, int Hi, int Wi, int k, int C, void* stream);
int tg_gp_normsq_img(const float* g, int N, long long per_img, float* nsq, void* stream);

/* ---- InstanceNorm2d(eps, biased var) (+affine) fused with ReLU / LeakyReLU, optional AvgPool2d(2) /
 *      MaxPool2d(2) copy and nearest Upsample(x2) copy (UNet_plusplus.py:23-24,40-41; BCDUNet.py:110;
 *      PatchDiscriminator.py:16-17) and their backward */
/* The passes whose inputs all sit at the unit's own resolution have two forms: register-staged loads and a
 * cp.async.bulk (1-D TMA) shared-memory ring with a persistent grid. policy 0: never the ring, 1 (default): the faster
 * form per shape as measured (profiles/r02_tail_microbench_*.txt), 2: the ring whenever the shape allows (tests).
 * Returns the previous policy; a negative argument only queries. Env TG_STREAM=0/1/2 sets the initial value. */
int tg_in_stream_policy(int policy);
/* Slim form of the ring's backward passes (4 KiB chunks, one CTA per SM, <= TG_SLIM_KB = 32 KiB of shared memory):
 * it fits on an SM beside a persistent weight-gradient GEMM CTA, so the engine switches it on (1) for the backward
 * passes it issues while a weight gradient runs on the side stream and off (0) otherwise. Returns the previous
 * value; any other argument only queries. */
int tg_in_stream_slim(int on);
/* Serpentine order: bit 0 / 1 / 2 makes the forward / statistics / apply pass walk its tensor from the end, so that it
 * starts on the part its predecessor (which walked the other way) left in L2. Results do not depend on it. Returns
 * the previous mask; an argument outside 0..7 only queries. Env TG_SERP sets the initial value (default 0: measured neutral). */
int tg_in_stream_serpentine(int mask);
int tg_in_finalize(const float* partial, float* mr, int N, int T, int C, int count, float eps, void* stream);
int tg_in_stats_direct(const void* raw, float* mr, int N, int HW, int C, float eps, void* stream);
int tg_in_act_fwd(const void* raw, const float* mr, const float* gamma, const float* beta, void* y,
                  void* pool, int pool_mode, void* up, int N, int H, int W, int C, int c_valid, int act,
                  float slope, void* stream);
/* g_up: gradient arriving through the 2x nearest-upsampled copy -- at the upsampled resolution (gathered 2x2 here)
 * or, with g_up_pooled, already summed to this tensor's resolution by a pool_out conv */
int tg_in_bwd_reduce(const void* raw, const void* y, const float* mr, const float* gamma,
                     const float* beta, const void* g_same, const void* g_pool, int pool_mode,
                     const void* g_up, int g_up_pooled, void* dn, float* red, int N, int H, int W, int C,
                     int c_valid, int act, float slope, void* stream);
/* second pass without a stored dn: re-reads raw and the gradient routes of tg_in_bwd_reduce (dn may be NULL there),
 * recomputes dn in fp32 and writes dz -- 5 instead of 6 tensor-sizes of HBM traffic per unit */
int tg_in_bwd_apply_re(const void* raw, const void* y, const float* mr, const float* gamma, const float* beta,
                       const void* g_same, const void* g_pool, int pool_mode, const void* g_up, int g_up_pooled,
                       const float* red, void* dz, int N, int H, int W, int C, int c_valid, int act, float slope,
                       float* dgamma, float* dbeta, void* stream);
/* dgamma / dbeta (optional, fp32 [c_valid]): += the affine gradients, i.e. what tg_affine_grad computes */
int tg_in_bwd_apply(const void* dn, const void* raw, const float* mr, const float* gamma,
                    const float* red, void* dz, int N, int HW, int C, int c_valid, float* dgamma, float* dbeta,
                    void* stream);
int tg_affine_grad(const float* red, float* dgamma, float* dbeta, int N, int C, int c_valid, void* stream);
int tg_bias_grad(const void* dz, float* db, long long rows, int C, int c_valid, void* stream);
/* InstanceNorm double backward for the gradient penalty (util.py:88-93, create_graph=True) */
int tg_in_bwd2(const void* u, const void* raw, const void* dn, const float* mr, const float* gamma,
               const float* beta, const float* red1, float* red2, void* adj_da, void* adj_z, int N,
               int HW, int C, int c_valid, int act, float slope, void* stream);
int tg_gamma_grad2(const float* red2, float* dgamma, int N, int C, int c_valid, void* stream);
int tg_add(const void* a, const void* b, void* out, long long numel, void* stream);
int tg_act_bwd(const void* g, const void* y, void* out, long long numel, int act, float slope, void* stream);

/* ---- FeatureMapBlock: 1x1 conv + bias (+Tanh) (UNet_plusplus.py:5-16) forward / backward */
/* w: the module's own fp32 [co][ci] weight (ci <= C = 64 real input channels of the 64-padded NHWC input) */
int tg_fmap_fwd(const void* x, const float* w, const float* b, float* out, int N, int HW, int C, int co,
                int use_tanh, int ci, void* stream);
/* dw: fp32 [dw_replicas][co][64] accumulated with atomics (blocks spread over the replicas);
 * tg_fmap_wgrad_fold adds their sum to grad [co][ci] and clears them for the next pass */
int tg_fmap_bwd(const void* x, const float* w, const float* out, const float* g1, const float* g2,
                void* dx, float* dw, int dw_replicas, float* db, int N, int HW, int C, int co, int use_tanh, int ci,
                void* stream);
int tg_fmap_wgrad_fold(float* dw, int dw_replicas, int co, int ci, float* grad, void* stream);

/* ---- version-1 perceptual term, VGGPerceptualLoss (util.py:100-144): input transform (channel repeat,
 *      ImageNet mean / std, bilinear resize align_corners=False) and its transpose, MaxPool2d(2) for layers without
 *      a normalise pass, and the gradient of weight * L1-mean between two feature tensors */
int tg_vgg_prep_fwd(const float* x, void* out, int N, int cs, int H, int W, int OH, int OW, int C, int resize,
                    void* stream);
int tg_vgg_prep_bwd(const void* g, float* grad, int N, int cs, int H, int W, int OH, int OW, int C, int resize,
                    float scale, void* stream);
int tg_pool_fwd(const void* y, void* pool, int N, int H, int W, int C, int mode, void* stream);
int tg_feat_loss_grad(const void* a, const void* b, long long numel, float weight, void* g, void* stream);

/* ---- input pipeline on the device (datasets/PairedDataset.py:30-44,80-92): HorizontalFlip + Affine of a uint8 HWC
 *      image / mask pair (image bilinear, mask nearest, zero border), then ToTensor (+ Normalize(.5,.5) on the image)
 *      into fp32 NCHW. params[n] = {flip, a00, a01, a02, a10, a11, a12, 0}: the inverse map in 16.16 fixed point */
int tg_augment_pair(const void* img_u8, const void* mask_u8, const long long* params, float* out_a, float* out_b,
                    int N, int H, int W, int ca, int cb, void* stream);

/* ---- evaluation (test.py:113-124, eval_pair fuzzy=True): per image stats[n][4] += (sum o*r, sum o^2 + r^2,
 *      sum min(o, r), sum r); accuracy = s2/s3, dice = 2 s0/s1, jaccard = s0/(s1 - s0) */
int tg_eval_fuzzy(const float* out, const float* real, int N, long long per_img, float* stats, void* stream);

/* ---- ConvLSTM / ConvBLSTM skip modules of BCDUNet (generators/BCDUNet.py:6-103; constructed at :145-152, never
 *      called by the reference forward -- SURVEY 8f row 4). The gate conv over cat[X, H_prev] is an ordinary
 *      two-source implicit GEMM (tg_conv_plan_create, 4*C output channels + bias); these two finish the cell:
 *      tg_pack_nchw_tiled: fp32 NCHW time slice (image stride in elements) -> bf16 NHWC, any channel count;
 *      tg_convlstm_gates: i/f/g/o peephole gates (BCDUNet.py:36-45) on z = bf16 NHWC [N][HW][zc] with the gate
 *      groups at channel offsets 0, C, 2C, 3C; W_c* fp32 [C][HW]; cell state and H in fp32 NCHW (image strides in
 *      elements; c_prev NULL = zero state, may alias c_out); H also as bf16 NHWC [N][HW][hc] for the next step's
 *      conv (h_out or h_nhwc may be NULL). act: 3 relu, 4 tanh. */
int tg_pack_nchw_tiled(const float* in, long long in_stride_n, void* out, int N, int C, int HW, int Cpad,
                       void* stream);
int tg_convlstm_gates(const void* z, int zc, const float* w_ci, const float* w_cf, const float* w_co,
                      const float* c_prev, long long c_prev_stride_n, float* c_out, long long c_out_stride_n,
                      float* h_out, long long h_out_stride_n, void* h_nhwc, int hc, int N, int HW, int C, int act,
                      void* stream);

/* Backward of the same cell (what autograd derives from BCDUNet.py:32-47 when the modules are trained): the gates are
 * recomputed from the saved z / c_prev / c; dh = dh_ext (fp32 NCHW gradient of the frame's output, may be NULL) +
 * dh_rec (bf16 NHWC d/dH_prev of the following step, may be NULL); dc_in / dc_out fp32 [N][C][HW] (dc_in NULL = 0,
 * may alias dc_out); dz bf16 NHWC [N][HW][zc] feeds tg_conv_plan (input gradients w.r.t. X and H_prev) and
 * tg_wgrad_plan; dW_c* fp32 [C][HW] are accumulated (+=, deterministic).
 * tg_unpack_nhwc_tiled: bf16 NHWC -> fp32 NCHW with an image stride (the dX time slice), optionally accumulating. */
int tg_convlstm_gates_bwd(const void* z, int zc, const float* w_ci, const float* w_cf, const float* w_co,
                          const float* c_prev, long long c_prev_stride_n, const float* c_cur, long long c_cur_stride_n,
                          const float* dh_ext, long long dh_ext_stride_n, const void* dh_rec, int hc,
                          const float* dc_in, float* dc_out, void* dz, float* dw_ci, float* dw_cf, float* dw_co,
                          int N, int HW, int C, int act, void* stream);
int tg_unpack_nhwc_tiled(const void* in, int Cpad, float* out, long long out_stride_n, int N, int C, int HW,
                         int accumulate, void* stream);

/* gradient-penalty interpolation weights (util.py:79-83): alpha = u, or (u + 1) / 2 for version 2; and 1 - alpha */
int tg_gp_alpha(const float* u, int version, float* alpha, float* one_minus_alpha, int n, void* stream);

/* ---- losses: GANLoss (generators/generators.py:80-105), nn.L1Loss (train.py:145), pan_loss
 *      (util.py:41-70), gradient_penalty norm (util.py:92-93) */
int tg_gan_loss(const void* pred, const float* label, float label_const, int mode, int target_is_real,
                int for_disc, int has_sigmoid, float scale, int n0, int n1, int HW, int C, float* loss,
                void* dz, void* stream);
int tg_gp_first_seed(const void* pred, int has_sigmoid, int n0, int n1, int HW, int C, void* dz, void* stream);
int tg_gp_top(const void* w, const void* pred, int has_sigmoid, long long numel, int C, void* out, void* stream);
int tg_l1_loss(const float* a, const float* b, long long numel, float scale, float* loss, float* grad_a,
               void* stream);
int tg_feat_loss(const void* a, const void* b, long long numel, float weight, int l2, float* loss, void* stream);
int tg_gp_normsq(const void* g, int N, int HW, int C, int c_off, int cj, float* nsq, void* stream);
int tg_gp_finish(const float* nsq, int N, float lambda, float constant, float* loss, float* coef, void* stream);
int tg_gp_seed(const void* g, const float* coef, int N, int HW, int C, int c_off, int cj, void* seed, void* stream);

/* ---- data-parallel gradient exchange for small buffers (new; the reference has no multi-GPU path): one-shot
 * sum over NVLink peer memory. local: this rank's fp32 buffer (numel % 4 == 0), summed in place. peer_bufs_dev /
 * peer_flags_dev: device arrays of `world` pointers to every rank's symmetric staging buffer (2 * half_stride floats,
 * half_stride >= numel: two halves used alternately by epoch parity) and flag
 * pad (ctas * world uint32, zero-initialised) -- memory all ranks of the node can address (torch symmetric memory /
 * cudaIpc). epoch: 1, 2, 3, ... the same on every rank for the same call. Ranks add in the order 0..world-1, so all
 * ranks obtain bit-identical sums. */
int tg_allreduce_oneshot(float* local, const void* peer_bufs_dev, const void* peer_flags_dev, int rank, int world,
                          long long numel, long long half_stride, unsigned epoch, int ctas, void* stream);

/* ---- torch.optim.Adam.step (train.py:135,168) fused with the bf16 weight re-pack.
 * table_dev: device array of `ntensors` rows of 136 bytes: 6 pointers (param, grad, exp_avg,
 * exp_avg_sq, pack_fwd, pack_bwd), int64 numel, int32 kind, kh, kw, dim1, o_pad, i_pad, nseg,
 * seg_end[6], seg_shift[6], pad. */
int tg_adam_step(const void* table_dev, int ntensors, long long max_numel, float lr, float beta1,
                 float beta2, float eps, int step, float grad_scale, void* stream);
/* The same step with its scalars in device memory: hyper_dev = {lr, beta1, beta2, eps, 1 - beta1^step, 1 - beta2^step,
 * grad_scale}. A CUDA graph that captured this launch can be replayed after the host rewrote the seven floats (the
 * learning-rate schedule and the step count are the only per-step scalars of the training iteration). */
int tg_adam_step_dev(const void* table_dev, int ntensors, long long max_numel, const float* hyper_dev, void* stream);
/* dst[0..n) = values_host[0..n), n <= 16: the values are kernel arguments (read from host memory at call time) */
int tg_write_floats(float* dst, const float* values_host, int n, void* stream);

#ifdef __cplusplus
}
#endif
#endif
