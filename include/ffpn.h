/*
 * ffpn.h -- C ABI of libfusionfpn.so: the sm_100a CUDA kernels behind the projective multimodal
 * fusion FPN forward/backward (reference: j-morano/multimodal-fusion-fpn, models/fusion_nets.py +
 * models/fpn/*).
 *
 * The reference has no FFI: its "operator boundary" is the set of ATen ops its nn.Modules dispatch
 * to.  Every entry point below replaces one of those ops (or a fused group of
 * them) and cites the reference call site it stands in for.  The Python host mirror
 * (multimodal-fusion-fpn_b200/ffpn/) binds them with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch types.  All tensor pointers are DEVICE pointers that
 *    the caller owns (PyTorch caching allocator); the library borrows them for the call.
 *  - every function returns 0 on success, non-zero on error (never throws/aborts); the message is
 *    available from ffpn_last_error(ctx).
 *  - all work is enqueued on the cudaStream_t passed in (as void*); no hidden synchronisation.
 *  - activations are channels-last: (B, S, W, H, C) contiguous, i.e. the reference's logical
 *    (B, C, S, W, H) tensor in torch.channels_last_3d memory format.  2-D features are H == 1.
 *    dtype: FFPN_F32 (exact path) or FFPN_BF16 (bf16 storage, fp32 accumulate).
 *  - weights, BatchNorm vectors and all statistics are fp32; conv weights stay in the reference's
 *    layout [Cout, Cin, kS, kW, kH] (state_dict compatible).
 */
#ifndef FFPN_H
#define FFPN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FFPN_ABI_VERSION 3
#define FFPN_F32 0
#define FFPN_BF16 1
/* rows of a per-block partial-statistics buffer: [FFPN_STAT_ROWS][ncols] floats */
#define FFPN_STAT_ROWS 1184

typedef struct ffpn_ctx ffpn_ctx;

/* Convolution geometry.  Only what the reference uses: stride and padding per axis, no dilation, no
 * groups, bias-free (every conv on the path is bias=False except final1, see ffpn_head_fwd). */
typedef struct ffpn_conv_desc {
  int64_t B, S, W, H;      /* input extent  */
  int64_t oS, oW, oH;      /* output extent */
  int32_t Cin, Cout;
  int32_t kS, kW, kH;
  int32_t sS, sW, sH;
  int32_t pS, pW, pH;
  int32_t dtype;           /* FFPN_F32 | FFPN_BF16 (activations in and out) */
  int32_t impl;            /* 0 = auto, 1 = force SIMT kernel, 2 = force tcgen05 kernel (error if unsupported) */
} ffpn_conv_desc;

int ffpn_abi_version(void);
int ffpn_create(ffpn_ctx** ctx, int device);
void ffpn_destroy(ffpn_ctx* ctx);
const char* ffpn_last_error(ffpn_ctx* ctx);
/* number of kernel launches issued through this ctx since creation (bench.py's gpu_launches) */
int64_t ffpn_launch_count(ffpn_ctx* ctx);
/* conv calls per kernel family since creation: out4 = {warp-specialised tcgen05, Cin==1 stem, first-generation tcgen05,
 * CUDA-core}.  The dispatch is not silent: bench.py prints the per-step counts and refuses a bf16 run whose convs left the
 * tcgen05 / stem kernels. */
int ffpn_route_counts(ffpn_ctx* ctx, int64_t* out4);
/* bit 0: built with -DFFPN_DEBUG (the result-invalidating ablation / trace switches FFPN_TC_DEBUG, FFPN_WS_TRACE are
 * compiled in).  The release library returns 0 and ignores those variables. */
int ffpn_build_info(void);
/* Host-only plan introspection (no GPU needed): writes a one-line description of how the forward (transposed = 0), dgrad
 * (transposed = 1) or weight gradient (transposed = 2) of this conv runs on the warp-specialised tcgen05 kernels, or that it is
 * not taken by them ("not on the warp-specialised kernel ..."). */
int ffpn_conv_plan_info(const ffpn_conv_desc* d, int transposed, char* buf, size_t n);
/* bytes of scratch a conv call of this geometry may use (packed bf16 weights for the tcgen05 path) */
size_t ffpn_conv_workspace_bytes(const ffpn_conv_desc* d);

/* ---- convolution (replaces aten::convolution / convolution_backward -> cuDNN; reference call sites
 *      fusion3D2D.py:601-647 (3-D), :762-808 (2-D), :232,267,318-324,925-931) ------------------------
 * fwd:  y = conv(f(x), w),  f(x) = in_relu ? max(in_scale*x+in_shift, 0) : in_scale*x+in_shift  applied on
 *       load (the producer's BatchNorm+ReLU, fusion3D2D.py:609-610), identity when in_scale == NULL.
 *       Zero padding is applied AFTER f, as in the reference.  When stat_partial != NULL the kernel also
 *       writes per-block partial sums of y and y*y: [rows][2][Cout] floats, *stat_rows = rows written. */
int ffpn_conv_fwd(ffpn_ctx* ctx, const ffpn_conv_desc* d, const void* x, const float* in_scale,
                  const float* in_shift, int in_relu, const float* w, void* y, float* stat_partial,
                  int* stat_rows, void* ws, size_t ws_bytes, void* stream);
/* conv_fwd followed by the BatchNorm statistics finalize of its output (ffpn_bn_finalize) as ONE call: on the tcgen05
 * path the last CTA of the conv kernel combines the partial sums itself, which removes a latency-bound launch per
 * conv (fusion3D2D.py:717-732: every conv is followed by a BatchNorm in training mode).  Same arguments as the two calls. */
int ffpn_conv_fwd_bn(ffpn_ctx* ctx, const ffpn_conv_desc* d, const void* x, const float* in_scale, const float* in_shift,
                     int in_relu, const float* w, void* y, float* stat_partial, int* stat_rows, double count,
                     const float* gamma, const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                     int training, float* scale, float* shift, float* save_mean, float* save_invstd, void* ws, size_t ws_bytes,
                     void* stream);

/* dgrad: dx = conv_transpose(dy, w) [+ addend]  (gradient w.r.t. f(x), i.e. before the producer's ReLU
 *        mask).  addend (nullable, same shape/dtype as dx) fuses the residual-branch gradient sum of
 *        fusion3D2D.py:724-725's backward. */
int ffpn_conv_dgrad(ffpn_ctx* ctx, const ffpn_conv_desc* d, const void* dy, const float* w,
                    const void* addend, void* dx, void* ws, size_t ws_bytes, void* stream);
/* dgrad + pass 1 of the backward of the BatchNorm + ReLU that produced the conv's input, as ONE call (the backward of
 *        fusion3D2D.py:717-732's conv -> BN -> ReLU -> conv: aten::convolution_backward's grad_input followed by
 *        aten::threshold_backward and the reductions of aten::native_batch_norm_backward).  y_prev is that BatchNorm's raw
 *        input (the previous conv's output, same shape/dtype as dx), bn_scale / bn_shift its affine.  dx receives
 *        conv_transpose(dy, w) exactly as ffpn_conv_dgrad writes it; partial receives [rows][2][Cin] = sums of G and
 *        G*y_prev with G = dx * (bn_scale*y_prev + bn_shift > 0), what ffpn_bn_bwd_reduce(dx, y_prev, relu=1) produces.  On the tcgen05 path the sums come
 *        out of the dgrad kernel's epilogue (no extra pass over dx and y_prev); otherwise the two kernels are launched. */
int ffpn_conv_dgrad_bnr(ffpn_ctx* ctx, const ffpn_conv_desc* d, const void* dy, const float* w, const void* y_prev,
                        const float* bn_scale, const float* bn_shift, void* dx, float* partial, int* rows, void* ws,
                        size_t ws_bytes, void* stream);
/* wgrad: dw += sum_pos dy[pos] (x) f(x)[pos*stride - pad + tap]; dw is fp32 [Cout,Cin,kS,kW,kH] and is
 *        ACCUMULATED into (caller zeroes it). */
int ffpn_conv_wgrad(ffpn_ctx* ctx, const ffpn_conv_desc* d, const void* x, const float* in_scale,
                    const float* in_shift, int in_relu, const void* dy, float* dw, void* ws,
                    size_t ws_bytes, void* stream);

/* ---- BatchNorm (replaces aten::native_batch_norm(+_backward); nn.BatchNorm3d/2d at fusion3D2D.py:609,
 *      622,634,647,770,...).  Training: batch mean / biased variance from the conv's partial sums,
 *      running stats momentum/unbiased update in place; eval: running stats.  Emits the affine
 *      scale = gamma*invstd, shift = beta - mean*scale that consumers apply on load. */
int ffpn_bn_finalize(ffpn_ctx* ctx, const float* stat_partial, int stat_rows, int C, double count,
                     const float* gamma, const float* beta, float* running_mean, float* running_var,
                     float momentum, float eps, int training, float* scale, float* shift,
                     float* save_mean, float* save_invstd, void* stream);
/* backward pass 1: G = dA * (relu ? (scale*y+shift > 0) : 1); partial sums of G and G*y ->
 *      [rows][2][C].  P = number of positions. */
int ffpn_bn_bwd_reduce(ffpn_ctx* ctx, int dtype, int64_t P, int C, const void* dA, const void* y,
                       const float* scale, const float* shift, int relu, float* partial, int* rows,
                       void* stream);
/* backward pass 2 (tiny): dgamma, dbeta (written) and the coefficients of
 *      dy = cA*G + cP*y + cQ   (per channel).  ycol selects which "G*y" column of a multi-column partial
 *      buffer to use (ffpn_block_end_bwd writes [rows][ncols][C] with col0 = sum G). */
int ffpn_bn_bwd_finalize(ffpn_ctx* ctx, const float* partial, int rows, int ncols, int ycol, int C,
                         double count, const float* gamma, const float* save_mean,
                         const float* save_invstd, float* dgamma, float* dbeta, float* cA, float* cP,
                         float* cQ, void* stream);
/* backward pass 3: dy = cA * (dA * mask) + cP*y + cQ   (mask as in pass 1; dy may alias dA) */
int ffpn_bn_bwd_apply(ffpn_ctx* ctx, int dtype, int64_t P, int C, const void* dA, const void* y,
                      const float* scale, const float* shift, int relu, const float* cA,
                      const float* cP, const float* cQ, void* dy, void* stream);

/* backward pass 3 of a block whose shortcut is conv + BN (fusion3D2D.py:717-727: out = relu(bn_k(conv_k(..)) + bn_s(conv_s(x)))):
 *      both BatchNorms receive the same gradient G (ffpn_block_end_bwd's output, already ReLU-masked), so
 *      dy1 = cA1*G + cP1*y1 + cQ1 and dy2 = cA2*G + cP2*y2 + cQ2 are produced by ONE pass that reads G once -- the same values
 *      as two ffpn_bn_bwd_apply(relu = 0) calls. */
int ffpn_bn_bwd_apply2(ffpn_ctx* ctx, int dtype, int64_t P, int C, const void* G, const void* y1, const void* y2,
                       const float* cA1, const float* cP1, const float* cQ1, const float* cA2, const float* cP2,
                       const float* cQ2, void* dy1, void* dy2, void* stream);

/* ---- residual block end (replaces BN-apply + add_ + relu_, fusion3D2D.py:724-727) --------------------
 * z = relu(a*y + b + r),  r = ra*res + rb (raw shortcut conv output), res (identity shortcut, ra==NULL)
 * or 0 (res == NULL, non-residual block). */
int ffpn_block_end_fwd(ffpn_ctx* ctx, int dtype, int64_t P, int C, const void* y, const float* a,
                       const float* b, const void* res, const float* ra, const float* rb, void* z,
                       void* stream);
/* G = (dz + route(dzp)) * (z > 0): gradient at the block-end ReLU input.  dzp (nullable) is the gradient
 * of the max-pooled tensor, routed to each window's first maximum (NaN wins) exactly like
 * max_pool3d_with_indices_backward (fusion3D2D.py:87-90); pool kernel (kS,kW,kH),
 * stride = kernel, floor.  Also writes partial sums [rows][ncols][C]: col0 = sum G, col1 = sum G*y,
 * col2 = sum G*yres (when yres != NULL); ncols = yres ? 3 : 2. */
int ffpn_block_end_bwd(ffpn_ctx* ctx, int dtype, int64_t B, int64_t S, int64_t W, int64_t H, int C,
                       int kS, int kW, int kH, const void* dz, const void* dzp, const void* z,
                       const void* y, const void* yres, void* G, float* partial, int* rows,
                       void* stream);

/* ---- max pooling (replaces aten::max_pool3d/2d_with_indices, fusion3D2D.py:87-90,168-171) -----------
 * idx (nullable): int64 flat index into (S,W,H) per (b,c), PyTorch's convention, laid out like zp. */
int ffpn_maxpool_fwd(ffpn_ctx* ctx, int dtype, int64_t B, int64_t S, int64_t W, int64_t H, int C, int kS,
                     int kW, int kH, const void* z, void* zp, int64_t* idx, void* stream);
/* stand-alone backward (argmax recomputed from z with the same rule): dz = route(dzp), zeros elsewhere */
int ffpn_maxpool_bwd(ffpn_ctx* ctx, int dtype, int64_t B, int64_t S, int64_t W, int64_t H, int C, int kS,
                     int kW, int kH, const void* z, const void* dzp, void* dz, void* stream);

/* ---- projection tail (replaces BN+ReLU of the (1,1,4) block and torch.mean(dim=4), fusion3D2D.py:527-536)
 * out[b,s,w,coff+c] = mean_h relu(a*y[b,s,w,h,c]+b); out has channel stride ostride (a concat buffer). */
int ffpn_proj_tail_fwd(ffpn_ctx* ctx, int dtype, int64_t EW, int64_t H, int C, const void* y,
                       const float* a, const float* b, void* out, int ostride, int coff, void* stream);
/* dA[b,s,w,h,c] = dout[b,s,w,coff+c] / H   (gradient w.r.t. the ReLU output; mask applied by bn_bwd_*) */
int ffpn_proj_tail_bwd(ffpn_ctx* ctx, int dtype, int64_t EW, int64_t H, int C, const void* dout,
                       int ostride, int coff, void* dA, void* stream);

/* ---- 2-D feature resize into a concat slice (fusion3D2D.py:544-564) ---------------------------------
 * mode 0: copy (interpolate=None), 1: adaptive max (F.adaptive_max_pool3d), 2: bilinear, half-pixel,
 * align_corners=False (F.interpolate 'trilinear' with depth 1).  idx (mode 1): int32 argmax into Si*Wi. */
int ffpn_resize2d_fwd(ffpn_ctx* ctx, int dtype, int mode, int64_t B, int64_t Si, int64_t Wi, int64_t So,
                      int64_t Wo, int C, const void* x, void* out, int ostride, int coff, int32_t* idx,
                      void* stream);
int ffpn_resize2d_bwd(ffpn_ctx* ctx, int dtype, int mode, int64_t B, int64_t Si, int64_t Wi, int64_t So,
                      int64_t Wo, int C, const void* dout, int ostride, int coff, const int32_t* idx,
                      void* dx, void* stream);

/* ---- nearest upsample by integer factors into a concat slice (Upsample_Custom3d_nearest,
 *      components.py:259-268: idx = ceil((i+1)/f) - 1 == i / f for integer f) --------------------------- */
int ffpn_upsample_fwd(ffpn_ctx* ctx, int dtype, int64_t B, int64_t Si, int64_t Wi, int fS, int fW, int C,
                      const void* x, void* out, int ostride, int coff, void* stream);
int ffpn_upsample_bwd(ffpn_ctx* ctx, int dtype, int64_t B, int64_t Si, int64_t Wi, int fS, int fW, int C,
                      const void* dout, int ostride, int coff, void* dx, void* stream);
/* copy a (P, C) tensor into / out of a channel slice (torch.cat and its backward, fusion3D2D.py:572,966) */
int ffpn_slice_copy(ffpn_ctx* ctx, int dtype, int64_t P, int C, const void* src, int sstride, int soff,
                    void* dst, int dstride, int doff, void* stream);

/* ---- head: final1 = Conv3d(C -> n, 1x1x1, bias) (fusion3D2D.py:223,579), optionally with the wrapper's sigmoid
 *      (fusion_nets.py:110,118) fused in: act 0 = logits, 1 = sigmoid.  Output fp32 in the reference layout (B, n, S, W, 1). - */
int ffpn_head_fwd(ffpn_ctx* ctx, int dtype, int64_t B, int64_t EW, int C, int n, int act, const void* x,
                  const float* w, const float* bias, float* out, void* stream);
/* dx (activations dtype), dw[n][C] and dbias[n] (fp32, written) from dout = dL/d(out).  pred != NULL: out was the sigmoid and
 * pred is that output (the gradient of the pre-activation is formed on the fly as dout * pred * (1 - pred)).  ws: scratch of
 * ffpn_head_bwd_workspace_bytes(C, n) bytes for the two-stage fixed-order weight-gradient reduction. */
size_t ffpn_head_bwd_workspace_bytes(int C, int n);
int ffpn_head_bwd(ffpn_ctx* ctx, int dtype, int64_t B, int64_t EW, int C, int n, const void* x,
                  const float* w, const float* dout, const float* pred, void* dx, float* dw, float* dbias, void* ws,
                  size_t ws_bytes, void* stream);

/* ---- training loss: Mix({Dice_loss_jointv2, BCE_Lossv2}) with unit coefficients (common/loss.py:9-90) on the fp32
 *      prediction and mask, both (B, n, EW) contiguous.  fwd: out[0..2] = loss, dice, bce and out[3 + 2k], out[4 + 2k] = the
 *      per-channel Dice sums the backward needs (out has 3 + 2n floats).  bwd: dpred = grad_scale[0] * dloss/dpred
 *      (grad_scale: device scalar, NULL = 1).  Replaces the ~45 ATen launches of the loss and its autograd backward. ------- */
size_t ffpn_mix_loss_workspace_bytes(int n);
int ffpn_mix_loss_fwd(ffpn_ctx* ctx, int64_t B, int n, int64_t EW, const float* pred, const float* mask, void* ws,
                      size_t ws_bytes, float* out, void* stream);
int ffpn_mix_loss_bwd(ffpn_ctx* ctx, int64_t B, int n, int64_t EW, const float* pred, const float* mask,
                      const float* stats, const float* grad_scale, float* dpred, void* stream);

/* ---- input packing: fp32 (R, H, W) -> activations dtype (R, W, H)  (the permute of
 *      fusion_nets.py:114 made physical; R = B*S) ----------------------------------------------------- */
int ffpn_pack_volume(ffpn_ctx* ctx, int dtype, int64_t R, int64_t H, int64_t W, const float* src, void* dst,
                     void* stream);
int ffpn_cast(ffpn_ctx* ctx, int dtype, int64_t n, const float* src, void* dst, void* stream);

/* ---- input pipeline / evaluation metric around the path ---------------------------------
 * z-score of every one of R slices of n contiguous fp32 values: y = (x - mean) / (std + eps), population std
 * (ZScoreNormalization(axis=(2,3)), common/mytransforms.py:277-296 with training_config.py:60: a slice = one B-scan of the
 * (B,1,S,H,W) volume, R = B*S, n = H*W).  stats: scratch of 2*R floats (mean, 1/(std+eps) per slice); y may alias x. */
int ffpn_zscore_slices(ffpn_ctx* ctx, int64_t R, int64_t n, const float* x, float eps, float* stats, float* y, void* stream);
/* Dice metric of common/metrics.py:216-253: out[b] = 2 |P & G| / (|P| + |G|) for sample b on channel `slice` of the
 * (B, n_channels, per_channel) prediction / mask thresholded at pred_threshold / target_threshold; 1 when both are empty. */
int ffpn_dice_metric(ffpn_ctx* ctx, int64_t B, int n_channels, int64_t per_channel, int slice, float pred_threshold,
                     float target_threshold, const float* pred, const float* mask, float* out, void* stream);

/* ---- optimiser: torch.optim.SGD(momentum, weight_decay) of train.py:126-133 on flat fp32 buffers;
 *      g is first scaled by grad_scale (1/world_size after the NCCL sum all-reduce). ----------------- */
int ffpn_sgd_step(ffpn_ctx* ctx, int64_t n, float* p, const float* g, float* mom, float lr, float momentum,
                  float weight_decay, float grad_scale, int first_step, void* stream);

/* ---- packed-weight arena (optional, training loops): the tcgen05 kernels consume bf16 weight images that the
 *      library packs from the fp32 master weights.  Without an arena every conv call packs its own image.  With
 *      one, the images of a whole step are recorded once (begin .. one forward+backward .. seal) into a caller-owned
 *      device buffer and afterwards regenerated by ONE kernel per step (ffpn_weight_arena_pack, to be called after
 *      the weights changed and before the next forward).  Calls whose image is not in the arena fall back to
 *      per-call packing.  The reference has no counterpart (cuDNN reads fp32 weights directly). ------------------- */
int ffpn_weight_arena_begin(ffpn_ctx* ctx, void* arena, size_t bytes);
int ffpn_weight_arena_seal(ffpn_ctx* ctx);
int ffpn_weight_arena_pack(ffpn_ctx* ctx, void* stream);
int ffpn_weight_arena_end(ffpn_ctx* ctx);
/* Images are only recorded into / taken from the arena while lookups are enabled (begin enables them).  The owner of the
 * arena (FusionTrainer.forward_backward) enables them around its own forward+backward and disables them afterwards, so
 * conv calls made by anyone else on the same device (an eval forward, another model, weights loaded behind the
 * trainer's back) pack their image from the CURRENT fp32 weights and can never read a stale one. */
int ffpn_weight_arena_enable(ffpn_ctx* ctx, int on);

#ifdef __cplusplus
}
#endif
#endif /* FFPN_H */
