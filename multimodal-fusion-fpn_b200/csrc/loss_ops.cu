// Training loss of the path as two kernels per direction: Mix({Dice_loss_jointv2, BCE_Lossv2}) of the reference
// (common/loss.py:9-90) on the (B, n, EW) fp32 prediction and mask.  Replaces ~45 ATen launches (reshape / mul / pow / sum /
// div / binary_cross_entropy and their backward) between the head's forward and backward.  The tensors are a few ten
// thousand elements: the kernels are latency-bound, what matters is that there are three of them.
//   dice = 1 - mean_c 2 (sum p g + 1e-6) / (sum (p^2 + g) + 2e-6)      sums over batch and space, per channel (loss.py:86-90)
//   bce  = mean( -(g log p + (1 - g) log(1 - p)) )                        logs clamped at -100 like torch (loss.py:54)
//   loss = (dice + bce) / 2                                               Mix with unit coefficients (loss.py:23-26)
// All reductions use per-block partials summed in a fixed order (fp64): bitwise reproducible.
#include "common.cuh"

namespace {

constexpr int LT = 256;
constexpr int L_MAXBLOCKS = 128;

__device__ __forceinline__ double block_sum(double v, double* red) {
  __syncthreads();
  red[threadIdx.x] = v;
  __syncthreads();
  for (int o = LT / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  return red[0];
}

// partial[block][k][3] = (sum p g, sum p^2 + g, sum bce terms) over this block's share of the (b, ew) positions of channel k
__global__ void __launch_bounds__(LT) mix_loss_partial_kernel(int B, int n, int64_t EW, const float* __restrict__ pred,
                                                              const float* __restrict__ mask, double* __restrict__ partial) {
  pdl_prologue();
  __shared__ double red[LT];
  const int64_t per = (int64_t)B * EW;
  for (int k = 0; k < n; k++) {
    double a = 0.0, b = 0.0, c = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * LT + threadIdx.x; i < per; i += (int64_t)gridDim.x * LT) {
      const int64_t bb = i / EW, ew = i - bb * EW;
      const int64_t o = (bb * n + k) * EW + ew;
      const float p = pred[o], g = mask[o];
      a += (double)(p * g);
      b += (double)(p * p + g);
      const float lp = fmaxf(logf(p), -100.f), lq = fmaxf(log1pf(-p), -100.f);
      c += (double)(-(g * lp + (1.f - g) * lq));
    }
    a = block_sum(a, red); b = block_sum(b, red); c = block_sum(c, red);
    if (threadIdx.x == 0) {
      double* out = partial + ((int64_t)blockIdx.x * n + k) * 3;
      out[0] = a; out[1] = b; out[2] = c;
    }
  }
}

// out[0..2] = loss, dice, bce; out[3 + 2k], out[4 + 2k] = inter_k, union_k (kept for the backward)
__global__ void __launch_bounds__(32) mix_loss_final_kernel(int nblocks, int n, double count, const double* __restrict__ partial,
                                                            float* __restrict__ out) {
  pdl_prologue();
  if (threadIdx.x != 0) return;
  double dice_sum = 0.0, bce = 0.0;
  for (int k = 0; k < n; k++) {
    double a = 0.0, b = 0.0, c = 0.0;
    for (int r = 0; r < nblocks; r++) {
      const double* q = partial + ((int64_t)r * n + k) * 3;
      a += q[0]; b += q[1]; c += q[2];
    }
    const double inter = a + 1e-6, uni = b + 2e-6;
    dice_sum += 2.0 * inter / uni;
    bce += c;
    out[3 + 2 * k] = (float)inter;
    out[4 + 2 * k] = (float)uni;
  }
  const double dice = 1.0 - dice_sum / n, b2 = bce / count;
  out[0] = (float)((dice + b2) * 0.5);
  out[1] = (float)dice;
  out[2] = (float)b2;
}

// dL/dp = gscale/2 * [ -(2/n) (g union - 2 inter p) / union^2  +  (p - g) / max(p (1 - p), 1e-12) / count ]
__global__ void __launch_bounds__(LT) mix_loss_bwd_kernel(int B, int n, int64_t EW, float inv_count, const float* __restrict__ pred,
                                                          const float* __restrict__ mask, const float* __restrict__ stats,
                                                          const float* __restrict__ gscale, float* __restrict__ dpred) {
  pdl_prologue();
  const int64_t tot = (int64_t)B * n * EW;
  const float gs = 0.5f * (gscale ? gscale[0] : 1.f);
  for (int64_t i = (int64_t)blockIdx.x * LT + threadIdx.x; i < tot; i += (int64_t)gridDim.x * LT) {
    const int k = (int)((i / EW) % n);
    const float inter = stats[3 + 2 * k], uni = stats[4 + 2 * k];
    const float p = pred[i], g = mask[i];
    const float ddice = -(2.f / n) * (g * uni - 2.f * inter * p) / (uni * uni);
    const float dbce = (p - g) / fmaxf(p * (1.f - p), 1e-12f) * inv_count;
    dpred[i] = gs * (ddice + dbce);
  }
}

}  // namespace

extern "C" size_t ffpn_mix_loss_workspace_bytes(int n) { return (size_t)L_MAXBLOCKS * n * 3 * sizeof(double); }

extern "C" int ffpn_mix_loss_fwd(ffpn_ctx* ctx, int64_t B, int n, int64_t EW, const float* pred, const float* mask, void* ws,
                                 size_t ws_bytes, float* out, void* stream) {
  if (!ctx) return 1;
  if (B <= 0 || n <= 0 || EW <= 0) FFPN_FAIL(ctx, "mix_loss_fwd: empty tensor");
  if (ws == nullptr || ws_bytes < ffpn_mix_loss_workspace_bytes(n)) FFPN_FAIL(ctx, "mix_loss_fwd: workspace too small");
  const int g = ffpn_grid_for(B * EW, LT * 4, L_MAXBLOCKS);
  ffpn_launch(mix_loss_partial_kernel, g, LT, 0, (cudaStream_t)stream, (int)B, n, EW, pred, mask, (double*)ws);
  FFPN_CHECK_LAUNCH(ctx, "mix_loss_partial");
  ffpn_launch(mix_loss_final_kernel, 1, 32, 0, (cudaStream_t)stream, g, n, (double)(B * n * EW), (const double*)ws, out);
  FFPN_CHECK_LAUNCH(ctx, "mix_loss_final");
  return 0;
}

extern "C" int ffpn_mix_loss_bwd(ffpn_ctx* ctx, int64_t B, int n, int64_t EW, const float* pred, const float* mask,
                                 const float* stats, const float* grad_scale, float* dpred, void* stream) {
  if (!ctx) return 1;
  const int g = ffpn_grid_for(B * n * EW, LT * 2, ctx->num_sms * 4);
  ffpn_launch(mix_loss_bwd_kernel, g, LT, 0, (cudaStream_t)stream, (int)B, n, EW, 1.f / (float)(B * n * EW), pred, mask, stats,
              grad_scale, dpred);
  FFPN_CHECK_LAUNCH(ctx, "mix_loss_bwd");
  return 0;
}
