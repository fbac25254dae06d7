// Input pipeline and evaluation-metric kernels around the hot path: they keep the per-step
// host work (numpy z-scoring of every B-scan, .cpu().numpy() metric updates with a device sync each) off the critical path once
// the model itself runs at hundreds of samples per second.
//   * z-score per B-scan: ZScoreNormalization(axis=(2,3)) of common/mytransforms.py:277-296 as configured by
//     training_config.py:60 -- for every (batch, B-scan) slice of the (B,1,S,H,W) volume: (x - mean) / (std + 1e-8), population std.
//   * Dice metric: common/metrics.py:216-253 -- per sample 2 |P & G| / (|P| + |G|) on thresholded prediction / mask, 1 when both
//     are empty; the per-sample values stay on the device until the epoch end.
#include "common.cuh"

namespace {

constexpr int PT = 256;

__device__ __forceinline__ double block_sum_d(double v, double* red) {
  __syncthreads();
  red[threadIdx.x] = v;
  __syncthreads();
  for (int o = PT / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  return red[0];
}

// one block per slice: mean and 1 / (std + eps) (fp64 accumulation, fixed order -> deterministic)
__global__ void __launch_bounds__(PT) slice_stats_kernel(int64_t n, const float* __restrict__ x, float eps, float* __restrict__ out) {
  pdl_prologue();
  __shared__ double red[PT];
  const float* p = x + (int64_t)blockIdx.x * n;
  double s = 0.0, q = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += PT) {
    const double v = (double)p[i];
    s += v;
    q += v * v;
  }
  s = block_sum_d(s, red);
  q = block_sum_d(q, red);
  if (threadIdx.x == 0) {
    const double mean = s / (double)n;
    double var = q / (double)n - mean * mean;
    if (var < 0.0) var = 0.0;
    out[2 * blockIdx.x] = (float)mean;
    out[2 * blockIdx.x + 1] = (float)(1.0 / (sqrt(var) + (double)eps));
  }
}

__global__ void __launch_bounds__(PT) slice_normalize_kernel(int64_t R, int64_t n, const float* __restrict__ x,
                                                             const float* __restrict__ st, float* __restrict__ y) {
  pdl_prologue();
  const int64_t tot = R * n;
  for (int64_t i = (int64_t)blockIdx.x * PT + threadIdx.x; i < tot; i += (int64_t)gridDim.x * PT) {
    const int64_t r = i / n;
    y[i] = (x[i] - st[2 * r]) * st[2 * r + 1];
  }
}

// one block per sample: counts[b] = (|P & G|, |P| + |G|) over elements [b][slice][*]
__global__ void __launch_bounds__(PT) dice_counts_kernel(int64_t per_sample, int64_t per_channel, int slice, float pt, float gt,
                                                         const float* __restrict__ pred, const float* __restrict__ mask,
                                                         float* __restrict__ out) {
  pdl_prologue();
  __shared__ double red[PT];
  const int64_t base = (int64_t)blockIdx.x * per_sample + (int64_t)slice * per_channel;
  double a = 0.0, b = 0.0;
  for (int64_t i = threadIdx.x; i < per_channel; i += PT) {
    const float p = pred[base + i] > pt ? 1.f : 0.f, g = mask[base + i] > gt ? 1.f : 0.f;
    a += (double)(p * g);
    b += (double)(p + g);
  }
  a = block_sum_d(a, red);
  b = block_sum_d(b, red);
  if (threadIdx.x == 0) out[blockIdx.x] = b == 0.0 ? 1.f : (float)(2.0 * a / b);
}

}  // namespace

extern "C" int ffpn_zscore_slices(ffpn_ctx* ctx, int64_t R, int64_t n, const float* x, float eps, float* stats, float* y, void* stream) {
  if (!ctx) return 1;
  if (R <= 0 || n <= 0 || R >= (1ll << 31)) FFPN_FAIL(ctx, "zscore_slices: bad extent");
  ffpn_launch(slice_stats_kernel, (int)R, PT, 0, (cudaStream_t)stream, n, x, eps, stats);
  FFPN_CHECK_LAUNCH(ctx, "slice_stats");
  ffpn_launch(slice_normalize_kernel, ffpn_grid_for(R * n, PT * 4, ctx->num_sms * 8), PT, 0, (cudaStream_t)stream, R, n, x, (const float*)stats, y);
  FFPN_CHECK_LAUNCH(ctx, "slice_normalize");
  return 0;
}

extern "C" int ffpn_dice_metric(ffpn_ctx* ctx, int64_t B, int n_channels, int64_t per_channel, int slice, float pred_threshold,
                                float target_threshold, const float* pred, const float* mask, float* out, void* stream) {
  if (!ctx) return 1;
  if (B <= 0 || slice < 0 || slice >= n_channels) FFPN_FAIL(ctx, "dice_metric: slice %d outside %d channels", slice, n_channels);
  ffpn_launch(dice_counts_kernel, (int)B, PT, 0, (cudaStream_t)stream, (int64_t)n_channels * per_channel, per_channel, slice, pred_threshold,
              target_threshold, pred, mask, out);
  FFPN_CHECK_LAUNCH(ctx, "dice_metric");
  return 0;
}
