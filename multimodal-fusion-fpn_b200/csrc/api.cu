// C-ABI entry points that are not defined next to their kernels: ctx lifetime and the convolution dispatcher.
#include <stdlib.h>

#include "common.cuh"

int ffpn_conv_fwd_simt(ffpn_ctx*, const ffpn_conv_desc*, const void*, const float*, const float*, int, const float*,
                       void*, float*, int*, cudaStream_t);
int ffpn_conv_dgrad_simt(ffpn_ctx*, const ffpn_conv_desc*, const void*, const float*, const void*, void*, cudaStream_t);
int ffpn_conv_wgrad_simt(ffpn_ctx*, const ffpn_conv_desc*, const void*, const float*, const float*, int, const void*,
                         float*, cudaStream_t);
// conv_stem.cu
bool ffpn_stem_supported(const ffpn_conv_desc* d);
int ffpn_stem_fwd(ffpn_ctx*, const ffpn_conv_desc*, const void*, const float*, void*, float*, int*, cudaStream_t);
int ffpn_stem_wgrad(ffpn_ctx*, const ffpn_conv_desc*, const void*, const void*, float*, void*, size_t, cudaStream_t);
size_t ffpn_stem_wgrad_workspace_bytes(const ffpn_conv_desc*);
// conv_tc.cu
bool ffpn_tc_fwd_supported(const ffpn_conv_desc* d);
bool ffpn_tc_dgrad_supported(const ffpn_conv_desc* d);
bool ffpn_tc_wgrad_supported(const ffpn_conv_desc* d);
size_t ffpn_tc_workspace_bytes(const ffpn_conv_desc* d);
int ffpn_conv_fwd_tc(ffpn_ctx*, const ffpn_conv_desc*, bool transposed, const void*, const float*, const float*, int,
                     const float*, const void* addend, void*, float*, int*, void*, size_t, cudaStream_t);
int ffpn_conv_wgrad_tc(ffpn_ctx*, const ffpn_conv_desc*, const void*, const float*, const float*, int, const void*,
                       float*, void*, size_t, cudaStream_t);
// conv_ws.cu (fin != nullptr: BatchNorm finalize fused into the kernel)
struct ffpn_bn_fin;
int ffpn_conv_fwd_ws_bn(ffpn_ctx*, const ffpn_conv_desc*, const void*, const float*, const float*, int, const float*, void*, float*, int*,
                        void*, size_t, cudaStream_t, double count, float momentum, float eps, const float* gamma, const float* beta,
                        float* rmean, float* rvar, float* scale, float* shift, float* smean, float* sinvstd);
int ffpn_conv_dgrad_ws_bnr(ffpn_ctx*, const ffpn_conv_desc*, const void*, const float*, const void*, const float*, const float*, void*, float*,
                           int*, void*, size_t, cudaStream_t);
extern "C" int ffpn_bn_bwd_reduce(ffpn_ctx* ctx, int dtype, int64_t P, int C, const void* dA, const void* y, const float* scale,
                                  const float* shift, int relu, float* partial, int* rows, void* stream);
extern "C" int ffpn_bn_finalize(ffpn_ctx* ctx, const float* stat_partial, int stat_rows, int C, double count, const float* gamma,
                                const float* beta, float* running_mean, float* running_var, float momentum, float eps, int training,
                                float* scale, float* shift, float* save_mean, float* save_invstd, void* stream);

static int check_desc(ffpn_ctx* ctx, const ffpn_conv_desc* d, const char* who) {
  if (!ctx) return 1;
  if (!d) FFPN_FAIL(ctx, "%s: null descriptor", who);
  if (d->dtype != FFPN_F32 && d->dtype != FFPN_BF16) FFPN_FAIL(ctx, "%s: unknown dtype %d", who, d->dtype);
  if (d->B <= 0 || d->S <= 0 || d->W <= 0 || d->H <= 0 || d->Cin <= 0 || d->Cout <= 0) FFPN_FAIL(ctx, "%s: empty tensor", who);
  if (d->kS <= 0 || d->kW <= 0 || d->kH <= 0 || d->sS <= 0 || d->sW <= 0 || d->sH <= 0) FFPN_FAIL(ctx, "%s: bad kernel/stride", who);
  const int64_t eS = (d->S + 2 * d->pS - d->kS) / d->sS + 1, eW = (d->W + 2 * d->pW - d->kW) / d->sW + 1,
                eH = (d->H + 2 * d->pH - d->kH) / d->sH + 1;
  if (eS != d->oS || eW != d->oW || eH != d->oH || eS <= 0 || eW <= 0 || eH <= 0)
    FFPN_FAIL(ctx, "%s: output extent (%lld,%lld,%lld) inconsistent with geometry (expected %lld,%lld,%lld)", who,
              (long long)d->oS, (long long)d->oW, (long long)d->oH, (long long)eS, (long long)eW, (long long)eH);
  if (d->B * d->S * d->W * d->H >= (1ll << 31) || d->B * d->oS * d->oW * d->oH >= (1ll << 31))
    FFPN_FAIL(ctx, "%s: more than 2^31 positions", who);
  return 0;
}

extern "C" int ffpn_abi_version(void) { return FFPN_ABI_VERSION; }

extern "C" int ffpn_create(ffpn_ctx** out, int device) {
  if (!out) return 1;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return 2;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return 3;
  if (prop.major != 10) return 4;   // sm_100a only: no fallback path exists
  ffpn_ctx* c = new ffpn_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  c->launches = 0;
  c->err[0] = 0;
  c->arena_state = 0; c->arena = nullptr; c->arena_bytes = c->arena_used = 0; c->njobs = 0; c->arena_elems = 0; c->d_jobs = nullptr; c->d_counter = nullptr;
  c->fin_slot = 0; c->arena_lookup = 0; c->attr_mask = 0;
  for (int i = 0; i < 4; i++) c->routes[i] = 0;
  *out = c;
  return 0;
}

extern "C" void ffpn_destroy(ffpn_ctx* ctx) {
  if (ctx && ctx->d_jobs) cudaFree(ctx->d_jobs);
  if (ctx && ctx->d_counter) cudaFree(ctx->d_counter);
  delete ctx;
}
extern "C" const char* ffpn_last_error(ffpn_ctx* ctx) { return ctx ? ctx->err : "null ctx"; }
extern "C" int64_t ffpn_launch_count(ffpn_ctx* ctx) { return ctx ? ctx->launches : -1; }
extern "C" int ffpn_route_counts(ffpn_ctx* ctx, int64_t* out4) {
  if (!ctx || !out4) return 1;
  for (int i = 0; i < 4; i++) out4[i] = ctx->routes[i];
  return 0;
}
extern "C" int ffpn_build_info(void) {
#ifdef FFPN_DEBUG
  return 1;
#else
  return 0;
#endif
}

extern "C" size_t ffpn_conv_workspace_bytes(const ffpn_conv_desc* d) {
  if (!d) return 0;
  size_t n = ffpn_tc_workspace_bytes(d);
  if (ffpn_stem_supported(d)) { const size_t s = ffpn_stem_wgrad_workspace_bytes(d); if (s > n) n = s; }
  return n;
}

extern "C" int ffpn_conv_fwd(ffpn_ctx* ctx, const ffpn_conv_desc* d, const void* x, const float* in_scale,
                             const float* in_shift, int in_relu, const float* w, void* y, float* stat_partial,
                             int* stat_rows, void* ws, size_t ws_bytes, void* stream) {
  if (check_desc(ctx, d, "conv_fwd")) return 1;
  if ((in_scale == nullptr) != (in_shift == nullptr)) FFPN_FAIL(ctx, "conv_fwd: in_scale/in_shift must both be set or both null");
  if (stat_partial != nullptr && stat_rows == nullptr) FFPN_FAIL(ctx, "conv_fwd: stat_rows is null");
  if (d->impl == 0 && in_scale == nullptr && ffpn_stem_supported(d)) {
    ctx->routes[FFPN_ROUTE_STEM]++;
    return ffpn_stem_fwd(ctx, d, x, w, y, stat_partial, stat_rows, (cudaStream_t)stream);
  }
  const bool tc_ok = ffpn_tc_fwd_supported(d);
  if (d->impl == 2 && !tc_ok) FFPN_FAIL(ctx, "conv_fwd: tcgen05 kernel does not support this geometry");
  if (tc_ok && d->impl != 1) {
    const int r = ffpn_conv_fwd_tc(ctx, d, false, x, in_scale, in_shift, in_relu, w, nullptr, y, stat_partial, stat_rows, ws, ws_bytes, (cudaStream_t)stream);
    if (r >= 0) return r;
    if (d->impl == 2) FFPN_FAIL(ctx, "conv_fwd: the tcgen05 kernel declined this call (input transform without ReLU, or workspace too small)");
  }
  ctx->routes[FFPN_ROUTE_SIMT]++;
  if (d->dtype == FFPN_BF16) ffpn_log_route("conv_fwd -> CUDA-core kernel", d);
  return ffpn_conv_fwd_simt(ctx, d, x, in_scale, in_shift, in_relu, w, y, stat_partial, stat_rows, (cudaStream_t)stream);
}

extern "C" int ffpn_conv_fwd_bn(ffpn_ctx* ctx, const ffpn_conv_desc* d, const void* x, const float* in_scale, const float* in_shift,
                                int in_relu, const float* w, void* y, float* stat_partial, int* stat_rows, double count,
                                const float* gamma, const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                                int training, float* scale, float* shift, float* save_mean, float* save_invstd, void* ws, size_t ws_bytes,
                                void* stream) {
  if (check_desc(ctx, d, "conv_fwd_bn")) return 1;
  if ((in_scale == nullptr) != (in_shift == nullptr)) FFPN_FAIL(ctx, "conv_fwd_bn: in_scale/in_shift must both be set or both null");
  if (training && (stat_partial == nullptr || stat_rows == nullptr)) FFPN_FAIL(ctx, "conv_fwd_bn: training needs the statistics buffer");
  if (training && d->impl != 1 && d->dtype == FFPN_BF16 && !(d->impl == 0 && in_scale == nullptr && ffpn_stem_supported(d)) &&
      ffpn_tc_fwd_supported(d)) {
    const int r = ffpn_conv_fwd_ws_bn(ctx, d, x, in_scale, in_shift, in_relu, w, y, stat_partial, stat_rows, ws, ws_bytes, (cudaStream_t)stream,
                                      count, momentum, eps, gamma, beta, running_mean, running_var, scale, shift, save_mean, save_invstd);
    if (r >= 0) return r;                                  // fused: conv + finalize by the last CTA
  }
  if (ffpn_conv_fwd(ctx, d, x, in_scale, in_shift, in_relu, w, y, training ? stat_partial : nullptr, stat_rows, ws, ws_bytes, stream)) return 1;
  return ffpn_bn_finalize(ctx, stat_partial, training ? *stat_rows : 0, d->Cout, count, gamma, beta, running_mean, running_var, momentum, eps,
                          training, scale, shift, save_mean, save_invstd, stream);
}

extern "C" int ffpn_conv_dgrad(ffpn_ctx* ctx, const ffpn_conv_desc* d, const void* dy, const float* w, const void* addend,
                               void* dx, void* ws, size_t ws_bytes, void* stream) {
  if (check_desc(ctx, d, "conv_dgrad")) return 1;
  const bool tc_ok = ffpn_tc_dgrad_supported(d);
  if (d->impl == 2 && !tc_ok) FFPN_FAIL(ctx, "conv_dgrad: tcgen05 kernel does not support this geometry");
  if (tc_ok && d->impl != 1) {
    const int r = ffpn_conv_fwd_tc(ctx, d, true, dy, nullptr, nullptr, 0, w, addend, dx, nullptr, nullptr, ws, ws_bytes, (cudaStream_t)stream);
    if (r >= 0) return r;
    if (d->impl == 2) FFPN_FAIL(ctx, "conv_dgrad: the tcgen05 kernel declined this call (workspace too small)");
  }
  ctx->routes[FFPN_ROUTE_SIMT]++;
  if (d->dtype == FFPN_BF16) ffpn_log_route("conv_dgrad -> CUDA-core kernel", d);
  return ffpn_conv_dgrad_simt(ctx, d, dy, w, addend, dx, (cudaStream_t)stream);
}

extern "C" int ffpn_conv_dgrad_bnr(ffpn_ctx* ctx, const ffpn_conv_desc* d, const void* dy, const float* w, const void* y_prev,
                                   const float* bn_scale, const float* bn_shift, void* dx, float* partial, int* rows, void* ws,
                                   size_t ws_bytes, void* stream) {
  if (check_desc(ctx, d, "conv_dgrad_bnr")) return 1;
  if (!y_prev || !bn_scale || !bn_shift || !partial || !rows) FFPN_FAIL(ctx, "conv_dgrad_bnr: null argument");
  if (d->impl != 1 && d->dtype == FFPN_BF16 && ffpn_tc_dgrad_supported(d)) {
    const int r = ffpn_conv_dgrad_ws_bnr(ctx, d, dy, w, y_prev, bn_scale, bn_shift, dx, partial, rows, ws, ws_bytes, (cudaStream_t)stream);
    if (r >= 0) return r;                                  // one kernel: the sums come out of the dgrad epilogue
  }
  if (ffpn_conv_dgrad(ctx, d, dy, w, nullptr, dx, ws, ws_bytes, stream)) return 1;
  return ffpn_bn_bwd_reduce(ctx, d->dtype, d->B * d->S * d->W * d->H, d->Cin, dx, y_prev, bn_scale, bn_shift, 1, partial, rows, stream);
}

extern "C" int ffpn_conv_wgrad(ffpn_ctx* ctx, const ffpn_conv_desc* d, const void* x, const float* in_scale,
                               const float* in_shift, int in_relu, const void* dy, float* dw, void* ws, size_t ws_bytes,
                               void* stream) {
  if (check_desc(ctx, d, "conv_wgrad")) return 1;
  if ((in_scale == nullptr) != (in_shift == nullptr)) FFPN_FAIL(ctx, "conv_wgrad: in_scale/in_shift must both be set or both null");
  if (d->impl == 0 && in_scale == nullptr && ffpn_stem_supported(d)) {
    ctx->routes[FFPN_ROUTE_STEM]++;
    return ffpn_stem_wgrad(ctx, d, x, dy, dw, ws, ws_bytes, (cudaStream_t)stream);
  }
  const bool tc_ok = ffpn_tc_wgrad_supported(d);
  if (d->impl == 2 && !tc_ok) FFPN_FAIL(ctx, "conv_wgrad: tcgen05 kernel does not support this geometry");
  if (tc_ok && d->impl != 1) {
    const int r = ffpn_conv_wgrad_tc(ctx, d, x, in_scale, in_shift, in_relu, dy, dw, ws, ws_bytes, (cudaStream_t)stream);
    if (r >= 0) return r;
    if (d->impl == 2) FFPN_FAIL(ctx, "conv_wgrad: the tcgen05 kernel declined this call (input transform without ReLU, or workspace too small)");
  }
  if (d->dtype == FFPN_BF16) ffpn_log_route("conv_wgrad -> CUDA-core kernel", d);
  ctx->routes[FFPN_ROUTE_SIMT]++;
  return ffpn_conv_wgrad_simt(ctx, d, x, in_scale, in_shift, in_relu, dy, dw, (cudaStream_t)stream);
}
