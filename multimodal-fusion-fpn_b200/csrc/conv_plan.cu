// Geometry, weight packing and dispatch shared by the warp-specialised tcgen05 convolution kernels (conv_ws.cu: forward and
// dgrad; conv_wgrad_ws.cu: weight gradient).
//
// Formulation.  A conv is canonicalised to (slices D with kD taps) x (inner plane Y x X with kY x kX taps, X contiguous).
// Positions of the zero-padded inner plane are flattened, i = y*Xp + x, so that for an output row m the input row of tap
// (dd,dy,dx) is simply m + dd*Lr + dy*Xp + dx: every tap's A operand is the SAME shared-memory tile viewed through a
// shifted UMMA descriptor start address (ffpn_tc_make_plan derives this geometry; the kernels tile on top of it).
// Weights stay fp32 in the reference's [Cout][Cin][taps] layout and are packed to the bf16 shared-memory image
// [N-chunk][K-group][tap][8-channel chunk][cout][8] per call, or once per step for all convs of a training step into the
// caller-owned packed-weight arena (ffpn_weight_arena_*).
//
// The first-generation kernels that used to live here (planar no-swizzle tiles, cp.async staging, atomically reduced weight
// gradient) are gone: every bf16 conv of the path runs on the warp-specialised kernels or the Cin == 1 stem kernels;
// geometries those decline (positions not a multiple of the flat tile, channel counts off the 16-grid) go to the CUDA-core
// kernels of conv_simt.cu, and ffpn_route_counts() reports how many did.
#include "tc_common.cuh"

namespace {

// Pack fp32 master weights [Cout][Cin][taps] to the bf16 smem image [nchunk][kg][tap][kc][n (Npad)][8].
// transposed: the operator applied is the dgrad conv: n <-> ci, k <-> co, taps flipped.
// Kc = reduction channels of the packed operator, Nc = its output channels, Npad = channels per N-chunk
// (r, e): r = index of the 8-element group (the innermost [8] of the image), e = element within it -- the eight elements of a
// group share everything but e, so a caller that produces whole groups pays the index decomposition once
__device__ __forceinline__ float pack_value_ge(const float* __restrict__ w, uint32_t r, const int e, int Cout, int Cin, int ntaps, int Kc,
                                               int Nc, int Npad, int KG, int mode, int sH, int pH, int kH, int tmin) {
  const uint32_t nkc = (uint32_t)KG >> 3;
  const uint32_t nkg = (uint32_t)(Kc / KG);
  int n = (int)(r % (uint32_t)Npad); r /= (uint32_t)Npad;
  const int kc = (int)(r % nkc); r /= nkc;
  const int tap = (int)(r % (uint32_t)ntaps); r /= (uint32_t)ntaps;
  const int kg = (int)(r % nkg);
  n += (int)(r / nkg) * Npad;
  const int k = kg * KG + kc * 8 + e;
  float v = 0.f;
  if (n < Nc) {
    if (mode == 0) v = w[((int64_t)n * Cin + k) * ntaps + tap];                       // forward
    else if (mode == 1) v = w[((int64_t)k * Cin + n) * ntaps + (ntaps - 1 - tap)];   // dgrad, stride 1
    else if (mode == 2) {                     // dgrad of a depth-strided (1,1,kH) conv: n = r*Cin + ci
      const int rr = n / Cin, ci = n - rr * Cin;
      const int dx = rr + pH - sH * (tap + tmin);
      if (dx >= 0 && dx < kH) v = w[((int64_t)k * Cin + ci) * kH + dx];
    } else if (mode == 5 || mode == 6) {      // stride-1 conv on the pair view of input and output (ffpn_make_pair2_desc): 5 forward, 6 dgrad
      // operator channels: forward n = (ho,co), k = (hi,ci); dgrad n = (hi,ci), k = (ho,co) and flipped taps.  Cout / Cin = the real ones
      const int nn = mode == 5 ? n : k, kk = mode == 5 ? k : n;
      const int ho = nn / Cout, co = nn - ho * Cout, hi = kk / Cin, ci = kk - hi * Cin;
      const int tp = mode == 5 ? tap : ntaps - 1 - tap;
      const int t3 = tp % 3, dyw = tp / 3;
      const int dx = 2 * (t3 - 1) + hi - ho + 1;
      if (dx >= 0 && dx < 3) v = w[((int64_t)co * Cin + ci) * ntaps + dyw * 3 + dx];
    } else {                                  // pair view of the (1,1,3) s2 p1 conv: 3 forward (k = (h,ci)), 4 dgrad (n = (h,ci))
      const int c2 = mode == 3 ? k : n, oc = mode == 3 ? n : k;
      const int hh = c2 / Cin, ci = c2 - hh * Cin;
      const int tp = mode == 3 ? tap : 1 - tap;
      const int dx = tp == 0 ? (hh == 1 ? 0 : -1) : (hh == 0 ? 1 : 2);
      if (dx >= 0) v = w[((int64_t)oc * Cin + ci) * 3 + dx];
    }
  }
  return v;
}

__device__ __forceinline__ float pack_value(const float* __restrict__ w, int64_t i, int Cout, int Cin, int ntaps, int Kc, int Nc,
                                            int Npad, int KG, int mode, int sH, int pH, int kH, int tmin) {
  const uint32_t u = (uint32_t)i;             // one image has < 2^31 elements: 32-bit index arithmetic
  return pack_value_ge(w, u >> 3, (int)(u & 7u), Cout, Cin, ntaps, Kc, Nc, Npad, KG, mode, sH, pH, kH, tmin);
}

__global__ void pack_weights_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cout, int Cin, int ntaps,
                                    int Kc, int Nc, int Npad, int KG, int nchunks, int mode, int sH, int pH, int kH,
                                    int tmin) {
  pdl_prologue();
  const int64_t total = (int64_t)nchunks * ntaps * Kc * Npad;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(pack_value(w, i, Cout, Cin, ntaps, Kc, Nc, Npad, KG, mode, sH, pH, kH, tmin));
}

// All images of the packed-weight arena in one launch: a thread produces one 8-element group (16 bytes of the image: job totals,
// prefixes and destinations are all multiples of 8 elements) -> job by binary search on the prefix sums, one index decomposition
// and one 16-byte store per group instead of per element (this launch sits on the serial tail of the step, after the optimiser).
__global__ void pack_all_kernel(const ffpn_pack_job* __restrict__ jobs, int njobs, long long total, bf16* __restrict__ arena) {
  pdl_prologue();
  const long long ngroups = total >> 3;
  for (long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x; gi < ngroups; gi += (long long)gridDim.x * blockDim.x) {
    const long long i = gi << 3;
    int lo = 0, hi = njobs - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (jobs[mid].prefix <= i) lo = mid; else hi = mid - 1;
    }
    const ffpn_pack_job& j = jobs[lo];
    const long long li = i - j.prefix;
    const uint32_t r = (uint32_t)(li >> 3);
    uint32_t pk[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const float v0 = pack_value_ge(j.w, r, 2 * q, j.Cout, j.Cin, j.ntaps, j.Kc, j.Nc, j.Npad, j.KG, j.mode, j.sH, j.pH, j.kH, j.tmin);
      const float v1 = pack_value_ge(j.w, r, 2 * q + 1, j.Cout, j.Cin, j.ntaps, j.Kc, j.Nc, j.Npad, j.KG, j.mode, j.sH, j.pH, j.kH, j.tmin);
      __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
      pk[q] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(arena + j.dst + li) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

}  // namespace

Plan ffpn_tc_make_plan(const ffpn_conv_desc* d, bool transposed, int num_sms) {
  Plan pl;
  memset(&pl, 0, sizeof(pl));
  pl.ok = false;
  TcParams& p = pl.p;
  if (d->dtype != FFPN_BF16) return pl;
  if (d->sS != 1 || d->sW != 1) return pl;
  const bool strided = d->sH != 1;
  if (strided && (d->kS != 1 || d->kW != 1)) return pl;          // only the projection's depth-strided convs
  // logical conv as seen by the kernel (dgrad: roles swapped, pads k-1-p)
  int64_t S = transposed ? d->oS : d->S, W = transposed ? d->oW : d->W, H = transposed ? d->oH : d->H;
  int64_t oS = transposed ? d->S : d->oS, oW = transposed ? d->W : d->oW, oH = transposed ? d->H : d->oH;
  int pS = transposed ? d->kS - 1 - d->pS : d->pS, pW = transposed ? d->kW - 1 - d->pW : d->pW,
      pH = transposed ? d->kH - 1 - d->pH : d->pH;
  int Cin = transposed ? d->Cout : d->Cin, Cout = transposed ? d->Cin : d->Cout;
  const int64_t B = d->B;
  int kS = d->kS, kW = d->kW, kH = d->kH;
  p.sX = 1; p.packmode = transposed ? 1 : 0;
  int tmin = 0;
  if (strided && transposed) {
    // dgrad of an X-strided conv == stride-1 conv over dy whose N = sH * Cin columns are the sH interleaved
    // input positions:  dx[s*j + r] = sum_t W[dx = r + p - s*(t + tmin)]^T dy[j + t + tmin]
    const int sH = d->sH;
    if (d->H % sH != 0 || d->H / sH != d->oH) return pl;
    tmin = -((d->kH - 1 - d->pH) / sH);                          // ceil((p - (k-1)) / s) for p <= k-1
    if (d->pH > d->kH - 1) return pl;
    const int tmax = (sH - 1 + d->pH) / sH;
    kH = tmax - tmin + 1; pH = -tmin;
    Cout = sH * d->Cin;
    H = d->oH; oH = d->oH;                                       // rows of dy in, rows of (sH*Cin)-wide dx out
    p.packmode = 2;
  } else if (strided) {
    p.sX = d->sH;
  }
  if (Cin % 16 != 0 || Cout % 8 != 0 || Cin < 16) return pl;
  if (!strided && H == 1 && oH == 1 && kH == 1 && (kW > 1)) {
    // en-face / 2-D maps: (S, W) becomes the inner plane
    p.NB = 1; p.D = (int)B; p.kD = 1; p.pD = 0; p.oD = (int)B;
    p.Y = (int)S; p.kY = kS; p.pY = pS; p.oY = (int)oS;
    p.X = (int)W; p.kX = kW; p.pX = pW; p.oX = (int)oW;
    p.inD = S * W; p.inY = W; p.inNB = 0;
    p.outD = oS * oW; p.outY = oW; p.outNB = 0;
  } else if (!strided && kW == 1 && kH == 1) {
    if (kS == 1) {                                    // 1x1x1: a plain GEMM over all positions
      p.NB = 1; p.D = 1; p.kD = 1; p.pD = 0; p.oD = 1;
      p.Y = 1; p.kY = 1; p.pY = 0; p.oY = 1;
      const int64_t P = B * S * W * H;
      if (P >= (1ll << 31)) return pl;
      p.X = (int)P; p.kX = 1; p.pX = 0; p.oX = (int)P;
      p.inD = p.inY = p.inNB = p.outD = p.outY = p.outNB = 0;
    } else {                                          // taps across slices only
      p.NB = (int)B; p.D = (int)S; p.kD = kS; p.pD = pS; p.oD = (int)oS;
      p.Y = 1; p.kY = 1; p.pY = 0; p.oY = 1;
      p.X = (int)(W * H); p.kX = 1; p.pX = 0; p.oX = (int)(W * H);
      p.inNB = S * W * H; p.inD = W * H; p.inY = 0;
      p.outNB = oS * W * H; p.outD = W * H; p.outY = 0;
    }
  } else if (kS == 1) {                               // taps inside the (W, H) plane
    p.NB = 1; p.D = (int)(B * S); p.kD = 1; p.pD = 0; p.oD = (int)(B * S);
    p.Y = (int)W; p.kY = kW; p.pY = pW; p.oY = (int)oW;
    p.X = (int)H; p.kX = kH; p.pX = pH; p.oX = (int)oH;
    p.inD = W * H; p.inY = H; p.inNB = 0;
    p.outD = oW * oH; p.outY = oH; p.outNB = 0;
  } else {
    return pl;
  }
  p.Cin = Cin; p.Cout = Cout;
  {                                                   // N-chunks of at most 256 output channels (blockIdx.y)
    const int npad = (Cout + 15) & ~15;
    pl.nchunks = (npad + 255) / 256;
    p.Npad = (((npad + pl.nchunks - 1) / pl.nchunks) + 15) & ~15;
  }
  // slots per X line: tap dx reads input x = sX*(ox + q) + res, q = floor((dx - pX) / sX); hl = max(-q), hr = max(q) + hl
  int qmin = 0, qmax = 0, nres = 0;
  bool seen[16] = {false};
  if (p.sX > 16) return pl;
  for (int dx = 0; dx < p.kX; dx++) {
    const int e = dx - p.pX;
    const int q = e >= 0 ? e / p.sX : -((-e + p.sX - 1) / p.sX);
    const int res = e - q * p.sX;
    if (dx == 0 || q < qmin) qmin = q;
    if (dx == 0 || q > qmax) qmax = q;
    if (!seen[res]) { seen[res] = true; nres++; }
  }
  p.hl = qmin < 0 ? -qmin : 0;
  if (qmin > 0) return pl;
  p.nsets = nres > 1 ? p.sX : 1;
  if (p.nsets > 2) return pl;
  const int hr = qmax + p.hl;
  p.Xp = p.oX + hr;
  p.Qout = (p.oY - 1) * p.Xp + p.oX;
  // Geometry only: the kernels tile on top of it (make_ws_plan, make_wgrad_ws_plan) and decide themselves what they take.
  if (p.kD * p.kY * p.kX > 27) return pl;
  pl.ok = true;
  return pl;
}



bool ffpn_wgrad_ws_supported(const ffpn_conv_desc* d);
size_t ffpn_wgrad_ws_workspace_bytes(const ffpn_conv_desc* d);
int ffpn_conv_wgrad_ws(ffpn_ctx*, const ffpn_conv_desc*, const void*, const float*, const float*, int, const void*, float*, void*, size_t,
                       cudaStream_t);
int ffpn_conv_fwd_ws(ffpn_ctx*, const ffpn_conv_desc*, bool transposed, const void*, const float*, const float*, int, const float*,
                     const void* addend, void*, float*, int*, void*, size_t, cudaStream_t);
bool ffpn_conv_ws_supported(const ffpn_conv_desc* d, bool transposed, bool has_aff, bool relu);

bool ffpn_tc_wgrad_supported(const ffpn_conv_desc* d) { return ffpn_wgrad_ws_supported(d); }
bool ffpn_tc_fwd_supported(const ffpn_conv_desc* d) { return ffpn_conv_ws_supported(d, false, false, false); }
bool ffpn_tc_dgrad_supported(const ffpn_conv_desc* d) { return ffpn_conv_ws_supported(d, true, false, false); }

size_t ffpn_tc_workspace_bytes(const ffpn_conv_desc* d) {
  const size_t taps = (size_t)d->kS * d->kW * d->kH;
  const size_t cin = (d->Cin + 63) & ~63, cout = (d->Cout + 63) & ~63;
  const size_t pack = taps * cin * cout * 2 + 65536;                  // packed weights | wgrad partial tiles
  const size_t wg = ffpn_wgrad_ws_workspace_bytes(d);
  return pack > wg ? pack : wg;
}

// Returns 0 = launched, 1 = error (message set), -1 = geometry not handled by the tcgen05 kernels (the caller uses conv_simt.cu).
int ffpn_conv_wgrad_tc(ffpn_ctx* ctx, const ffpn_conv_desc* d, const void* x, const float* in_scale, const float* in_shift,
                       int in_relu, const void* dy, float* dw, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int r = ffpn_conv_wgrad_ws(ctx, d, x, in_scale, in_shift, in_relu, dy, dw, ws, ws_bytes, st);
  if (r >= 0) ctx->routes[FFPN_ROUTE_WS]++;
  return r;
}

int ffpn_conv_fwd_tc(ffpn_ctx* ctx, const ffpn_conv_desc* d, bool transposed, const void* x, const float* in_scale,
                     const float* in_shift, int in_relu, const float* w, const void* addend, void* y, float* stat_partial,
                     int* stat_rows, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int r = ffpn_conv_fwd_ws(ctx, d, transposed, x, in_scale, in_shift, in_relu, w, addend, y, stat_partial, stat_rows, ws, ws_bytes, st);
  if (r >= 0) ctx->routes[FFPN_ROUTE_WS]++;
  return r;
}

const void* ffpn_tc_pack_weights(ffpn_ctx* ctx, const float* w, void* ws, const ffpn_conv_desc* d, const TcParams& p, int nchunks,
                                 int KG, cudaStream_t st, bool* launched) {
  const int ntaps = p.kD * p.kY * p.kX;
  const int64_t total = (int64_t)nchunks * ntaps * p.Cin * p.Npad;
  *launched = false;
  ffpn_pack_job j;
  j.w = w; j.total = total; j.Cout = d->Cout; j.Cin = d->Cin; j.ntaps = ntaps; j.Kc = p.Cin; j.Nc = p.Cout; j.Npad = p.Npad; j.KG = KG;
  j.nchunks = nchunks; j.mode = p.packmode; j.sH = d->sH; j.pH = d->pH; j.kH = d->kH; j.tmin = -p.pX;
  if (ctx->arena_state == 2 && ctx->arena_lookup) {
    for (int i = 0; i < ctx->njobs; i++) {
      const ffpn_pack_job& q = ctx->jobs[i];
      if (q.w == w && q.mode == j.mode && q.KG == KG && q.Npad == j.Npad && q.nchunks == nchunks && q.total == total && q.tmin == j.tmin &&
          q.Nc == j.Nc && q.Kc == j.Kc)
        return ctx->arena + (size_t)q.dst * 2;                       // image regenerated by ffpn_weight_arena_pack this step
    }
  }
  void* out = ws;
  if (ctx->arena_state == 1 && ctx->arena_lookup && ctx->njobs < FFPN_MAX_PACK_JOBS && ctx->arena_used + (size_t)total * 2 + 1024 <= ctx->arena_bytes) {
    j.dst = (long long)(ctx->arena_used / 2);
    j.prefix = ctx->arena_elems;
    ctx->jobs[ctx->njobs++] = j;
    out = ctx->arena + ctx->arena_used;
    ctx->arena_used += ((size_t)total * 2 + 1023) & ~(size_t)1023;   // cp.async.bulk sources stay 16-byte aligned
    ctx->arena_elems += total;
  }
  const int g = (int)((total + 255) / 256 < 1024 ? (total + 255) / 256 : 1024);
  ffpn_launch(pack_weights_kernel, g, 256, 0, st, w, (bf16*)out, j.Cout, j.Cin, ntaps, j.Kc, j.Nc, j.Npad, KG, nchunks, j.mode, j.sH, j.pH, j.kH,
                                         j.tmin);
  *launched = true;
  return out;
}

// ---- packed-weight arena (include/ffpn.h) -----------------------------------------------------------------
extern "C" int ffpn_weight_arena_begin(ffpn_ctx* ctx, void* arena, size_t bytes) {
  if (!ctx) return 1;
  if (arena == nullptr || bytes < (1u << 20) || ((uintptr_t)arena & 1023)) FFPN_FAIL(ctx, "weight_arena_begin: need a 1 KiB-aligned buffer of >= 1 MiB");
  ctx->arena = (char*)arena; ctx->arena_bytes = bytes; ctx->arena_used = 0; ctx->njobs = 0; ctx->arena_elems = 0;
  ctx->arena_state = 1;
  ctx->arena_lookup = 1;
  return 0;
}
extern "C" int ffpn_weight_arena_enable(ffpn_ctx* ctx, int on) {
  if (!ctx) return 1;
  ctx->arena_lookup = on ? 1 : 0;
  return 0;
}
extern "C" int ffpn_weight_arena_seal(ffpn_ctx* ctx) {
  if (!ctx) return 1;
  if (ctx->arena_state != 1) FFPN_FAIL(ctx, "weight_arena_seal: not recording");
  if (ctx->d_jobs == nullptr && cudaMalloc(&ctx->d_jobs, sizeof(ffpn_pack_job) * FFPN_MAX_PACK_JOBS) != cudaSuccess)
    FFPN_FAIL(ctx, "weight_arena_seal: cannot allocate the job table");
  if (ctx->njobs > 0 && cudaMemcpy(ctx->d_jobs, ctx->jobs, sizeof(ffpn_pack_job) * ctx->njobs, cudaMemcpyHostToDevice) != cudaSuccess)
    FFPN_FAIL(ctx, "weight_arena_seal: cannot upload the job table");
  ctx->arena_state = ctx->njobs > 0 ? 2 : 0;
  return 0;
}
extern "C" int ffpn_weight_arena_pack(ffpn_ctx* ctx, void* stream) {
  if (!ctx) return 1;
  if (ctx->arena_state != 2) return 0;
  const long long total = ctx->arena_elems;
  if (total & 7) FFPN_FAIL(ctx, "weight_arena_pack: an image is not a whole number of 8-element groups");
  const long long ngroups = total >> 3;
  const int g = (int)((ngroups + 255) / 256 < 148 * 16 ? (ngroups + 255) / 256 : 148 * 16);
  ffpn_launch(pack_all_kernel, g, 256, 0, (cudaStream_t)stream, ctx->d_jobs, ctx->njobs, total, (bf16*)ctx->arena);
  FFPN_CHECK_LAUNCH(ctx, "weight_arena_pack");
  return 0;
}
extern "C" int ffpn_weight_arena_end(ffpn_ctx* ctx) {
  if (!ctx) return 1;
  ctx->arena_state = 0; ctx->njobs = 0; ctx->arena = nullptr; ctx->arena_bytes = ctx->arena_used = 0; ctx->arena_elems = 0;
  ctx->arena_lookup = 0;
  return 0;
}
