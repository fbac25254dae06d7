// Generic direct convolution on CUDA cores (fp32 accumulate): forward, dgrad and wgrad for every kernel
// shape on the path, any channel count, fp32 or bf16 activations.  This is the shape-complete kernel
// set: it serves the Cin=1 stems, odd shapes and the fp32 "exact" mode, and it is the on-device
// cross-check for the tcgen05 implicit-GEMM kernels in conv_tc.cu.
#include "common.cuh"

namespace {

constexpr int MAXTAPS = 9;     // largest kernel on the path: (1,3,3) / (3,3,1)
constexpr int CO_T = 16;       // output channels per thread
constexpr int CI_CHUNK = 32;   // reduction channels staged in smem per step
constexpr int CONV_THREADS = 128;

struct Geom {
  int B, S, W, H, oS, oW, oH, Cin, Cout, kS, kW, kH, sS, sW, sH, pS, pW, pH;
};

__host__ Geom make_geom(const ffpn_conv_desc* d) {
  Geom g;
  g.B = (int)d->B; g.S = (int)d->S; g.W = (int)d->W; g.H = (int)d->H;
  g.oS = (int)d->oS; g.oW = (int)d->oW; g.oH = (int)d->oH;
  g.Cin = d->Cin; g.Cout = d->Cout;
  g.kS = d->kS; g.kW = d->kW; g.kH = d->kH;
  g.sS = d->sS; g.sW = d->sW; g.sH = d->sH;
  g.pS = d->pS; g.pW = d->pW; g.pH = d->pH;
  return g;
}

// TRANSPOSED == false: forward.   dst = output position, src = input position = dst*stride - pad + tap
// TRANSPOSED == true : dgrad.     dst = input position,  src = output position = (dst + pad - tap)/stride
template <typename T, bool TRANSPOSED>
__global__ void __launch_bounds__(CONV_THREADS)
conv_simt_kernel(Geom g, const T* __restrict__ x, const float* __restrict__ in_scale,
                 const float* __restrict__ in_shift, int in_relu, const float* __restrict__ w,
                 const T* __restrict__ addend, T* __restrict__ y, float* __restrict__ stat_partial) {
  pdl_prologue();
  extern __shared__ float smem[];
  const int ntaps = g.kS * g.kW * g.kH;
  const int KIN = TRANSPOSED ? g.Cout : g.Cin;
  const int KOUT = TRANSPOSED ? g.Cin : g.Cout;
  float* w_s = smem;                                   // [ntaps][CI_CHUNK][CO_T]
  float* sc_s = w_s + ntaps * CI_CHUNK * CO_T;         // [CI_CHUNK]
  float* sh_s = sc_s + CI_CHUNK;                       // [CI_CHUNK]
  const int dS = TRANSPOSED ? g.S : g.oS, dW = TRANSPOSED ? g.W : g.oW, dH = TRANSPOSED ? g.H : g.oH;
  const int qS = TRANSPOSED ? g.oS : g.S, qW = TRANSPOSED ? g.oW : g.W, qH = TRANSPOSED ? g.oH : g.H;
  const int64_t P = (int64_t)g.B * dS * dW * dH;
  const int64_t ntiles = (P + CONV_THREADS - 1) / CONV_THREADS;
  const int co0 = blockIdx.y * CO_T;
  constexpr int VEC = Elem<T>::VEC;
  const bool vec_in = (KIN % VEC) == 0;
  const bool vec_out = (KOUT % VEC) == 0;
  const bool has_aff = (!TRANSPOSED) && in_scale != nullptr;

  float ssum[CO_T], ssq[CO_T];
#pragma unroll
  for (int i = 0; i < CO_T; i++) { ssum[i] = 0.f; ssq[i] = 0.f; }

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t pos = tile * CONV_THREADS + threadIdx.x;
    const bool active = pos < P;
    int srcpos[MAXTAPS];
    {
      int64_t r = active ? pos : 0;
      const int h = (int)(r % dH); r /= dH;
      const int ww = (int)(r % dW); r /= dW;
      const int s = (int)(r % dS);
      const int b = (int)(r / dS);
#pragma unroll
      for (int t = 0; t < MAXTAPS; t++) {
        int sp = -1;
        if (active && t < ntaps) {
          const int th = t % g.kH, tw = (t / g.kH) % g.kW, ts = t / (g.kH * g.kW);
          int qs, qw, qh;
          bool ok;
          if (!TRANSPOSED) {
            qs = s * g.sS - g.pS + ts; qw = ww * g.sW - g.pW + tw; qh = h * g.sH - g.pH + th;
            ok = true;
          } else {
            const int ns = s + g.pS - ts, nw = ww + g.pW - tw, nh = h + g.pH - th;
            ok = (ns >= 0) && (nw >= 0) && (nh >= 0) && (ns % g.sS == 0) && (nw % g.sW == 0) && (nh % g.sH == 0);
            qs = ns / g.sS; qw = nw / g.sW; qh = nh / g.sH;
          }
          ok = ok && qs >= 0 && qs < qS && qw >= 0 && qw < qW && qh >= 0 && qh < qH;
          if (ok) sp = ((b * qS + qs) * qW + qw) * qH + qh;
        }
        srcpos[t] = sp;
      }
    }
    float acc[CO_T];
#pragma unroll
    for (int i = 0; i < CO_T; i++) acc[i] = 0.f;

    for (int c0 = 0; c0 < KIN; c0 += CI_CHUNK) {
      const int cc = min(CI_CHUNK, KIN - c0);
      __syncthreads();
      for (int i = threadIdx.x; i < ntaps * cc * CO_T; i += CONV_THREADS) {
        const int co = i % CO_T, ci = (i / CO_T) % cc, t = i / (CO_T * cc);
        float v = 0.f;
        if (co0 + co < KOUT) {
          // master layout [Cout][Cin][tap]
          const int64_t wi = TRANSPOSED ? ((int64_t)(c0 + ci) * g.Cin + (co0 + co)) * ntaps + t
                                        : ((int64_t)(co0 + co) * g.Cin + (c0 + ci)) * ntaps + t;
          v = w[wi];
        }
        w_s[(t * CI_CHUNK + ci) * CO_T + co] = v;
      }
      if (has_aff) {
        for (int i = threadIdx.x; i < cc; i += CONV_THREADS) { sc_s[i] = in_scale[c0 + i]; sh_s[i] = in_shift[c0 + i]; }
      }
      __syncthreads();
#pragma unroll
      for (int t = 0; t < MAXTAPS; t++) {
        if (t >= ntaps) break;
        const int sp = srcpos[t];
        if (sp < 0) continue;
        const T* xp = x + (int64_t)sp * KIN + c0;
        const float* wt = w_s + t * CI_CHUNK * CO_T;
        if (vec_in) {
          for (int cv = 0; cv < cc; cv += VEC) {
            float v[VEC];
            Elem<T>::load(xp + cv, v);
#pragma unroll
            for (int j = 0; j < VEC; j++) {
              float xv = v[j];
              if (has_aff) { xv = fmaf(xv, sc_s[cv + j], sh_s[cv + j]); if (in_relu) xv = fmaxf(xv, 0.f); }
              const float4* wr = reinterpret_cast<const float4*>(wt + (cv + j) * CO_T);
#pragma unroll
              for (int q = 0; q < CO_T / 4; q++) {
                const float4 wv = wr[q];
                acc[4 * q + 0] = fmaf(xv, wv.x, acc[4 * q + 0]);
                acc[4 * q + 1] = fmaf(xv, wv.y, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(xv, wv.z, acc[4 * q + 2]);
                acc[4 * q + 3] = fmaf(xv, wv.w, acc[4 * q + 3]);
              }
            }
          }
        } else {
          for (int ci = 0; ci < cc; ci++) {
            float xv = Elem<T>::ld1(xp + ci);
            if (has_aff) { xv = fmaf(xv, sc_s[ci], sh_s[ci]); if (in_relu) xv = fmaxf(xv, 0.f); }
            const float4* wr = reinterpret_cast<const float4*>(wt + ci * CO_T);
#pragma unroll
            for (int q = 0; q < CO_T / 4; q++) {
              const float4 wv = wr[q];
              acc[4 * q + 0] = fmaf(xv, wv.x, acc[4 * q + 0]);
              acc[4 * q + 1] = fmaf(xv, wv.y, acc[4 * q + 1]);
              acc[4 * q + 2] = fmaf(xv, wv.z, acc[4 * q + 2]);
              acc[4 * q + 3] = fmaf(xv, wv.w, acc[4 * q + 3]);
            }
          }
        }
      }
    }
    if (active) {
      T* yp = y + pos * KOUT + co0;
      if (addend != nullptr) {
#pragma unroll
        for (int i = 0; i < CO_T; i++)
          if (co0 + i < KOUT) acc[i] += Elem<T>::ld1(addend + pos * KOUT + co0 + i);
      }
#pragma unroll
      for (int i = 0; i < CO_T; i++) acc[i] = Elem<T>::rnd(acc[i]);
      if (vec_out && co0 + CO_T <= KOUT) {
#pragma unroll
        for (int q = 0; q < CO_T / VEC; q++) {
          float v[VEC];
#pragma unroll
          for (int j = 0; j < VEC; j++) v[j] = acc[q * VEC + j];
          Elem<T>::store(yp + q * VEC, v);
        }
      } else {
#pragma unroll
        for (int i = 0; i < CO_T; i++)
          if (co0 + i < KOUT) Elem<T>::st1(yp + i, acc[i]);
      }
#pragma unroll
      for (int i = 0; i < CO_T; i++) { ssum[i] += acc[i]; ssq[i] = fmaf(acc[i], acc[i], ssq[i]); }
    }
  }
  if (stat_partial != nullptr) {
    __syncthreads();
    float* red = smem;  // [warps][2*CO_T]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < CO_T; i++) {
      const float a = warp_sum(ssum[i]), b = warp_sum(ssq[i]);
      if (lane == 0) { red[wid * 2 * CO_T + i] = a; red[wid * 2 * CO_T + CO_T + i] = b; }
    }
    __syncthreads();
    if (threadIdx.x < 2 * CO_T) {
      float v = 0.f;
      for (int k = 0; k < CONV_THREADS / 32; k++) v += red[k * 2 * CO_T + threadIdx.x];
      const int which = threadIdx.x / CO_T, co = co0 + threadIdx.x % CO_T;
      if (co < KOUT) stat_partial[((int64_t)blockIdx.x * 2 + which) * KOUT + co] = v;
    }
  }
}

// ---- wgrad -------------------------------------------------------------------------------------------
constexpr int WG_PT = 32;  // positions staged per step

template <typename T, int KI_B, int KO_B>
__global__ void __launch_bounds__((KI_B / 2) * (KO_B / 2))
conv_wgrad_simt_kernel(Geom g, const T* __restrict__ x, const float* __restrict__ in_scale,
                       const float* __restrict__ in_shift, int in_relu, const T* __restrict__ dy,
                       float* __restrict__ dw, int n_ki_tiles) {
  pdl_prologue();
  constexpr int NT = (KI_B / 2) * (KO_B / 2);
  extern __shared__ float smem[];
  const int ntaps = g.kS * g.kW * g.kH;
  float* dy_s = smem;                          // [WG_PT][KO_B]
  float* x_s = dy_s + WG_PT * KO_B;            // [ntaps][WG_PT][KI_B]
  int* sp_s = reinterpret_cast<int*>(x_s + ntaps * WG_PT * KI_B);  // [ntaps][WG_PT]
  const int ki0 = (blockIdx.y % n_ki_tiles) * KI_B, ko0 = (blockIdx.y / n_ki_tiles) * KO_B;
  const int kit = threadIdx.x % (KI_B / 2), kot = threadIdx.x / (KI_B / 2);
  const int64_t P = (int64_t)g.B * g.oS * g.oW * g.oH;
  const int64_t nsteps = (P + WG_PT - 1) / WG_PT;
  const bool has_aff = in_scale != nullptr;

  float acc[MAXTAPS][4];
#pragma unroll
  for (int t = 0; t < MAXTAPS; t++) { acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f; }

  for (int64_t step = blockIdx.x; step < nsteps; step += gridDim.x) {
    const int64_t p0 = step * WG_PT;
    __syncthreads();
    for (int i = threadIdx.x; i < ntaps * WG_PT; i += NT) {
      const int p = i % WG_PT, t = i / WG_PT;
      const int64_t pos = p0 + p;
      int sp = -1;
      if (pos < P) {
        int64_t r = pos;
        const int h = (int)(r % g.oH); r /= g.oH;
        const int ww = (int)(r % g.oW); r /= g.oW;
        const int s = (int)(r % g.oS);
        const int b = (int)(r / g.oS);
        const int th = t % g.kH, tw = (t / g.kH) % g.kW, ts = t / (g.kH * g.kW);
        const int qs = s * g.sS - g.pS + ts, qw = ww * g.sW - g.pW + tw, qh = h * g.sH - g.pH + th;
        if (qs >= 0 && qs < g.S && qw >= 0 && qw < g.W && qh >= 0 && qh < g.H)
          sp = ((b * g.S + qs) * g.W + qw) * g.H + qh;
      }
      sp_s[i] = sp;
    }
    for (int i = threadIdx.x; i < WG_PT * KO_B; i += NT) {
      const int ko = i % KO_B, p = i / KO_B;
      float v = 0.f;
      if (p0 + p < P && ko0 + ko < g.Cout) v = Elem<T>::ld1(dy + (p0 + p) * g.Cout + ko0 + ko);
      dy_s[i] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ntaps * WG_PT * KI_B; i += NT) {
      const int ki = i % KI_B, tp = i / KI_B;
      const int sp = sp_s[tp];
      float v = 0.f;
      if (sp >= 0 && ki0 + ki < g.Cin) {
        v = Elem<T>::ld1(x + (int64_t)sp * g.Cin + ki0 + ki);
        if (has_aff) { v = fmaf(v, in_scale[ki0 + ki], in_shift[ki0 + ki]); if (in_relu) v = fmaxf(v, 0.f); }
      }
      x_s[i] = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int p = 0; p < WG_PT; p++) {
      const float2 d2 = *reinterpret_cast<const float2*>(dy_s + p * KO_B + kot * 2);
#pragma unroll
      for (int t = 0; t < MAXTAPS; t++) {
        if (t >= ntaps) break;
        const float2 x2 = *reinterpret_cast<const float2*>(x_s + (t * WG_PT + p) * KI_B + kit * 2);
        acc[t][0] = fmaf(d2.x, x2.x, acc[t][0]);
        acc[t][1] = fmaf(d2.x, x2.y, acc[t][1]);
        acc[t][2] = fmaf(d2.y, x2.x, acc[t][2]);
        acc[t][3] = fmaf(d2.y, x2.y, acc[t][3]);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < MAXTAPS; t++) {
    if (t >= ntaps) break;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int ko = ko0 + kot * 2 + (q >> 1), ki = ki0 + kit * 2 + (q & 1);
      if (ko < g.Cout && ki < g.Cin) atomicAdd(dw + ((int64_t)ko * g.Cin + ki) * ntaps + t, acc[t][q]);
    }
  }
}

template <typename T>
int launch_conv(ffpn_ctx* ctx, const ffpn_conv_desc* d, bool transposed, const void* x, const float* in_scale,
                const float* in_shift, int in_relu, const float* w, const void* addend, void* y, float* stat_partial, int* stat_rows,
                cudaStream_t st) {
  Geom g = make_geom(d);
  const int ntaps = g.kS * g.kW * g.kH;
  if (ntaps > MAXTAPS) FFPN_FAIL(ctx, "conv: %d taps unsupported (max %d)", ntaps, MAXTAPS);
  const int64_t P = transposed ? (int64_t)g.B * g.S * g.W * g.H : (int64_t)g.B * g.oS * g.oW * g.oH;
  if (P * (int64_t)max(g.Cin, g.Cout) <= 0 || P >= (1ll << 31)) FFPN_FAIL(ctx, "conv: bad position count");
  const int kout = transposed ? g.Cin : g.Cout;
  const int chunks = (kout + CO_T - 1) / CO_T;
  const int64_t ntiles = (P + CONV_THREADS - 1) / CONV_THREADS;
  int gx = (int)(ntiles < (int64_t)FFPN_STAT_ROWS ? ntiles : (int64_t)FFPN_STAT_ROWS);
  dim3 grid(gx, chunks);
  size_t smem = (size_t)(ntaps * CI_CHUNK * CO_T + 2 * CI_CHUNK) * sizeof(float);
  smem = max(smem, (size_t)(CONV_THREADS / 32) * 2 * CO_T * sizeof(float));
  if (transposed)
    ffpn_launch(conv_simt_kernel<T, true>, grid, CONV_THREADS, smem, st, g, (const T*)x, nullptr, nullptr, 0, w, (const T*)addend, (T*)y, nullptr);
  else
    ffpn_launch(conv_simt_kernel<T, false>, grid, CONV_THREADS, smem, st, g, (const T*)x, in_scale, in_shift, in_relu, w, (const T*)addend,
                                                               (T*)y, stat_partial);
  FFPN_CHECK_LAUNCH(ctx, transposed ? "conv_dgrad_simt" : "conv_fwd_simt");
  if (stat_rows) *stat_rows = gx;
  return 0;
}

template <typename T>
int launch_wgrad(ffpn_ctx* ctx, const ffpn_conv_desc* d, const void* x, const float* in_scale, const float* in_shift,
                 int in_relu, const void* dy, float* dw, cudaStream_t st) {
  Geom g = make_geom(d);
  const int ntaps = g.kS * g.kW * g.kH;
  if (ntaps > MAXTAPS) FFPN_FAIL(ctx, "wgrad: %d taps unsupported", ntaps);
  const int64_t P = (int64_t)g.B * g.oS * g.oW * g.oH;
  const int64_t nsteps = (P + WG_PT - 1) / WG_PT;
  const bool small = (g.Cin <= 16 && g.Cout <= 16);
  const int kib = small ? 16 : 32, kob = small ? 16 : 32;
  const int nki = (g.Cin + kib - 1) / kib, nko = (g.Cout + kob - 1) / kob;
  int64_t want = (int64_t)ctx->num_sms * 8 / ((int64_t)nki * nko);
  if (want < 1) want = 1;
  const int gx = (int)(nsteps < want ? nsteps : want);
  dim3 grid(gx, nki * nko);
  const size_t smem = (size_t)(WG_PT * kob + ntaps * WG_PT * kib) * sizeof(float) + (size_t)ntaps * WG_PT * sizeof(int);
  if (small)
    ffpn_launch(conv_wgrad_simt_kernel<T, 16, 16>, grid, 64, smem, st, g, (const T*)x, in_scale, in_shift, in_relu, (const T*)dy, dw, nki);
  else
    ffpn_launch(conv_wgrad_simt_kernel<T, 32, 32>, grid, 256, smem, st, g, (const T*)x, in_scale, in_shift, in_relu, (const T*)dy, dw, nki);
  FFPN_CHECK_LAUNCH(ctx, "conv_wgrad_simt");
  return 0;
}

}  // namespace

int ffpn_conv_fwd_simt(ffpn_ctx* ctx, const ffpn_conv_desc* d, const void* x, const float* in_scale,
                       const float* in_shift, int in_relu, const float* w, void* y, float* stat_partial,
                       int* stat_rows, cudaStream_t st) {
  if (d->dtype == FFPN_F32) return launch_conv<float>(ctx, d, false, x, in_scale, in_shift, in_relu, w, nullptr, y, stat_partial, stat_rows, st);
  return launch_conv<bf16>(ctx, d, false, x, in_scale, in_shift, in_relu, w, nullptr, y, stat_partial, stat_rows, st);
}
int ffpn_conv_dgrad_simt(ffpn_ctx* ctx, const ffpn_conv_desc* d, const void* dy, const float* w, const void* addend, void* dx, cudaStream_t st) {
  if (d->dtype == FFPN_F32) return launch_conv<float>(ctx, d, true, dy, nullptr, nullptr, 0, w, addend, dx, nullptr, nullptr, st);
  return launch_conv<bf16>(ctx, d, true, dy, nullptr, nullptr, 0, w, addend, dx, nullptr, nullptr, st);
}
int ffpn_conv_wgrad_simt(ffpn_ctx* ctx, const ffpn_conv_desc* d, const void* x, const float* in_scale,
                         const float* in_shift, int in_relu, const void* dy, float* dw, cudaStream_t st) {
  if (d->dtype == FFPN_F32) return launch_wgrad<float>(ctx, d, x, in_scale, in_shift, in_relu, dy, dw, st);
  return launch_wgrad<bf16>(ctx, d, x, in_scale, in_shift, in_relu, dy, dw, st);
}
