// Stem convolutions (Cin = 1 -> 16 channels): the first conv of each encoder and its 1x1x1 shortcut.  They are
// pure streaming kernels (read 1 channel, write 16): one thread per output position along the contiguous axis,
// no integer division per element, 16-byte stores, BatchNorm partial sums in registers.  wgrad keeps the
// 16 x taps accumulators in registers and reduces once per block.
#include "common.cuh"

namespace {

constexpr int ST_THREADS = 128;
constexpr int ST_C = 16;
constexpr int ST_MAXTAPS = 9;

struct StemGeom {
  int B, S, W, H, oS, oW, oH, kS, kW, kH, pS, pW, pH;
};

template <typename T>
__global__ void __launch_bounds__(ST_THREADS)
stem_fwd_kernel(StemGeom g, const T* __restrict__ x, const float* __restrict__ w, T* __restrict__ y,
                float* __restrict__ stat_partial) {
  pdl_prologue();
  __shared__ float w_s[ST_MAXTAPS * ST_C];
  __shared__ float red[(ST_THREADS / 32) * 2 * ST_C];
  const int ntaps = g.kS * g.kW * g.kH;
  for (int i = threadIdx.x; i < ntaps * ST_C; i += ST_THREADS) w_s[i] = w[(i % ST_C) * ntaps + i / ST_C];   // [tap][co]
  __syncthreads();
  float ssum[ST_C], ssq[ST_C];
#pragma unroll
  for (int i = 0; i < ST_C; i++) { ssum[i] = 0.f; ssq[i] = 0.f; }
  const int64_t nlines = (int64_t)g.B * g.oS * g.oW;
  for (int64_t line = blockIdx.x; line < nlines; line += gridDim.x) {
    const int ow = (int)(line % g.oW);
    const int os = (int)((line / g.oW) % g.oS);
    const int b = (int)(line / ((int64_t)g.oW * g.oS));
    for (int oh = threadIdx.x; oh < g.oH; oh += ST_THREADS) {
      float acc[ST_C];
#pragma unroll
      for (int i = 0; i < ST_C; i++) acc[i] = 0.f;
      int tap = 0;
      for (int ts = 0; ts < g.kS; ts++) {
        const int s = os - g.pS + ts;
        for (int tw = 0; tw < g.kW; tw++) {
          const int ww = ow - g.pW + tw;
          const bool ok_sw = s >= 0 && s < g.S && ww >= 0 && ww < g.W;
          const T* xl = x + (((int64_t)b * g.S + s) * g.W + ww) * g.H;
          for (int th = 0; th < g.kH; th++, tap++) {
            const int h = oh - g.pH + th;
            if (ok_sw && h >= 0 && h < g.H) {
              const float xv = Elem<T>::ld1(xl + h);
              const float4* wr = reinterpret_cast<const float4*>(w_s + tap * ST_C);
#pragma unroll
              for (int q = 0; q < ST_C / 4; q++) {
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
      T* yp = y + (line * g.oH + oh) * ST_C;
#pragma unroll
      for (int i = 0; i < ST_C; i++) acc[i] = Elem<T>::rnd(acc[i]);
      constexpr int VEC = Elem<T>::VEC;
#pragma unroll
      for (int q = 0; q < ST_C / VEC; q++) {
        float v[VEC];
#pragma unroll
        for (int j = 0; j < VEC; j++) v[j] = acc[q * VEC + j];
        Elem<T>::store(yp + q * VEC, v);
      }
#pragma unroll
      for (int i = 0; i < ST_C; i++) { ssum[i] += acc[i]; ssq[i] = fmaf(acc[i], acc[i], ssq[i]); }
    }
  }
  if (stat_partial != nullptr) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < ST_C; i++) {
      const float a = warp_sum(ssum[i]), b2 = warp_sum(ssq[i]);
      if (lane == 0) { red[wid * 2 * ST_C + i] = a; red[wid * 2 * ST_C + ST_C + i] = b2; }
    }
    __syncthreads();
    if (threadIdx.x < 2 * ST_C) {
      float v = 0.f;
      for (int k = 0; k < ST_THREADS / 32; k++) v += red[k * 2 * ST_C + threadIdx.x];
      stat_partial[(int64_t)blockIdx.x * 2 * ST_C + threadIdx.x] = v;     // [row][2][16]
    }
  }
}

template <typename T, int NT>
__global__ void __launch_bounds__(ST_THREADS)
stem_wgrad_kernel(StemGeom g, const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dw, float* __restrict__ part) {
  pdl_prologue();
  __shared__ float red[(ST_THREADS / 32) * NT * ST_C];
  float acc[NT][ST_C];
#pragma unroll
  for (int t = 0; t < NT; t++)
#pragma unroll
    for (int i = 0; i < ST_C; i++) acc[t][i] = 0.f;
  const int64_t nlines = (int64_t)g.B * g.oS * g.oW;
  constexpr int VEC = Elem<T>::VEC;
  for (int64_t line = blockIdx.x; line < nlines; line += gridDim.x) {
    const int ow = (int)(line % g.oW);
    const int os = (int)((line / g.oW) % g.oS);
    const int b = (int)(line / ((int64_t)g.oW * g.oS));
    for (int oh = threadIdx.x; oh < g.oH; oh += ST_THREADS) {
      float d[ST_C];
      const T* dp = dy + (line * g.oH + oh) * ST_C;
#pragma unroll
      for (int q = 0; q < ST_C / VEC; q++) {
        float v[VEC];
        Elem<T>::load(dp + q * VEC, v);
#pragma unroll
        for (int j = 0; j < VEC; j++) d[q * VEC + j] = v[j];
      }
      int tap = 0;
#pragma unroll
      for (int ts = 0; ts < 3; ts++) {
        if (ts >= g.kS) break;
        const int s = os - g.pS + ts;
#pragma unroll
        for (int tw = 0; tw < 3; tw++) {
          if (tw >= g.kW) break;
          const int ww = ow - g.pW + tw;
          const bool ok_sw = s >= 0 && s < g.S && ww >= 0 && ww < g.W;
          const T* xl = x + (((int64_t)b * g.S + s) * g.W + ww) * g.H;
#pragma unroll
          for (int th = 0; th < 3; th++) {
            if (th >= g.kH) break;
            const int h = oh - g.pH + th;
            float xv = 0.f;
            if (ok_sw && h >= 0 && h < g.H) xv = Elem<T>::ld1(xl + h);
            // tap index is compile-time only when the loops are fully unrolled: select by comparison
#pragma unroll
            for (int t = 0; t < NT; t++)
              if (t == tap) {
#pragma unroll
                for (int i = 0; i < ST_C; i++) acc[t][i] = fmaf(xv, d[i], acc[t][i]);
              }
            tap++;
          }
        }
      }
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int t = 0; t < NT; t++)
#pragma unroll
    for (int i = 0; i < ST_C; i++) {
      const float a = warp_sum(acc[t][i]);
      if (lane == 0) red[(wid * NT + t) * ST_C + i] = a;
    }
  __syncthreads();
  for (int i = threadIdx.x; i < NT * ST_C; i += ST_THREADS) {
    float v = 0.f;
    for (int k = 0; k < ST_THREADS / 32; k++) v += red[k * NT * ST_C + i];
    const int t = i / ST_C, co = i % ST_C;
    if (part) part[(size_t)blockIdx.x * (NT * ST_C) + co * NT + t] = v;   // per-block slice, summed in block order afterwards
    else atomicAdd(dw + co * NT + t, v);     // master layout [Cout][Cin=1][taps]
  }
}

// Same kernel with compile-time kernel extents: the tap loops unroll completely.
template <typename T, int KS, int KW, int KH>
__global__ void __launch_bounds__(ST_THREADS)
stem_fwd_kernel_t(StemGeom g, const T* __restrict__ x, const float* __restrict__ w, T* __restrict__ y,
                float* __restrict__ stat_partial) {
  pdl_prologue();
  __shared__ float w_s[ST_MAXTAPS * ST_C];
  __shared__ float red[(ST_THREADS / 32) * 2 * ST_C];
  constexpr int ntaps = KS * KW * KH;
  for (int i = threadIdx.x; i < ntaps * ST_C; i += ST_THREADS) w_s[i] = w[(i % ST_C) * ntaps + i / ST_C];   // [tap][co]
  __syncthreads();
  float ssum[ST_C], ssq[ST_C];
#pragma unroll
  for (int i = 0; i < ST_C; i++) { ssum[i] = 0.f; ssq[i] = 0.f; }
  const int64_t nlines = (int64_t)g.B * g.oS * g.oW;
  for (int64_t line = blockIdx.x; line < nlines; line += gridDim.x) {
    const int ow = (int)(line % g.oW);
    const int os = (int)((line / g.oW) % g.oS);
    const int b = (int)(line / ((int64_t)g.oW * g.oS));
    for (int oh = threadIdx.x; oh < g.oH; oh += ST_THREADS) {
      float acc[ST_C];
#pragma unroll
      for (int i = 0; i < ST_C; i++) acc[i] = 0.f;
#pragma unroll
      for (int ts = 0; ts < KS; ts++) {
        const int s = os - g.pS + ts;
#pragma unroll
        for (int tw = 0; tw < KW; tw++) {
          const int ww = ow - g.pW + tw;
          const bool ok_sw = s >= 0 && s < g.S && ww >= 0 && ww < g.W;
          const T* xl = x + (((int64_t)b * g.S + s) * g.W + ww) * g.H;
#pragma unroll
          for (int th = 0; th < KH; th++) {
            constexpr int dummy = 0; (void)dummy;
            const int tap = (ts * KW + tw) * KH + th;
            const int h = oh - g.pH + th;
            if (ok_sw && h >= 0 && h < g.H) {
              const float xv = Elem<T>::ld1(xl + h);
              const float4* wr = reinterpret_cast<const float4*>(w_s + tap * ST_C);
#pragma unroll
              for (int q = 0; q < ST_C / 4; q++) {
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
      T* yp = y + (line * g.oH + oh) * ST_C;
#pragma unroll
      for (int i = 0; i < ST_C; i++) acc[i] = Elem<T>::rnd(acc[i]);
      constexpr int VEC = Elem<T>::VEC;
#pragma unroll
      for (int q = 0; q < ST_C / VEC; q++) {
        float v[VEC];
#pragma unroll
        for (int j = 0; j < VEC; j++) v[j] = acc[q * VEC + j];
        Elem<T>::store(yp + q * VEC, v);
      }
#pragma unroll
      for (int i = 0; i < ST_C; i++) { ssum[i] += acc[i]; ssq[i] = fmaf(acc[i], acc[i], ssq[i]); }
    }
  }
  if (stat_partial != nullptr) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < ST_C; i++) {
      const float a = warp_sum(ssum[i]), b2 = warp_sum(ssq[i]);
      if (lane == 0) { red[wid * 2 * ST_C + i] = a; red[wid * 2 * ST_C + ST_C + i] = b2; }
    }
    __syncthreads();
    if (threadIdx.x < 2 * ST_C) {
      float v = 0.f;
      for (int k = 0; k < ST_THREADS / 32; k++) v += red[k * 2 * ST_C + threadIdx.x];
      stat_partial[(int64_t)blockIdx.x * 2 * ST_C + threadIdx.x] = v;     // [row][2][16]
    }
  }
}


// Compile-time kernel extents: the tap index is a constant in every unrolled iteration (the generic kernel has to
// compare it against all NT accumulator rows).
template <typename T, int KS, int KW, int KH>
__global__ void __launch_bounds__(ST_THREADS)
stem_wgrad_kernel_t(StemGeom g, const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dw, float* __restrict__ part) {
  pdl_prologue();
  constexpr int NT = KS * KW * KH;
  __shared__ float red[(ST_THREADS / 32) * NT * ST_C];
  float acc[NT][ST_C];
#pragma unroll
  for (int t = 0; t < NT; t++)
#pragma unroll
    for (int i = 0; i < ST_C; i++) acc[t][i] = 0.f;
  const int64_t nlines = (int64_t)g.B * g.oS * g.oW;
  constexpr int VEC = Elem<T>::VEC;
  for (int64_t line = blockIdx.x; line < nlines; line += gridDim.x) {
    const int ow = (int)(line % g.oW);
    const int os = (int)((line / g.oW) % g.oS);
    const int b = (int)(line / ((int64_t)g.oW * g.oS));
    for (int oh = threadIdx.x; oh < g.oH; oh += ST_THREADS) {
      float d[ST_C];
      const T* dp = dy + (line * g.oH + oh) * ST_C;
#pragma unroll
      for (int q = 0; q < ST_C / VEC; q++) {
        float v[VEC];
        Elem<T>::load(dp + q * VEC, v);
#pragma unroll
        for (int j = 0; j < VEC; j++) d[q * VEC + j] = v[j];
      }
#pragma unroll
      for (int ts = 0; ts < KS; ts++) {
        const int s = os - g.pS + ts;
#pragma unroll
        for (int tw = 0; tw < KW; tw++) {
          const int ww = ow - g.pW + tw;
          const bool ok_sw = s >= 0 && s < g.S && ww >= 0 && ww < g.W;
          const T* xl = x + (((int64_t)b * g.S + s) * g.W + ww) * g.H;
#pragma unroll
          for (int th = 0; th < KH; th++) {
            constexpr int unused = 0; (void)unused;
            const int tap = (ts * KW + tw) * KH + th;
            const int h = oh - g.pH + th;
            float xv = 0.f;
            if (ok_sw && h >= 0 && h < g.H) xv = Elem<T>::ld1(xl + h);
#pragma unroll
            for (int i = 0; i < ST_C; i++) acc[tap][i] = fmaf(xv, d[i], acc[tap][i]);
          }
        }
      }
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int t = 0; t < NT; t++)
#pragma unroll
    for (int i = 0; i < ST_C; i++) {
      const float a = warp_sum(acc[t][i]);
      if (lane == 0) red[(wid * NT + t) * ST_C + i] = a;
    }
  __syncthreads();
  for (int i = threadIdx.x; i < NT * ST_C; i += ST_THREADS) {
    float v = 0.f;
    for (int k = 0; k < ST_THREADS / 32; k++) v += red[k * NT * ST_C + i];
    const int t = i / ST_C, co = i % ST_C;
    if (part) part[(size_t)blockIdx.x * (NT * ST_C) + co * NT + t] = v;   // per-block slice, summed in block order afterwards
    else atomicAdd(dw + co * NT + t, v);     // master layout [Cout][Cin=1][taps]
  }
}


// Forward, four consecutive depth positions per thread: every weight vector fetched from shared memory is used for four
// outputs (the one-position kernel is bound by its 36 LDS.128 per output), 128 contiguous bytes stored per thread.
template <typename T, int KS, int KW, int KH>
__global__ void __launch_bounds__(ST_THREADS)
stem_fwd_kernel_v4(StemGeom g, const T* __restrict__ x, const float* __restrict__ w, T* __restrict__ y,
                   float* __restrict__ stat_partial) {
  pdl_prologue();
  constexpr int NT = KS * KW * KH;
  __shared__ float w_s[NT * ST_C];
  __shared__ float red[(ST_THREADS / 32) * 2 * ST_C];
  for (int i = threadIdx.x; i < NT * ST_C; i += ST_THREADS) w_s[i] = w[(i % ST_C) * NT + i / ST_C];   // [tap][co]
  __syncthreads();
  float ssum[ST_C], ssq[ST_C];
#pragma unroll
  for (int i = 0; i < ST_C; i++) { ssum[i] = 0.f; ssq[i] = 0.f; }
  const uint32_t nq = (uint32_t)(g.oH + 3) >> 2;
  const uint32_t nlines = (uint32_t)g.B * g.oS * g.oW;
  const uint32_t total = nlines * nq;
  for (uint32_t it = blockIdx.x * ST_THREADS + threadIdx.x; it < total; it += gridDim.x * ST_THREADS) {
    const uint32_t line = it / nq;
    const int oh0 = (int)(it - line * nq) * 4;
    const int ow = (int)(line % (uint32_t)g.oW);
    const int os = (int)((line / (uint32_t)g.oW) % (uint32_t)g.oS);
    const int b = (int)(line / ((uint32_t)g.oW * g.oS));
    float acc[4][ST_C];
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
      for (int i = 0; i < ST_C; i++) acc[j][i] = 0.f;
#pragma unroll
    for (int ts = 0; ts < KS; ts++) {
      const int s = os - g.pS + ts;
#pragma unroll
      for (int tw = 0; tw < KW; tw++) {
        const int ww = ow - g.pW + tw;
        const bool ok_sw = s >= 0 && s < g.S && ww >= 0 && ww < g.W;
        const T* xl = x + (((int64_t)b * g.S + s) * g.W + ww) * g.H;
        float xv[KH + 3];
#pragma unroll
        for (int u = 0; u < KH + 3; u++) {
          const int h = oh0 - g.pH + u;
          xv[u] = (ok_sw && h >= 0 && h < g.H) ? Elem<T>::ld1(xl + h) : 0.f;
        }
#pragma unroll
        for (int th = 0; th < KH; th++) {
          const float4* wr = reinterpret_cast<const float4*>(w_s + ((ts * KW + tw) * KH + th) * ST_C);
#pragma unroll
          for (int q = 0; q < ST_C / 4; q++) {
            const float4 wv = wr[q];
#pragma unroll
            for (int j = 0; j < 4; j++) {
              acc[j][4 * q + 0] = fmaf(xv[th + j], wv.x, acc[j][4 * q + 0]);
              acc[j][4 * q + 1] = fmaf(xv[th + j], wv.y, acc[j][4 * q + 1]);
              acc[j][4 * q + 2] = fmaf(xv[th + j], wv.z, acc[j][4 * q + 2]);
              acc[j][4 * q + 3] = fmaf(xv[th + j], wv.w, acc[j][4 * q + 3]);
            }
          }
        }
      }
    }
    constexpr int VEC = Elem<T>::VEC;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if (oh0 + j < g.oH) {
        T* yp = y + ((int64_t)line * g.oH + oh0 + j) * ST_C;
#pragma unroll
        for (int i = 0; i < ST_C; i++) acc[j][i] = Elem<T>::rnd(acc[j][i]);
#pragma unroll
        for (int q = 0; q < ST_C / VEC; q++) {
          float v[VEC];
#pragma unroll
          for (int e = 0; e < VEC; e++) v[e] = acc[j][q * VEC + e];
          Elem<T>::store(yp + q * VEC, v);
        }
#pragma unroll
        for (int i = 0; i < ST_C; i++) { ssum[i] += acc[j][i]; ssq[i] = fmaf(acc[j][i], acc[j][i], ssq[i]); }
      }
    }
  }
  if (stat_partial != nullptr) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < ST_C; i++) {
      const float a = warp_sum(ssum[i]), b2 = warp_sum(ssq[i]);
      if (lane == 0) { red[wid * 2 * ST_C + i] = a; red[wid * 2 * ST_C + ST_C + i] = b2; }
    }
    __syncthreads();
    if (threadIdx.x < 2 * ST_C) {
      float v = 0.f;
      for (int k = 0; k < ST_THREADS / 32; k++) v += red[k * 2 * ST_C + threadIdx.x];
      stat_partial[(int64_t)blockIdx.x * 2 * ST_C + threadIdx.x] = v;     // [row][2][16]
    }
  }
}

// Weight gradient, channels split over two threads per position (72 accumulators each instead of 144): three times
// the occupancy of the one-thread kernel, whose 188 registers leave two blocks per SM.
template <typename T, int KS, int KW, int KH>
__global__ void __launch_bounds__(ST_THREADS)
stem_wgrad_kernel_v2(StemGeom g, const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dw, float* __restrict__ part) {
  pdl_prologue();
  constexpr int NT = KS * KW * KH;
  constexpr int HC = ST_C / 2;
  __shared__ float red[(ST_THREADS / 32) * NT * ST_C];
  float acc[NT][HC];
#pragma unroll
  for (int t = 0; t < NT; t++)
#pragma unroll
    for (int i = 0; i < HC; i++) acc[t][i] = 0.f;
  const int half = threadIdx.x & 1;
  const uint32_t nlines = (uint32_t)g.B * g.oS * g.oW;
  const uint32_t total = nlines * (uint32_t)g.oH;
  for (uint32_t it = (blockIdx.x * ST_THREADS + threadIdx.x) >> 1; it < total; it += (gridDim.x * ST_THREADS) >> 1) {
    const uint32_t line = it / (uint32_t)g.oH;
    const int oh = (int)(it - line * (uint32_t)g.oH);
    const int ow = (int)(line % (uint32_t)g.oW);
    const int os = (int)((line / (uint32_t)g.oW) % (uint32_t)g.oS);
    const int b = (int)(line / ((uint32_t)g.oW * g.oS));
    float d[HC];
    {
      const T* dp = dy + ((int64_t)line * g.oH + oh) * ST_C + half * HC;
      if (sizeof(T) == 2) {
        float v[Elem<T>::VEC];
        Elem<T>::load(dp, v);
#pragma unroll
        for (int i = 0; i < HC; i++) d[i] = v[i % Elem<T>::VEC];
      } else {
#pragma unroll
        for (int q = 0; q < HC / Elem<T>::VEC; q++) {
          float v[Elem<T>::VEC];
          Elem<T>::load(dp + q * Elem<T>::VEC, v);
#pragma unroll
          for (int e = 0; e < Elem<T>::VEC; e++) d[q * Elem<T>::VEC + e] = v[e];
        }
      }
    }
#pragma unroll
    for (int ts = 0; ts < KS; ts++) {
      const int s = os - g.pS + ts;
#pragma unroll
      for (int tw = 0; tw < KW; tw++) {
        const int ww = ow - g.pW + tw;
        const bool ok_sw = s >= 0 && s < g.S && ww >= 0 && ww < g.W;
        const T* xl = x + (((int64_t)b * g.S + s) * g.W + ww) * g.H;
#pragma unroll
        for (int th = 0; th < KH; th++) {
          const int h = oh - g.pH + th;
          float xv = 0.f;
          if (ok_sw && h >= 0 && h < g.H) xv = Elem<T>::ld1(xl + h);
#pragma unroll
          for (int i = 0; i < HC; i++) acc[(ts * KW + tw) * KH + th][i] = fmaf(xv, d[i], acc[(ts * KW + tw) * KH + th][i]);
        }
      }
    }
  }
  // lanes with equal parity hold the same channel half: sum over the 16 lanes of each parity, then over the warps
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int t = 0; t < NT; t++)
#pragma unroll
    for (int i = 0; i < HC; i++) {
      float a = acc[t][i];
#pragma unroll
      for (int o = 16; o >= 2; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane < 2) red[(wid * NT + t) * ST_C + lane * HC + i] = a;
    }
  __syncthreads();
  for (int i = threadIdx.x; i < NT * ST_C; i += ST_THREADS) {
    float v = 0.f;
    for (int k = 0; k < ST_THREADS / 32; k++) v += red[k * NT * ST_C + i];
    const int t = i / ST_C, co = i % ST_C;
    if (part) part[(size_t)blockIdx.x * (NT * ST_C) + co * NT + t] = v;   // per-block slice, summed in block order afterwards
    else atomicAdd(dw + co * NT + t, v);     // master layout [Cout][Cin=1][taps]
  }
}

bool stem_geom(const ffpn_conv_desc* d, StemGeom& g) {
  if (d->Cin != 1 || d->Cout != ST_C) return false;
  if (d->sS != 1 || d->sW != 1 || d->sH != 1) return false;
  if (d->kS > 3 || d->kW > 3 || d->kH > 3 || d->kS * d->kW * d->kH > ST_MAXTAPS) return false;
  if (d->H == 1 && d->oH == 1 && d->kH == 1) {
    // 2-D image: make W the contiguous streaming axis (batch stays separate so that taps along S stay correct)
    g.B = (int)d->B; g.S = 1; g.W = (int)d->S; g.H = (int)d->W; g.oS = 1; g.oW = (int)d->oS; g.oH = (int)d->oW;
    g.kS = 1; g.kW = d->kS; g.kH = d->kW; g.pS = 0; g.pW = d->pS; g.pH = d->pW;
  } else {
    g.B = (int)d->B; g.S = (int)d->S; g.W = (int)d->W; g.H = (int)d->H; g.oS = (int)d->oS; g.oW = (int)d->oW; g.oH = (int)d->oH;
    g.kS = d->kS; g.kW = d->kW; g.kH = d->kH; g.pS = d->pS; g.pW = d->pW; g.pH = d->pH;
  }
  return true;
}

}  // namespace

bool ffpn_stem_supported(const ffpn_conv_desc* d) {
  StemGeom g;
  const int nt = d->kS * d->kW * d->kH;
  return stem_geom(d, g) && (nt == 1 || nt == 3 || nt == 9);
}

int ffpn_stem_fwd(ffpn_ctx* ctx, const ffpn_conv_desc* d, const void* x, const float* w, void* y, float* stat_partial,
                  int* stat_rows, cudaStream_t st) {
  StemGeom g;
  if (!stem_geom(d, g)) FFPN_FAIL(ctx, "stem_fwd: unsupported geometry");
  const int64_t nlines = (int64_t)g.B * g.oS * g.oW;
  const int grid = (int)(nlines < FFPN_STAT_ROWS ? nlines : FFPN_STAT_ROWS);
#define STEM_FW(KS, KW, KH)                                                                                                  \
  if (g.kS == KS && g.kW == KW && g.kH == KH) {                                                                              \
    if (d->dtype == FFPN_F32) ffpn_launch(stem_fwd_kernel_v4<float, KS, KW, KH>, grid, ST_THREADS, 0, st, g, (const float*)x, w, (float*)y, stat_partial); \
    else ffpn_launch(stem_fwd_kernel_v4<bf16, KS, KW, KH>, grid, ST_THREADS, 0, st, g, (const bf16*)x, w, (bf16*)y, stat_partial);       \
  } else
  STEM_FW(1, 3, 3) STEM_FW(1, 1, 1) STEM_FW(1, 1, 3) STEM_FW(1, 3, 1)
#undef STEM_FW
  if (d->dtype == FFPN_F32) ffpn_launch(stem_fwd_kernel<float>, grid, ST_THREADS, 0, st, g, (const float*)x, w, (float*)y, stat_partial);
  else ffpn_launch(stem_fwd_kernel<bf16>, grid, ST_THREADS, 0, st, g, (const bf16*)x, w, (bf16*)y, stat_partial);
  FFPN_CHECK_LAUNCH(ctx, "stem_fwd");
  if (stat_rows) *stat_rows = grid;
  return 0;
}

namespace {
// dw[i] += sum over the per-block slices in block order (bitwise reproducible)
__global__ void __launch_bounds__(128) stem_slice_reduce_kernel(const float* __restrict__ part, int nslices, int n, float* __restrict__ dw) {
  pdl_prologue();
  // one block per output: thread t adds slices t, t + 128, ... then a fixed-shape tree -> same bits every run
  __shared__ float red[128];
  const int i = blockIdx.x;
  float a = 0.f;
  for (int k = threadIdx.x; k < nslices; k += 128) a += part[(size_t)k * n + i];
  red[threadIdx.x] = a;
  __syncthreads();
#pragma unroll
  for (int o = 64; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) dw[i] += red[0];
}
}  // namespace

size_t ffpn_stem_wgrad_workspace_bytes(const ffpn_conv_desc* d) {
  return (size_t)148 * 8 * d->kS * d->kW * d->kH * ST_C * 4;     // one dW slice per block (grid <= 8 per SM)
}

int ffpn_stem_wgrad(ffpn_ctx* ctx, const ffpn_conv_desc* d, const void* x, const void* dy, float* dw, void* ws, size_t ws_bytes,
                    cudaStream_t st) {
  StemGeom g;
  if (!stem_geom(d, g)) FFPN_FAIL(ctx, "stem_wgrad: unsupported geometry");
  const int nt = g.kS * g.kW * g.kH;
  const int64_t nlines = (int64_t)g.B * g.oS * g.oW;
  const int64_t cap = (int64_t)ctx->num_sms * 4;
  const int grid = (int)(nlines < cap ? nlines : cap);
  // per-block slices in the workspace + ordered reduce when it fits, atomics on dw otherwise
  float* part = (ws && (size_t)grid * 2 * nt * ST_C * 4 <= ws_bytes) ? (float*)ws : nullptr;
  int nblocks = grid;
#define STEM_FINISH()                                                                                                        \
  if (part) {                                                                                                                \
    ffpn_launch(stem_slice_reduce_kernel, nt * ST_C, 128, 0, st, (const float*)part, nblocks, nt * ST_C, dw);  \
    FFPN_CHECK_LAUNCH(ctx, "stem_slice_reduce");                                                                             \
  }
#define STEM_WT(KS, KW, KH)                                                                                                  \
  if (g.kS == KS && g.kW == KW && g.kH == KH) {                                                                              \
    if (d->dtype == FFPN_F32) ffpn_launch(stem_wgrad_kernel_v2<float, KS, KW, KH>, grid * 2, ST_THREADS, 0, st, g, (const float*)x, (const float*)dy, dw, part); \
    else ffpn_launch(stem_wgrad_kernel_v2<bf16, KS, KW, KH>, grid * 2, ST_THREADS, 0, st, g, (const bf16*)x, (const bf16*)dy, dw, part);      \
    FFPN_CHECK_LAUNCH(ctx, "stem_wgrad");                                                                                    \
    nblocks = grid * 2;                                                                                                      \
    STEM_FINISH()                                                                                                            \
    return 0;                                                                                                                \
  }
  STEM_WT(1, 3, 3) STEM_WT(1, 1, 1) STEM_WT(1, 1, 3) STEM_WT(1, 3, 1)
#undef STEM_WT
#define STEM_WG(T, NT) ffpn_launch(stem_wgrad_kernel<T, NT>, grid, ST_THREADS, 0, st, g, (const T*)x, (const T*)dy, dw, part)
  if (d->dtype == FFPN_F32) { if (nt == 1) STEM_WG(float, 1); else if (nt == 3) STEM_WG(float, 3); else if (nt == 9) STEM_WG(float, 9); else FFPN_FAIL(ctx, "stem_wgrad: taps"); }
  else { if (nt == 1) STEM_WG(bf16, 1); else if (nt == 3) STEM_WG(bf16, 3); else if (nt == 9) STEM_WG(bf16, 9); else FFPN_FAIL(ctx, "stem_wgrad: taps"); }
#undef STEM_WG
  FFPN_CHECK_LAUNCH(ctx, "stem_wgrad");
  STEM_FINISH()
#undef STEM_FINISH
  return 0;
}
