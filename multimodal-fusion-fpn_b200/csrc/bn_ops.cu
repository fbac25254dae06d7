// BatchNorm statistics / backward, residual block end (+ max-pool gradient routing) and max pooling.
// All of these are HBM-bound streaming kernels: 16-byte vector accesses, channels innermost so that a warp
// touches a contiguous span, grid-stride with a bounded grid so the per-block partial sums fit a fixed buffer.
#include "common.cuh"

namespace {

constexpr int EW_THREADS = 256;

// ---- finalize: partial sums -> mean / invstd / scale / shift, running stats -------------------------
// 32 channels per block, 8 row lanes; loads are coalesced across channels and the 8 lane partials are added
// in a fixed order, so the result is deterministic.
constexpr int FIN_LANES = 32;

__device__ __forceinline__ void column_sums(const float* __restrict__ partial, int rows, int ncols, int col_a, int col_b,
                                            int C, int c, int rl, double (*red)[32][2], double& sa, double& sb) {
  double a = 0.0, b = 0.0;
  if (c < C) {
    int r = rl;
    for (; r + 3 * FIN_LANES < rows; r += 4 * FIN_LANES) {        // 8 independent loads in flight
      float va[4], vb[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        va[u] = partial[((int64_t)(r + u * FIN_LANES) * ncols + col_a) * C + c];
        vb[u] = partial[((int64_t)(r + u * FIN_LANES) * ncols + col_b) * C + c];
      }
#pragma unroll
      for (int u = 0; u < 4; u++) { a += (double)va[u]; b += (double)vb[u]; }
    }
    for (; r < rows; r += FIN_LANES) {
      a += (double)partial[((int64_t)r * ncols + col_a) * C + c];
      b += (double)partial[((int64_t)r * ncols + col_b) * C + c];
    }
  }
  red[rl][threadIdx.x & 31][0] = a;
  red[rl][threadIdx.x & 31][1] = b;
  __syncthreads();
  sa = sb = 0.0;
  for (int l = 0; l < FIN_LANES; l++) { sa += red[l][threadIdx.x & 31][0]; sb += red[l][threadIdx.x & 31][1]; }
}

__global__ void __launch_bounds__(32 * FIN_LANES)
bn_finalize_kernel(const float* __restrict__ partial, int rows, int C, double count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, float momentum,
                                   float eps, int training, float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ save_mean, float* __restrict__ save_invstd) {
  pdl_prologue();
  __shared__ double red[FIN_LANES][32][2];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rl = threadIdx.x >> 5;
  double mean, var;
  if (training) {
    double s, q;
    column_sums(partial, rows, 2, 0, 1, C, c, rl, red, s, q);
    mean = s / count;
    var = q / count - mean * mean;
    if (var < 0.0) var = 0.0;
  } else {
    mean = c < C ? (double)running_mean[c] : 0.0;
    var = c < C ? (double)running_var[c] : 1.0;
  }
  if (rl == 0 && c < C) {
    const double invstd = 1.0 / sqrt(var + (double)eps);
    const float a = (float)((double)gamma[c] * invstd);
    scale[c] = a;
    shift[c] = (float)((double)beta[c] - mean * (double)gamma[c] * invstd);
    if (save_mean) save_mean[c] = (float)mean;
    if (save_invstd) save_invstd[c] = (float)invstd;
    if (training && running_mean != nullptr) {
      const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
      running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * mean);
      running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * unb);
    }
  }
}

// Block-level reduction of NCOL per-thread channel-vector partials into partial[blockIdx.x][col][C].
// Thread layout: threadIdx.x % cvecs = channel vector, threadIdx.x / cvecs = position lane.
template <int VEC, int NCOL>
__device__ __forceinline__ void block_reduce_cols(float (&acc)[NCOL][VEC], int C, int cvecs, float* smem,
                                                  float* __restrict__ partial) {
  // smem: [NCOL][EW_THREADS][VEC]
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NCOL; k++)
#pragma unroll
    for (int j = 0; j < VEC; j++) smem[(k * EW_THREADS + threadIdx.x) * VEC + j] = acc[k][j];
  __syncthreads();
  const int lanes = EW_THREADS / cvecs;   // position lanes per channel vector
  for (int i = threadIdx.x; i < NCOL * C; i += EW_THREADS) {
    const int k = i / C, c = i % C;
    const int cv = c / VEC, j = c % VEC;
    float s = 0.f;
    for (int l = 0; l < lanes; l++) s += smem[(k * EW_THREADS + l * cvecs + cv) * VEC + j];
    partial[((int64_t)blockIdx.x * NCOL + k) * C + c] = s;
  }
}

// ---- backward pass 1: sums of G and G*y ---------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(EW_THREADS)
bn_bwd_reduce_kernel(int64_t P, int C, const T* __restrict__ dA, const T* __restrict__ y,
                     const float* __restrict__ scale, const float* __restrict__ shift, int relu,
                     float* __restrict__ partial) {
  pdl_prologue();
  constexpr int VEC = Elem<T>::VEC;
  __shared__ float smem[2 * EW_THREADS * VEC];
  const int cvecs = C / VEC;
  const int lanes = EW_THREADS / cvecs;
  const int cv = threadIdx.x % cvecs, pl = threadIdx.x / cvecs;
  float acc[2][VEC];
#pragma unroll
  for (int j = 0; j < VEC; j++) acc[0][j] = acc[1][j] = 0.f;
  float sc[VEC], sh[VEC];
#pragma unroll
  for (int j = 0; j < VEC; j++) { sc[j] = scale[cv * VEC + j]; sh[j] = shift[cv * VEC + j]; }
  if (pl < lanes) {
    for (int64_t p = (int64_t)blockIdx.x * lanes + pl; p < P; p += (int64_t)gridDim.x * lanes) {
      float g[VEC], yv[VEC];
      Elem<T>::load(dA + p * C + cv * VEC, g);
      Elem<T>::load(y + p * C + cv * VEC, yv);
#pragma unroll
      for (int j = 0; j < VEC; j++) {
        float gj = g[j];
        if (relu && !(fmaf(yv[j], sc[j], sh[j]) > 0.f)) gj = 0.f;
        acc[0][j] += gj;
        acc[1][j] = fmaf(gj, yv[j], acc[1][j]);
      }
    }
  }
  block_reduce_cols<VEC, 2>(acc, C, cvecs, smem, partial);
}

// ---- backward pass 2: coefficients ---------------------------------------------------------------------
__global__ void __launch_bounds__(32 * FIN_LANES)
bn_bwd_finalize_kernel(const float* __restrict__ partial, int rows, int ncols, int ycol, int C,
                                       double count, const float* __restrict__ gamma,
                                       const float* __restrict__ save_mean, const float* __restrict__ save_invstd,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ cA,
                                       float* __restrict__ cP, float* __restrict__ cQ) {
  pdl_prologue();
  __shared__ double red[FIN_LANES][32][2];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rl = threadIdx.x >> 5;
  double s1, s2;
  column_sums(partial, rows, ncols, 0, ycol, C, c, rl, red, s1, s2);
  if (rl == 0 && c < C) {
    const double mu = (double)save_mean[c], is = (double)save_invstd[c], gm = (double)gamma[c];
    const double dg = is * (s2 - mu * s1);   // sum G * xhat
    const double a = gm * is;
    const double k = is * is * (s2 - mu * s1) / count;
    dgamma[c] = (float)dg;
    dbeta[c] = (float)s1;
    cA[c] = (float)a;
    cP[c] = (float)(-a * k);
    cQ[c] = (float)(a * (mu * k - s1 / count));
  }
}

// ---- backward pass 3: dy = cA*G + cP*y + cQ ---------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(EW_THREADS)
bn_bwd_apply_kernel(int64_t nvec, int C, const T* __restrict__ dA, const T* __restrict__ y,
                    const float* __restrict__ scale, const float* __restrict__ shift, int relu,
                    const float* __restrict__ cA, const float* __restrict__ cP, const float* __restrict__ cQ,
                    T* __restrict__ dy) {
  pdl_prologue();
  constexpr int VEC = Elem<T>::VEC;
  const int cvecs = C / VEC;
  for (int64_t i = (int64_t)blockIdx.x * EW_THREADS + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * EW_THREADS) {
    const int c0 = (int)(i % cvecs) * VEC;
    float g[VEC], yv[VEC], o[VEC];
    Elem<T>::load(dA + i * VEC, g);
    Elem<T>::load(y + i * VEC, yv);
#pragma unroll
    for (int j = 0; j < VEC; j++) {
      float gj = g[j];
      if (relu && !(fmaf(yv[j], scale[c0 + j], shift[c0 + j]) > 0.f)) gj = 0.f;
      o[j] = fmaf(cA[c0 + j], gj, fmaf(cP[c0 + j], yv[j], cQ[c0 + j]));
    }
    Elem<T>::store(dy + i * VEC, o);
  }
}

// ---- backward pass 3 for a block with a conv + BN shortcut: the block's last BN and the shortcut's BN both receive the
// same masked gradient G (block_end_bwd).  One pass reads G once and writes both conv gradients.
template <typename T>
__global__ void __launch_bounds__(EW_THREADS)
bn_bwd_apply2_kernel(int64_t nvec, int C, const T* __restrict__ G, const T* __restrict__ y1, const T* __restrict__ y2,
                     const float* __restrict__ cA1, const float* __restrict__ cP1, const float* __restrict__ cQ1,
                     const float* __restrict__ cA2, const float* __restrict__ cP2, const float* __restrict__ cQ2,
                     T* __restrict__ dy1, T* __restrict__ dy2) {
  pdl_prologue();
  constexpr int VEC = Elem<T>::VEC;
  const int cvecs = C / VEC;
  for (int64_t i = (int64_t)blockIdx.x * EW_THREADS + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * EW_THREADS) {
    const int c0 = (int)(i % cvecs) * VEC;
    float g[VEC], a[VEC], b[VEC], o1[VEC], o2[VEC];
    Elem<T>::load(G + i * VEC, g);
    Elem<T>::load(y1 + i * VEC, a);
    Elem<T>::load(y2 + i * VEC, b);
#pragma unroll
    for (int j = 0; j < VEC; j++) {
      o1[j] = fmaf(cA1[c0 + j], g[j], fmaf(cP1[c0 + j], a[j], cQ1[c0 + j]));
      o2[j] = fmaf(cA2[c0 + j], g[j], fmaf(cP2[c0 + j], b[j], cQ2[c0 + j]));
    }
    Elem<T>::store(dy1 + i * VEC, o1);
    Elem<T>::store(dy2 + i * VEC, o2);
  }
}

// ---- block end forward ----------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(EW_THREADS)
block_end_fwd_kernel(int64_t nvec, int C, const T* __restrict__ y, const float* __restrict__ a,
                     const float* __restrict__ b, const T* __restrict__ res, const float* __restrict__ ra,
                     const float* __restrict__ rb, T* __restrict__ z) {
  pdl_prologue();
  constexpr int VEC = Elem<T>::VEC;
  const int cvecs = C / VEC;
  for (int64_t i = (int64_t)blockIdx.x * EW_THREADS + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * EW_THREADS) {
    const int c0 = (int)(i % cvecs) * VEC;
    float yv[VEC], o[VEC];
    Elem<T>::load(y + i * VEC, yv);
#pragma unroll
    for (int j = 0; j < VEC; j++) o[j] = fmaf(yv[j], a[c0 + j], b[c0 + j]);
    if (res != nullptr) {
      float rv[VEC];
      Elem<T>::load(res + i * VEC, rv);
#pragma unroll
      for (int j = 0; j < VEC; j++) o[j] += (ra != nullptr) ? fmaf(rv[j], ra[c0 + j], rb[c0 + j]) : rv[j];
    }
#pragma unroll
    for (int j = 0; j < VEC; j++) o[j] = (o[j] < 0.f) ? 0.f : o[j];   // NaN propagates, like torch.relu
    Elem<T>::store(z + i * VEC, o);
  }
}

// ---- block end backward (+ pool routing) ---------------------------------------------------------------
// POOL = false (no pooled-branch gradient to route: most launches of a step) compiles the window scan out: 121 -> ~60 registers,
// i.e. four instead of two resident CTAs per SM and twice the loads in flight.
template <typename T, int NCOL, bool POOL>
__global__ void __launch_bounds__(EW_THREADS, POOL ? 1 : 4)
block_end_bwd_kernel(int B, int S, int W, int H, int C, int kS, int kW, int kH, const T* __restrict__ dz,
                     const T* __restrict__ dzp, const T* __restrict__ z, const T* __restrict__ y,
                     const T* __restrict__ yres, T* __restrict__ G, float* __restrict__ partial) {
  pdl_prologue();
  constexpr int VEC = Elem<T>::VEC;
  __shared__ float smem[NCOL * EW_THREADS * VEC];
  const int cvecs = C / VEC;
  const int lanes = EW_THREADS / cvecs;
  const int cv = threadIdx.x % cvecs, pl = threadIdx.x / cvecs;
  const int64_t P = (int64_t)B * S * W * H;
  const int oS = kS ? S / kS : 0, oW = kW ? W / kW : 0, oH = kH ? H / kH : 0;
  float acc[NCOL][VEC];
#pragma unroll
  for (int k = 0; k < NCOL; k++)
#pragma unroll
    for (int j = 0; j < VEC; j++) acc[k][j] = 0.f;
  if (pl < lanes) {
    for (int64_t p = (int64_t)blockIdx.x * lanes + pl; p < P; p += (int64_t)gridDim.x * lanes) {
      const int64_t e = p * C + cv * VEC;
      float zv[VEC], g[VEC];
      Elem<T>::load(z + e, zv);
      if (dz != nullptr) Elem<T>::load(dz + e, g);
      else {
#pragma unroll
        for (int j = 0; j < VEC; j++) g[j] = 0.f;
      }
      if (POOL && dzp != nullptr) {
        int64_t r = p;
        const int h = (int)(r % H); r /= H;
        const int w = (int)(r % W); r /= W;
        const int s = (int)(r % S);
        const int b = (int)(r / S);
        const int os = s / kS, ow = w / kW, oh = h / kH;
        if (os < oS && ow < oW && oh < oH) {
          // first-max scan of the window in (s, w, h) order; this position wins channel j iff the scan's
          // winner is this position.  Rule: cand replaces best when cand > best or cand is NaN.
          float best[VEC];
          bool mine[VEC];
#pragma unroll
          for (int j = 0; j < VEC; j++) { best[j] = -INFINITY; mine[j] = false; }
          bool first = true;
          for (int ds = 0; ds < kS; ds++)
            for (int dw = 0; dw < kW; dw++)
              for (int dh = 0; dh < kH; dh++) {
                const int ss = os * kS + ds, ww = ow * kW + dw, hh = oh * kH + dh;
                const bool self = (ss == s) && (ww == w) && (hh == h);
                float cv_[VEC];
                if (self) {
#pragma unroll
                  for (int j = 0; j < VEC; j++) cv_[j] = zv[j];
                } else {
                  Elem<T>::load(z + ((((int64_t)b * S + ss) * W + ww) * H + hh) * C + cv * VEC, cv_);
                }
#pragma unroll
                for (int j = 0; j < VEC; j++) {
                  if (first || pool_better(cv_[j], best[j])) { best[j] = cv_[j]; mine[j] = self; }
                }
                first = false;
              }
          float gp[VEC];
          Elem<T>::load(dzp + ((((int64_t)b * oS + os) * oW + ow) * oH + oh) * C + cv * VEC, gp);
#pragma unroll
          for (int j = 0; j < VEC; j++)
            if (mine[j]) g[j] += gp[j];
        }
      }
      float yv[VEC], yr[VEC];
      Elem<T>::load(y + e, yv);
      if (NCOL == 3) Elem<T>::load(yres + e, yr);
#pragma unroll
      for (int j = 0; j < VEC; j++) {
        if (!(zv[j] > 0.f)) g[j] = 0.f;
        g[j] = Elem<T>::rnd(g[j]);
        acc[0][j] += g[j];
        acc[1][j] = fmaf(g[j], yv[j], acc[1][j]);
        if (NCOL == 3) acc[2][j] = fmaf(g[j], yr[j], acc[2][j]);
      }
      Elem<T>::store(G + e, g);
    }
  }
  block_reduce_cols<VEC, NCOL>(acc, C, cvecs, smem, partial);
}

// Window-major variant for extents that the pool kernel divides: one thread owns one pool window (per channel vector),
// reads its KS*KW*KH activations once, finds the first-max winner (same scan order / NaN rule as the forward pool) and
// then emits G for every position of the window.  The generic kernel above re-reads the whole window per position and
// pays 64-bit divisions per element (396 us at level 1 vs the 87 us its bytes need).
template <typename T, int NCOL, int KS, int KW, int KH>
__global__ void __launch_bounds__(EW_THREADS)
block_end_bwd_pool_kernel(int B, int S, int W, int H, int C, const T* __restrict__ dz, const T* __restrict__ dzp,
                          const T* __restrict__ z, const T* __restrict__ y, const T* __restrict__ yres, T* __restrict__ G,
                          float* __restrict__ partial) {
  pdl_prologue();
  constexpr int VEC = Elem<T>::VEC;
  constexpr int NW = KS * KW * KH;
  __shared__ float smem[NCOL * EW_THREADS * VEC];
  const int cvecs = C / VEC;
  const int lanes = EW_THREADS / cvecs;
  const int cv = threadIdx.x % cvecs, pl = threadIdx.x / cvecs;
  const int oS = S / KS, oW = W / KW, oH = H / KH;
  const uint32_t nwin = (uint32_t)B * oS * oW * oH;
  float acc[NCOL][VEC];
#pragma unroll
  for (int k = 0; k < NCOL; k++)
#pragma unroll
    for (int j = 0; j < VEC; j++) acc[k][j] = 0.f;
  if (pl < lanes) {
    for (uint32_t win = blockIdx.x * lanes + pl; win < nwin; win += gridDim.x * lanes) {
      uint32_t r = win;
      const int oh = (int)(r % (uint32_t)oH); r /= (uint32_t)oH;
      const int ow = (int)(r % (uint32_t)oW); r /= (uint32_t)oW;
      const int os = (int)(r % (uint32_t)oS);
      const int b = (int)(r / (uint32_t)oS);
      float gp[VEC];
      Elem<T>::load(dzp + (int64_t)win * C + cv * VEC, gp);
      float zw[NW][VEC];
      int winner[VEC];
      float best[VEC];
#pragma unroll
      for (int k = 0; k < NW; k++) {
        const int ds = k / (KW * KH), dw = (k / KH) % KW, dh = k % KH;
        const int64_t pos = (((int64_t)b * S + os * KS + ds) * W + ow * KW + dw) * H + oh * KH + dh;
        Elem<T>::load(z + pos * C + cv * VEC, zw[k]);
#pragma unroll
        for (int j = 0; j < VEC; j++)
          if (k == 0 || pool_better(zw[k][j], best[j])) { best[j] = zw[k][j]; winner[j] = k; }
      }
#pragma unroll
      for (int k = 0; k < NW; k++) {
        const int ds = k / (KW * KH), dw = (k / KH) % KW, dh = k % KH;
        const int64_t e = ((((int64_t)b * S + os * KS + ds) * W + ow * KW + dw) * H + oh * KH + dh) * C + cv * VEC;
        float g[VEC], yv[VEC], yr[VEC];
        if (dz != nullptr) Elem<T>::load(dz + e, g);
        else {
#pragma unroll
          for (int j = 0; j < VEC; j++) g[j] = 0.f;
        }
        Elem<T>::load(y + e, yv);
        if (NCOL == 3) Elem<T>::load(yres + e, yr);
#pragma unroll
        for (int j = 0; j < VEC; j++) {
          if (winner[j] == k) g[j] += gp[j];
          if (!(zw[k][j] > 0.f)) g[j] = 0.f;
          g[j] = Elem<T>::rnd(g[j]);
          acc[0][j] += g[j];
          acc[1][j] = fmaf(g[j], yv[j], acc[1][j]);
          if (NCOL == 3) acc[2][j] = fmaf(g[j], yr[j], acc[2][j]);
        }
        Elem<T>::store(G + e, g);
      }
    }
  }
  block_reduce_cols<VEC, NCOL>(acc, C, cvecs, smem, partial);
}

// ---- max pool forward -----------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(EW_THREADS)
maxpool_fwd_kernel(int B, int S, int W, int H, int C, int kS, int kW, int kH, const T* __restrict__ z,
                   T* __restrict__ zp, int64_t* __restrict__ idx) {
  pdl_prologue();
  constexpr int VEC = Elem<T>::VEC;
  const int cvecs = C / VEC;
  const int oS = S / kS, oW = W / kW, oH = H / kH;
  const int64_t nvec = (int64_t)B * oS * oW * oH * cvecs;
  for (int64_t i = (int64_t)blockIdx.x * EW_THREADS + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * EW_THREADS) {
    const int cv = (int)(i % cvecs);
    int64_t r = i / cvecs;
    const int oh = (int)(r % oH); r /= oH;
    const int ow = (int)(r % oW); r /= oW;
    const int os = (int)(r % oS);
    const int b = (int)(r / oS);
    float best[VEC];
    int bi[VEC];
    bool first = true;
    for (int ds = 0; ds < kS; ds++)
      for (int dw = 0; dw < kW; dw++)
        for (int dh = 0; dh < kH; dh++) {
          const int ss = os * kS + ds, ww = ow * kW + dw, hh = oh * kH + dh;
          float v[VEC];
          Elem<T>::load(z + ((((int64_t)b * S + ss) * W + ww) * H + hh) * C + cv * VEC, v);
          const int flat = (ss * W + ww) * H + hh;
#pragma unroll
          for (int j = 0; j < VEC; j++)
            if (first || pool_better(v[j], best[j])) { best[j] = v[j]; bi[j] = flat; }
          first = false;
        }
    Elem<T>::store(zp + i * VEC, best);
    if (idx != nullptr) {
#pragma unroll
      for (int j = 0; j < VEC; j++) idx[i * VEC + j] = bi[j];
    }
  }
}

// ---- stand-alone max pool backward (gather form: one thread per input position) ---------------------------
template <typename T>
__global__ void __launch_bounds__(EW_THREADS)
maxpool_bwd_kernel(int B, int S, int W, int H, int C, int kS, int kW, int kH, const T* __restrict__ z,
                   const T* __restrict__ dzp, T* __restrict__ dz) {
  pdl_prologue();
  constexpr int VEC = Elem<T>::VEC;
  const int cvecs = C / VEC;
  const int oS = S / kS, oW = W / kW, oH = H / kH;
  const int64_t nvec = (int64_t)B * S * W * H * cvecs;
  for (int64_t i = (int64_t)blockIdx.x * EW_THREADS + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * EW_THREADS) {
    const int cv = (int)(i % cvecs);
    int64_t r = i / cvecs;
    const int h = (int)(r % H); r /= H;
    const int w = (int)(r % W); r /= W;
    const int s = (int)(r % S);
    const int b = (int)(r / S);
    float g[VEC];
#pragma unroll
    for (int j = 0; j < VEC; j++) g[j] = 0.f;
    const int os = s / kS, ow = w / kW, oh = h / kH;
    if (os < oS && ow < oW && oh < oH) {
      float best[VEC];
      bool mine[VEC];
      bool first = true;
      for (int ds = 0; ds < kS; ds++)
        for (int dw = 0; dw < kW; dw++)
          for (int dh = 0; dh < kH; dh++) {
            const int ss = os * kS + ds, ww = ow * kW + dw, hh = oh * kH + dh;
            const bool self = (ss == s) && (ww == w) && (hh == h);
            float v[VEC];
            Elem<T>::load(z + ((((int64_t)b * S + ss) * W + ww) * H + hh) * C + cv * VEC, v);
#pragma unroll
            for (int j = 0; j < VEC; j++)
              if (first || pool_better(v[j], best[j])) { best[j] = v[j]; mine[j] = self; }
            first = false;
          }
      float gp[VEC];
      Elem<T>::load(dzp + ((((int64_t)b * oS + os) * oW + ow) * oH + oh) * C + cv * VEC, gp);
#pragma unroll
      for (int j = 0; j < VEC; j++)
        if (mine[j]) g[j] = gp[j];
    }
    Elem<T>::store(dz + i * VEC, g);
  }
}

inline int ew_grid(ffpn_ctx* ctx, int64_t items) { return ffpn_grid_for(items, EW_THREADS, ctx->num_sms * 8); }

}  // namespace

#define CHECK_C(ctx, C, vec, name) \
  if ((C) % (vec) != 0 || (C) / (vec) > EW_THREADS || (C) <= 0) FFPN_FAIL(ctx, "%s: C=%d must be a positive multiple of %d (<= %d vectors)", name, C, vec, EW_THREADS)

extern "C" int ffpn_bn_finalize(ffpn_ctx* ctx, const float* stat_partial, int stat_rows, int C, double count,
                                const float* gamma, const float* beta, float* running_mean, float* running_var,
                                float momentum, float eps, int training, float* scale, float* shift, float* save_mean,
                                float* save_invstd, void* stream) {
  if (training && (stat_partial == nullptr || stat_rows <= 0)) FFPN_FAIL(ctx, "bn_finalize: training needs partial sums");
  ffpn_launch(bn_finalize_kernel, (C + 31) / 32, 32 * FIN_LANES, 0, (cudaStream_t)stream, 
      stat_partial, stat_rows, C, count, gamma, beta, running_mean, running_var, momentum, eps, training, scale, shift,
      save_mean, save_invstd);
  FFPN_CHECK_LAUNCH(ctx, "bn_finalize");
  return 0;
}

extern "C" int ffpn_bn_bwd_reduce(ffpn_ctx* ctx, int dtype, int64_t P, int C, const void* dA, const void* y,
                                  const float* scale, const float* shift, int relu, float* partial, int* rows,
                                  void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == FFPN_F32) {
    CHECK_C(ctx, C, 4, "bn_bwd_reduce");
    const int lanes = EW_THREADS / (C / 4);
    const int g = ffpn_grid_for(P, lanes * 8, min(ctx->num_sms * 4, FFPN_STAT_ROWS));
    ffpn_launch(bn_bwd_reduce_kernel<float>, g, EW_THREADS, 0, st, P, C, (const float*)dA, (const float*)y, scale, shift, relu, partial);
    *rows = g;
  } else {
    CHECK_C(ctx, C, 8, "bn_bwd_reduce");
    const int lanes = EW_THREADS / (C / 8);
    const int g = ffpn_grid_for(P, lanes * 8, min(ctx->num_sms * 4, FFPN_STAT_ROWS));
    ffpn_launch(bn_bwd_reduce_kernel<bf16>, g, EW_THREADS, 0, st, P, C, (const bf16*)dA, (const bf16*)y, scale, shift, relu, partial);
    *rows = g;
  }
  FFPN_CHECK_LAUNCH(ctx, "bn_bwd_reduce");
  return 0;
}

extern "C" int ffpn_bn_bwd_finalize(ffpn_ctx* ctx, const float* partial, int rows, int ncols, int ycol, int C,
                                    double count, const float* gamma, const float* save_mean, const float* save_invstd,
                                    float* dgamma, float* dbeta, float* cA, float* cP, float* cQ, void* stream) {
  ffpn_launch(bn_bwd_finalize_kernel, (C + 31) / 32, 32 * FIN_LANES, 0, (cudaStream_t)stream, 
      partial, rows, ncols, ycol, C, count, gamma, save_mean, save_invstd, dgamma, dbeta, cA, cP, cQ);
  FFPN_CHECK_LAUNCH(ctx, "bn_bwd_finalize");
  return 0;
}

extern "C" int ffpn_bn_bwd_apply(ffpn_ctx* ctx, int dtype, int64_t P, int C, const void* dA, const void* y,
                                 const float* scale, const float* shift, int relu, const float* cA, const float* cP,
                                 const float* cQ, void* dy, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == FFPN_F32) {
    CHECK_C(ctx, C, 4, "bn_bwd_apply");
    const int64_t nvec = P * C / 4;
    ffpn_launch(bn_bwd_apply_kernel<float>, ew_grid(ctx, nvec), EW_THREADS, 0, st, nvec, C, (const float*)dA, (const float*)y, scale, shift, relu, cA, cP, cQ, (float*)dy);
  } else {
    CHECK_C(ctx, C, 8, "bn_bwd_apply");
    const int64_t nvec = P * C / 8;
    ffpn_launch(bn_bwd_apply_kernel<bf16>, ew_grid(ctx, nvec), EW_THREADS, 0, st, nvec, C, (const bf16*)dA, (const bf16*)y, scale, shift, relu, cA, cP, cQ, (bf16*)dy);
  }
  FFPN_CHECK_LAUNCH(ctx, "bn_bwd_apply");
  return 0;
}

extern "C" int ffpn_bn_bwd_apply2(ffpn_ctx* ctx, int dtype, int64_t P, int C, const void* G, const void* y1, const void* y2,
                                  const float* cA1, const float* cP1, const float* cQ1, const float* cA2, const float* cP2,
                                  const float* cQ2, void* dy1, void* dy2, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!G || !y1 || !y2 || !dy1 || !dy2 || !cA1 || !cP1 || !cQ1 || !cA2 || !cP2 || !cQ2) FFPN_FAIL(ctx, "bn_bwd_apply2: null argument");
  if (dtype == FFPN_F32) {
    CHECK_C(ctx, C, 4, "bn_bwd_apply2");
    const int64_t nvec = P * C / 4;
    ffpn_launch(bn_bwd_apply2_kernel<float>, ew_grid(ctx, nvec), EW_THREADS, 0, st, nvec, C, (const float*)G, (const float*)y1, (const float*)y2, cA1, cP1, cQ1, cA2, cP2, cQ2, (float*)dy1, (float*)dy2);
  } else {
    CHECK_C(ctx, C, 8, "bn_bwd_apply2");
    const int64_t nvec = P * C / 8;
    ffpn_launch(bn_bwd_apply2_kernel<bf16>, ew_grid(ctx, nvec), EW_THREADS, 0, st, nvec, C, (const bf16*)G, (const bf16*)y1, (const bf16*)y2, cA1, cP1, cQ1, cA2, cP2, cQ2, (bf16*)dy1, (bf16*)dy2);
  }
  FFPN_CHECK_LAUNCH(ctx, "bn_bwd_apply2");
  return 0;
}

extern "C" int ffpn_block_end_fwd(ffpn_ctx* ctx, int dtype, int64_t P, int C, const void* y, const float* a,
                                  const float* b, const void* res, const float* ra, const float* rb, void* z,
                                  void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == FFPN_F32) {
    CHECK_C(ctx, C, 4, "block_end_fwd");
    const int64_t nvec = P * C / 4;
    ffpn_launch(block_end_fwd_kernel<float>, ew_grid(ctx, nvec), EW_THREADS, 0, st, nvec, C, (const float*)y, a, b, (const float*)res, ra, rb, (float*)z);
  } else {
    CHECK_C(ctx, C, 8, "block_end_fwd");
    const int64_t nvec = P * C / 8;
    ffpn_launch(block_end_fwd_kernel<bf16>, ew_grid(ctx, nvec), EW_THREADS, 0, st, nvec, C, (const bf16*)y, a, b, (const bf16*)res, ra, rb, (bf16*)z);
  }
  FFPN_CHECK_LAUNCH(ctx, "block_end_fwd");
  return 0;
}

extern "C" int ffpn_block_end_bwd(ffpn_ctx* ctx, int dtype, int64_t B, int64_t S, int64_t W, int64_t H, int C, int kS,
                                  int kW, int kH, const void* dz, const void* dzp, const void* z, const void* y,
                                  const void* yres, void* G, float* partial, int* rows, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t P = B * S * W * H;
  if (dzp != nullptr && (kS <= 0 || kW <= 0 || kH <= 0)) FFPN_FAIL(ctx, "block_end_bwd: pool kernel missing");
  if (dzp == nullptr) { kS = kW = kH = 0; }
  const int vec = dtype == FFPN_F32 ? 4 : 8;
  CHECK_C(ctx, C, vec, "block_end_bwd");
  const int lanes = EW_THREADS / (C / vec);
  const int g = ffpn_grid_for(P, lanes * 4, min(ctx->num_sms * 4, FFPN_STAT_ROWS));
  if (dzp != nullptr && S % kS == 0 && W % kW == 0 && H % kH == 0 && B * (S / kS) * (W / kW) * (H / kH) < (1ll << 31)) {
    // window-major fast path for the pool kernels of the model
    const int64_t nwin = B * (S / kS) * (W / kW) * (H / kH);
    const int gw = ffpn_grid_for(nwin, lanes * 2, min(ctx->num_sms * 4, FFPN_STAT_ROWS));
#define LAUNCH_BP(T, N, KS_, KW_, KH_)                                                                                              \
    ffpn_launch(block_end_bwd_pool_kernel<T, N, KS_, KW_, KH_>, gw, EW_THREADS, 0, st, (int)B, (int)S, (int)W, (int)H, C, (const T*)dz, (const T*)dzp, \
                                                                              (const T*)z, (const T*)y, (const T*)yres, (T*)G, partial)
#define DISPATCH_BP(KS_, KW_, KH_)                                                                  \
    if (kS == KS_ && kW == KW_ && kH == KH_) {                                                      \
      if (dtype == FFPN_F32) { if (yres) LAUNCH_BP(float, 3, KS_, KW_, KH_); else LAUNCH_BP(float, 2, KS_, KW_, KH_); } \
      else { if (yres) LAUNCH_BP(bf16, 3, KS_, KW_, KH_); else LAUNCH_BP(bf16, 2, KS_, KW_, KH_); }  \
      *rows = gw;                                                                                   \
      FFPN_CHECK_LAUNCH(ctx, "block_end_bwd");                                                      \
      return 0;                                                                                     \
    }
    DISPATCH_BP(1, 2, 2) DISPATCH_BP(2, 2, 2) DISPATCH_BP(1, 2, 1) DISPATCH_BP(2, 2, 1) DISPATCH_BP(1, 1, 2)
#undef DISPATCH_BP
#undef LAUNCH_BP
  }
#define LAUNCH_BE(T, N, PL) ffpn_launch(block_end_bwd_kernel<T, N, PL>, g, EW_THREADS, 0, st, (int)B, (int)S, (int)W, (int)H, C, kS, kW, kH, (const T*)dz, (const T*)dzp, (const T*)z, (const T*)y, (const T*)yres, (T*)G, partial)
  if (dzp != nullptr) {
    if (dtype == FFPN_F32) { if (yres) LAUNCH_BE(float, 3, true); else LAUNCH_BE(float, 2, true); }
    else { if (yres) LAUNCH_BE(bf16, 3, true); else LAUNCH_BE(bf16, 2, true); }
  } else {
    if (dtype == FFPN_F32) { if (yres) LAUNCH_BE(float, 3, false); else LAUNCH_BE(float, 2, false); }
    else { if (yres) LAUNCH_BE(bf16, 3, false); else LAUNCH_BE(bf16, 2, false); }
  }
#undef LAUNCH_BE
  *rows = g;
  FFPN_CHECK_LAUNCH(ctx, "block_end_bwd");
  return 0;
}

extern "C" int ffpn_maxpool_fwd(ffpn_ctx* ctx, int dtype, int64_t B, int64_t S, int64_t W, int64_t H, int C, int kS,
                                int kW, int kH, const void* z, void* zp, int64_t* idx, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (kS <= 0 || kW <= 0 || kH <= 0) FFPN_FAIL(ctx, "maxpool: bad kernel");
  const int vec = dtype == FFPN_F32 ? 4 : 8;
  CHECK_C(ctx, C, vec, "maxpool_fwd");
  const int64_t nvec = B * (S / kS) * (W / kW) * (H / kH) * (C / vec);
  if (nvec <= 0) FFPN_FAIL(ctx, "maxpool: empty output");
  if (dtype == FFPN_F32)
    ffpn_launch(maxpool_fwd_kernel<float>, ew_grid(ctx, nvec), EW_THREADS, 0, st, (int)B, (int)S, (int)W, (int)H, C, kS, kW, kH, (const float*)z, (float*)zp, idx);
  else
    ffpn_launch(maxpool_fwd_kernel<bf16>, ew_grid(ctx, nvec), EW_THREADS, 0, st, (int)B, (int)S, (int)W, (int)H, C, kS, kW, kH, (const bf16*)z, (bf16*)zp, idx);
  FFPN_CHECK_LAUNCH(ctx, "maxpool_fwd");
  return 0;
}

extern "C" int ffpn_maxpool_bwd(ffpn_ctx* ctx, int dtype, int64_t B, int64_t S, int64_t W, int64_t H, int C, int kS,
                                int kW, int kH, const void* z, const void* dzp, void* dz, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (kS <= 0 || kW <= 0 || kH <= 0) FFPN_FAIL(ctx, "maxpool_bwd: bad kernel");
  const int vec = dtype == FFPN_F32 ? 4 : 8;
  CHECK_C(ctx, C, vec, "maxpool_bwd");
  const int64_t nvec = B * S * W * H * (C / vec);
  if (dtype == FFPN_F32)
    ffpn_launch(maxpool_bwd_kernel<float>, ew_grid(ctx, nvec), EW_THREADS, 0, st, (int)B, (int)S, (int)W, (int)H, C, kS, kW, kH, (const float*)z, (const float*)dzp, (float*)dz);
  else
    ffpn_launch(maxpool_bwd_kernel<bf16>, ew_grid(ctx, nvec), EW_THREADS, 0, st, (int)B, (int)S, (int)W, (int)H, C, kS, kW, kH, (const bf16*)z, (const bf16*)dzp, (bf16*)dz);
  FFPN_CHECK_LAUNCH(ctx, "maxpool_bwd");
  return 0;
}
