// Projection tail (BN+ReLU+depth mean), 2-D feature resize (copy / adaptive max / bilinear), integer nearest
// upsample, concat-slice copies, the 1x1x1 head, input packing and the fused SGD step.
#include "common.cuh"

namespace {

constexpr int TH = 256;

// Projection tail: out[ew, coff + c] = mean_h relu(a[c] * y[ew, h, c] + b[c])  (BatchNorm + ReLU + torch.mean(dim=4),
// fusion3D2D.py:527-536).  One thread owns a 16-byte channel chunk (8 bf16 / 4 fp32) of one en-face position and walks
// the surviving depth taps with vector loads: consecutive threads read consecutive chunks, so a warp streams whole
// channels-last rows; coefficients live in registers.  C % VEC != 0 falls back to the scalar kernel below.
template <typename T>
__global__ void __launch_bounds__(TH) proj_tail_fwd_vec_kernel(int64_t EW, int H, int C, const T* __restrict__ y,
                                                              const float* __restrict__ a, const float* __restrict__ b,
                                                              T* __restrict__ out, int ostride, int coff) {
  pdl_prologue();
  constexpr int VEC = Elem<T>::VEC;
  const int cvecs = C / VEC;
  const int64_t n = EW * cvecs;
  const float inv = 1.f / (float)H;
  for (int64_t i = (int64_t)blockIdx.x * TH + threadIdx.x; i < n; i += (int64_t)gridDim.x * TH) {
    const int cv = (int)(i % cvecs);
    const int64_t ew = i / cvecs;
    float sa[VEC], sb[VEC], acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; j++) { sa[j] = a[cv * VEC + j]; sb[j] = b[cv * VEC + j]; acc[j] = 0.f; }
    const T* src = y + (ew * H) * C + cv * VEC;
    for (int h = 0; h < H; h++) {
      float v[VEC];
      Elem<T>::load(src + (int64_t)h * C, v);
#pragma unroll
      for (int j = 0; j < VEC; j++) {
        const float t = fmaf(v[j], sa[j], sb[j]);
        acc[j] += (t < 0.f) ? 0.f : t;
      }
    }
#pragma unroll
    for (int j = 0; j < VEC; j++) acc[j] *= inv;
    Elem<T>::store(out + ew * ostride + coff + cv * VEC, acc);
  }
}

template <typename T>
__global__ void __launch_bounds__(TH) proj_tail_bwd_vec_kernel(int64_t EW, int H, int C, const T* __restrict__ dout, int ostride,
                                                              int coff, T* __restrict__ dA) {
  pdl_prologue();
  constexpr int VEC = Elem<T>::VEC;
  const int cvecs = C / VEC;
  const int64_t n = EW * cvecs;
  const float inv = 1.f / (float)H;
  for (int64_t i = (int64_t)blockIdx.x * TH + threadIdx.x; i < n; i += (int64_t)gridDim.x * TH) {
    const int cv = (int)(i % cvecs);
    const int64_t ew = i / cvecs;
    float g[VEC];
    Elem<T>::load(dout + ew * ostride + coff + cv * VEC, g);
#pragma unroll
    for (int j = 0; j < VEC; j++) g[j] *= inv;
    T* dst = dA + (ew * H) * C + cv * VEC;
    for (int h = 0; h < H; h++) Elem<T>::store(dst + (int64_t)h * C, g);
  }
}

template <typename T>
__global__ void proj_tail_fwd_kernel(int64_t EW, int H, int C, const T* __restrict__ y, const float* __restrict__ a,
                                     const float* __restrict__ b, T* __restrict__ out, int ostride, int coff) {
  pdl_prologue();
  const int64_t n = EW * C;
  const float inv = 1.f / (float)H;
  for (int64_t i = (int64_t)blockIdx.x * TH + threadIdx.x; i < n; i += (int64_t)gridDim.x * TH) {
    const int c = (int)(i % C);
    const int64_t ew = i / C;
    const float sa = a[c], sb = b[c];
    float s = 0.f;
    for (int h = 0; h < H; h++) {
      float v = fmaf(Elem<T>::ld1(y + (ew * H + h) * C + c), sa, sb);
      s += (v < 0.f) ? 0.f : v;
    }
    Elem<T>::st1(out + ew * ostride + coff + c, s * inv);
  }
}

template <typename T>
__global__ void proj_tail_bwd_kernel(int64_t EW, int H, int C, const T* __restrict__ dout, int ostride, int coff,
                                     T* __restrict__ dA) {
  pdl_prologue();
  const int64_t n = EW * H * C;
  const float inv = 1.f / (float)H;
  for (int64_t i = (int64_t)blockIdx.x * TH + threadIdx.x; i < n; i += (int64_t)gridDim.x * TH) {
    const int c = (int)(i % C);
    const int64_t ew = i / ((int64_t)C * H);
    Elem<T>::st1(dA + i, Elem<T>::ld1(dout + ew * ostride + coff + c) * inv);
  }
}

__device__ __forceinline__ int win_start(int i, int n_in, int n_out) { return (int)(((int64_t)i * n_in) / n_out); }
__device__ __forceinline__ int win_end(int i, int n_in, int n_out) {
  return (int)((((int64_t)(i + 1)) * n_in + n_out - 1) / n_out);
}

// area_pixel_compute_source_index(scale, dst, align_corners=false, cubic=false): max(0, (dst+0.5)*scale-0.5)
__device__ __forceinline__ void lin_src(int o, int n_in, int n_out, int& i0, int& i1, float& l1) {
  const float scale = (float)n_in / (float)n_out;
  float r = ((float)o + 0.5f) * scale - 0.5f;
  if (r < 0.f) r = 0.f;
  i0 = (int)r;
  if (i0 > n_in - 1) i0 = n_in - 1;
  i1 = i0 + ((i0 < n_in - 1) ? 1 : 0);
  l1 = r - (float)i0;
}

template <typename T>
__global__ void resize2d_fwd_kernel(int mode, int B, int Si, int Wi, int So, int Wo, int C, const T* __restrict__ x,
                                    T* __restrict__ out, int ostride, int coff, int32_t* __restrict__ idx) {
  pdl_prologue();
  const int64_t n = (int64_t)B * So * Wo * C;
  for (int64_t i = (int64_t)blockIdx.x * TH + threadIdx.x; i < n; i += (int64_t)gridDim.x * TH) {
    const int c = (int)(i % C);
    int64_t r = i / C;
    const int w = (int)(r % Wo); r /= Wo;
    const int s = (int)(r % So);
    const int b = (int)(r / So);
    const T* xb = x + (int64_t)b * Si * Wi * C + c;
    float v;
    if (mode == 0) {
      v = Elem<T>::ld1(xb + ((int64_t)s * Wi + w) * C);
    } else if (mode == 1) {
      const int s0 = win_start(s, Si, So), s1 = win_end(s, Si, So);
      const int w0 = win_start(w, Wi, Wo), w1 = win_end(w, Wi, Wo);
      float best = 0.f;
      int bi = 0;
      bool first = true;
      for (int ss = s0; ss < s1; ss++)
        for (int ww = w0; ww < w1; ww++) {
          const float cand = Elem<T>::ld1(xb + ((int64_t)ss * Wi + ww) * C);
          if (first || pool_better(cand, best)) { best = cand; bi = ss * Wi + ww; }
          first = false;
        }
      v = best;
      if (idx) idx[i] = bi;
    } else {
      int sa, sb, wa, wb;
      float ls, lw;
      lin_src(s, Si, So, sa, sb, ls);
      lin_src(w, Wi, Wo, wa, wb, lw);
      const float v00 = Elem<T>::ld1(xb + ((int64_t)sa * Wi + wa) * C), v01 = Elem<T>::ld1(xb + ((int64_t)sa * Wi + wb) * C);
      const float v10 = Elem<T>::ld1(xb + ((int64_t)sb * Wi + wa) * C), v11 = Elem<T>::ld1(xb + ((int64_t)sb * Wi + wb) * C);
      v = (1.f - ls) * ((1.f - lw) * v00 + lw * v01) + ls * ((1.f - lw) * v10 + lw * v11);
    }
    Elem<T>::st1(out + (((int64_t)b * So + s) * Wo + w) * ostride + coff + c, v);
  }
}

// Gather-form backward: one thread per INPUT element, loops over the outputs that can reference it
// (deterministic, no atomics).
template <typename T>
__global__ void resize2d_bwd_kernel(int mode, int B, int Si, int Wi, int So, int Wo, int C, const T* __restrict__ dout,
                                    int ostride, int coff, const int32_t* __restrict__ idx, T* __restrict__ dx) {
  pdl_prologue();
  const int64_t n = (int64_t)B * Si * Wi * C;
  for (int64_t i = (int64_t)blockIdx.x * TH + threadIdx.x; i < n; i += (int64_t)gridDim.x * TH) {
    uint32_t r32 = (uint32_t)i;                       // n < 2^31 is checked by the caller
    const int c = (int)(r32 % (uint32_t)C); r32 /= (uint32_t)C;
    const int w = (int)(r32 % (uint32_t)Wi); r32 /= (uint32_t)Wi;
    const int s = (int)(r32 % (uint32_t)Si);
    const int b = (int)(r32 / (uint32_t)Si);
    const T* db = dout + (int64_t)b * So * Wo * ostride + coff + c;
    float g = 0.f;
    if (mode == 0) {
      g = Elem<T>::ld1(db + ((int64_t)s * Wo + w) * ostride);
    } else if (mode == 1 && Si % So == 0 && Wi % Wo == 0) {
      // adaptive windows tile the input exactly: one output window per input position (the shapes of the model)
      const int os = s / (Si / So), ow = w / (Wi / Wo);
      const int64_t o = (((int64_t)b * So + os) * Wo + ow);
      if (idx[o * C + c] == s * Wi + w) g = Elem<T>::ld1(dout + o * ostride + coff + c);
    } else if (mode == 1) {
      // outputs o whose window [floor(o*in/out), ceil((o+1)*in/out)) contains s:  o in [lo, hi]
      int slo = (int)(((int64_t)s * So) / Si); while (slo > 0 && win_end(slo - 1, Si, So) > s) slo--;
      int wlo = (int)(((int64_t)w * Wo) / Wi); while (wlo > 0 && win_end(wlo - 1, Wi, Wo) > w) wlo--;
      const int self = s * Wi + w;
      for (int os = slo; os < So && win_start(os, Si, So) <= s; os++)
        for (int ow = wlo; ow < Wo && win_start(ow, Wi, Wo) <= w; ow++) {
          const int64_t o = (((int64_t)b * So + os) * Wo + ow);
          if (idx[o * C + c] == self) g += Elem<T>::ld1(dout + o * ostride + coff + c);
        }
    } else {
      // bilinear: scan the (small) set of outputs whose two taps can include s / w
      const float scs = (float)Si / (float)So, scw = (float)Wi / (float)Wo;
      int os0 = (int)floorf(((float)s - 1.f + 0.5f) / scs - 0.5f) - 1; if (os0 < 0) os0 = 0;
      int os1 = (int)ceilf(((float)s + 1.f + 0.5f) / scs - 0.5f) + 1; if (os1 > So - 1) os1 = So - 1;
      int ow0 = (int)floorf(((float)w - 1.f + 0.5f) / scw - 0.5f) - 1; if (ow0 < 0) ow0 = 0;
      int ow1 = (int)ceilf(((float)w + 1.f + 0.5f) / scw - 0.5f) + 1; if (ow1 > Wo - 1) ow1 = Wo - 1;
      for (int os = os0; os <= os1; os++) {
        int sa, sb; float ls;
        lin_src(os, Si, So, sa, sb, ls);
        float ws = 0.f;
        if (sa == s) ws += 1.f - ls;
        if (sb == s) ws += ls;
        if (ws == 0.f) continue;
        for (int ow = ow0; ow <= ow1; ow++) {
          int wa, wb; float lw;
          lin_src(ow, Wi, Wo, wa, wb, lw);
          float wwt = 0.f;
          if (wa == w) wwt += 1.f - lw;
          if (wb == w) wwt += lw;
          if (wwt == 0.f) continue;
          g += ws * wwt * Elem<T>::ld1(db + ((int64_t)os * Wo + ow) * ostride);
        }
      }
    }
    Elem<T>::st1(dx + i, g);
  }
}

// Adaptive-max backward when the windows tile the input exactly (Si % So == 0, Wi % Wo == 0): one output window per input
// position, a vector of channels per thread (16-byte accesses instead of one bf16 at a time).
template <typename T>
__global__ void resize2d_bwd_max_tiled_kernel(int B, int Si, int Wi, int So, int Wo, int C, const T* __restrict__ dout, int ostride,
                                              int coff, const int32_t* __restrict__ idx, T* __restrict__ dx) {
  pdl_prologue();
  constexpr int VEC = Elem<T>::VEC;
  const int cvecs = C / VEC;
  const uint32_t n = (uint32_t)B * Si * Wi * cvecs;
  const int fs = Si / So, fw = Wi / Wo;
  for (uint32_t i = blockIdx.x * TH + threadIdx.x; i < n; i += gridDim.x * TH) {
    uint32_t r = i;
    const int cv = (int)(r % (uint32_t)cvecs); r /= (uint32_t)cvecs;
    const int w = (int)(r % (uint32_t)Wi); r /= (uint32_t)Wi;
    const int s = (int)(r % (uint32_t)Si);
    const int b = (int)(r / (uint32_t)Si);
    const int64_t o = ((int64_t)b * So + s / fs) * Wo + w / fw;
    const int self = s * Wi + w;
    float g[VEC], out[VEC];
    Elem<T>::load(dout + o * ostride + coff + cv * VEC, g);
    const int4* ip = reinterpret_cast<const int4*>(idx + o * C + cv * VEC);
#pragma unroll
    for (int q = 0; q < VEC / 4; q++) {
      const int4 iv = ip[q];
      out[4 * q + 0] = iv.x == self ? g[4 * q + 0] : 0.f;
      out[4 * q + 1] = iv.y == self ? g[4 * q + 1] : 0.f;
      out[4 * q + 2] = iv.z == self ? g[4 * q + 2] : 0.f;
      out[4 * q + 3] = iv.w == self ? g[4 * q + 3] : 0.f;
    }
    Elem<T>::store(dx + (int64_t)i * VEC, out);
  }
}

template <typename T>
__global__ void upsample_fwd_kernel(int B, int Si, int Wi, int fS, int fW, int C, const T* __restrict__ x,
                                    T* __restrict__ out, int ostride, int coff) {
  pdl_prologue();
  const int So = Si * fS, Wo = Wi * fW;
  const int64_t n = (int64_t)B * So * Wo * C;
  for (int64_t i = (int64_t)blockIdx.x * TH + threadIdx.x; i < n; i += (int64_t)gridDim.x * TH) {
    const int c = (int)(i % C);
    int64_t r = i / C;
    const int w = (int)(r % Wo); r /= Wo;
    const int s = (int)(r % So);
    const int b = (int)(r / So);
    const float v = Elem<T>::ld1(x + (((int64_t)b * Si + s / fS) * Wi + w / fW) * C + c);
    Elem<T>::st1(out + (((int64_t)b * So + s) * Wo + w) * ostride + coff + c, v);
  }
}

template <typename T>
__global__ void upsample_bwd_kernel(int B, int Si, int Wi, int fS, int fW, int C, const T* __restrict__ dout,
                                    int ostride, int coff, T* __restrict__ dx) {
  pdl_prologue();
  const int So = Si * fS, Wo = Wi * fW;
  const int64_t n = (int64_t)B * Si * Wi * C;
  for (int64_t i = (int64_t)blockIdx.x * TH + threadIdx.x; i < n; i += (int64_t)gridDim.x * TH) {
    const int c = (int)(i % C);
    int64_t r = i / C;
    const int w = (int)(r % Wi); r /= Wi;
    const int s = (int)(r % Si);
    const int b = (int)(r / Si);
    float g = 0.f;
    for (int ds = 0; ds < fS; ds++)
      for (int dw = 0; dw < fW; dw++)
        g += Elem<T>::ld1(dout + (((int64_t)b * So + s * fS + ds) * Wo + w * fW + dw) * ostride + coff + c);
    Elem<T>::st1(dx + i, g);
  }
}

template <typename T>
__global__ void slice_copy_kernel(int64_t P, int C, const T* __restrict__ src, int sstride, int soff, T* __restrict__ dst,
                                  int dstride, int doff) {
  pdl_prologue();
  const int64_t n = P * C;
  for (int64_t i = (int64_t)blockIdx.x * TH + threadIdx.x; i < n; i += (int64_t)gridDim.x * TH) {
    const int c = (int)(i % C);
    const int64_t p = i / C;
    dst[p * dstride + doff + c] = src[p * sstride + soff + c];
  }
}

// logits[b][k][ew] = bias[k] + sum_c x[b,ew,c] * w[k][c]
// final1 (+ the model's sigmoid, fusion_nets.py:110,118, when act == 1): out[b][k][ew] = act(bias[k] + sum_c w[k][c] x[b][ew][c])
template <typename T>
__global__ void head_fwd_kernel(int B, int64_t EW, int C, int n, int act, const T* __restrict__ x, const float* __restrict__ w,
                                const float* __restrict__ bias, float* __restrict__ logits) {
  pdl_prologue();
  const int64_t tot = (int64_t)B * EW;
  for (int64_t i = (int64_t)blockIdx.x * TH + threadIdx.x; i < tot; i += (int64_t)gridDim.x * TH) {
    const int64_t b = i / EW, ew = i % EW;
    for (int k = 0; k < n; k++) {
      float s = bias ? bias[k] : 0.f;
      for (int c = 0; c < C; c++) s = fmaf(Elem<T>::ld1(x + i * C + c), w[k * C + c], s);
      if (act == 1) s = 1.f / (1.f + expf(-s));
      logits[(b * n + k) * EW + ew] = s;
    }
  }
}

// gradient wrt the pre-activation: dl = dout (act == 0) or dout * p * (1 - p) with p = the forward's sigmoid output
__device__ __forceinline__ float head_dl(const float* __restrict__ dout, const float* __restrict__ pred, int64_t o) {
  const float g = dout[o];
  if (pred == nullptr) return g;
  const float p = pred[o];
  return g * p * (1.f - p);
}

// dx[pos][c] = sum_k dl[k] w[k][c];  dw[k][c] = sum_pos dl[k] x[pos][c];  db[k] = sum_pos dl[k]
template <typename T>
__global__ void head_bwd_dx_kernel(int B, int64_t EW, int C, int n, const float* __restrict__ w,
                                   const float* __restrict__ dl, const float* __restrict__ pred, T* __restrict__ dx) {
  pdl_prologue();
  const int64_t tot = (int64_t)B * EW * C;
  for (int64_t i = (int64_t)blockIdx.x * TH + threadIdx.x; i < tot; i += (int64_t)gridDim.x * TH) {
    const int c = (int)(i % C);
    const int64_t pos = i / C;
    const int64_t b = pos / EW, ew = pos % EW;
    float s = 0.f;
    for (int k = 0; k < n; k++) s = fmaf(head_dl(dl, pred, (b * n + k) * EW + ew), w[k * C + c], s);
    Elem<T>::st1(dx + i, s);
  }
}

// Two stages, both deterministic: HB_BLOCKS blocks each reduce a slice of the positions for every (k, c) (thread = one
// (position lane, c) pair, so a warp reads whole channel rows), then one block adds the partials in a fixed order.
constexpr int HB_BLOCKS = 64;
template <typename T>
__global__ void __launch_bounds__(TH) head_bwd_dw_partial_kernel(int B, int64_t EW, int C, int n, const T* __restrict__ x,
                                                                const float* __restrict__ dl, const float* __restrict__ pred,
                                                                float* __restrict__ partial) {
  pdl_prologue();
  __shared__ float red[TH];
  const int C1 = C + 1;                                   // column C is the bias
  const int lanes = TH / C1;                              // position lanes per block (C = 16: 15)
  const int c = threadIdx.x % C1, pl = threadIdx.x / C1;
  const int64_t tot = (int64_t)B * EW;
  for (int k = 0; k < n; k++) {
    float s = 0.f;
    if (pl < lanes) {
      for (int64_t i = (int64_t)blockIdx.x * lanes + pl; i < tot; i += (int64_t)gridDim.x * lanes) {
        const int64_t b = i / EW, ew = i - b * EW;
        const float g = head_dl(dl, pred, (b * n + k) * EW + ew);
        s = fmaf(g, c < C ? Elem<T>::ld1(x + i * C + c) : 1.f, s);
      }
    }
    __syncthreads();
    red[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x < C1) {
      float t = 0.f;
      for (int l = 0; l < lanes; l++) t += red[l * C1 + threadIdx.x];
      partial[((int64_t)blockIdx.x * n + k) * C1 + threadIdx.x] = t;
    }
  }
}
__global__ void __launch_bounds__(TH) head_bwd_dw_final_kernel(int nblocks, int C, int n, const float* __restrict__ partial,
                                                              float* __restrict__ dw, float* __restrict__ dbias) {
  pdl_prologue();
  const int C1 = C + 1;
  for (int i = threadIdx.x; i < n * C1; i += TH) {
    const int k = i / C1, c = i % C1;
    double t = 0.0;
    for (int r = 0; r < nblocks; r++) t += (double)partial[((int64_t)r * n + k) * C1 + c];
    if (c < C) dw[k * C + c] = (float)t;
    else dbias[k] = (float)t;
  }
}

// (R, H, W) fp32 -> (R, W, H) T, 32x32 smem tile transpose
template <typename T>
__global__ void pack_volume_kernel(int64_t R, int H, int W, const float* __restrict__ src, T* __restrict__ dst) {
  pdl_prologue();
  __shared__ float tile[32][33];
  const int tw = (W + 31) / 32, thh = (H + 31) / 32;
  const int64_t ntile = R * tw * thh;
  for (int64_t t = blockIdx.x; t < ntile; t += gridDim.x) {
    const int64_t r = t / (tw * thh);
    const int th_ = (int)((t / tw) % thh), tw_ = (int)(t % tw);
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;   // 256 threads: 8 rows per pass
    __syncthreads();
    for (int yy = ly; yy < 32; yy += 8) {
      const int h = th_ * 32 + yy, w = tw_ * 32 + lx;
      if (h < H && w < W) tile[yy][lx] = src[(r * H + h) * W + w];
    }
    __syncthreads();
    for (int yy = ly; yy < 32; yy += 8) {
      const int w = tw_ * 32 + yy, h = th_ * 32 + lx;
      if (h < H && w < W) Elem<T>::st1(dst + (r * W + w) * H + h, tile[lx][yy]);
    }
  }
}

template <typename T>
__global__ void cast_kernel(int64_t n, const float* __restrict__ src, T* __restrict__ dst) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * TH + threadIdx.x; i < n; i += (int64_t)gridDim.x * TH) Elem<T>::st1(dst + i, src[i]);
}

__global__ void sgd_kernel(int64_t n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ mom, float lr,
                           float momentum, float wd, float gscale, int first) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * TH + threadIdx.x; i < n; i += (int64_t)gridDim.x * TH) {
    const float pv = p[i];
    const float d = fmaf(wd, pv, g[i] * gscale);
    const float m = first ? d : fmaf(momentum, mom[i], d);
    mom[i] = m;
    p[i] = pv - lr * m;
  }
}

inline int grid_of(ffpn_ctx* ctx, int64_t n) { return ffpn_grid_for(n, TH, ctx->num_sms * 8); }

}  // namespace

#define DISPATCH(dtype, KERNEL, grid, ...)                                                     \
  do {                                                                                         \
    if ((dtype) == FFPN_F32) ffpn_launch(KERNEL<float>, grid, TH, 0, (cudaStream_t)stream, __VA_ARGS__); \
    else ffpn_launch(KERNEL<bf16>, grid, TH, 0, (cudaStream_t)stream, __VA_ARGS__);                    \
  } while (0)

extern "C" int ffpn_proj_tail_fwd(ffpn_ctx* ctx, int dtype, int64_t EW, int64_t H, int C, const void* y, const float* a,
                                  const float* b, void* out, int ostride, int coff, void* stream) {
  if (H <= 0) FFPN_FAIL(ctx, "proj_tail_fwd: empty depth");
  const int vec = dtype == FFPN_F32 ? 4 : 8;
  if (C % vec == 0 && ostride % vec == 0 && coff % vec == 0) {
    const int g = grid_of(ctx, EW * (C / vec));
    if (dtype == FFPN_F32) ffpn_launch(proj_tail_fwd_vec_kernel<float>, g, TH, 0, (cudaStream_t)stream, EW, (int)H, C, (const float*)y, a, b, (float*)out, ostride, coff);
    else ffpn_launch(proj_tail_fwd_vec_kernel<bf16>, g, TH, 0, (cudaStream_t)stream, EW, (int)H, C, (const bf16*)y, a, b, (bf16*)out, ostride, coff);
  } else if (dtype == FFPN_F32) ffpn_launch(proj_tail_fwd_kernel<float>, grid_of(ctx, EW * C), TH, 0, (cudaStream_t)stream, EW, (int)H, C, (const float*)y, a, b, (float*)out, ostride, coff);
  else ffpn_launch(proj_tail_fwd_kernel<bf16>, grid_of(ctx, EW * C), TH, 0, (cudaStream_t)stream, EW, (int)H, C, (const bf16*)y, a, b, (bf16*)out, ostride, coff);
  FFPN_CHECK_LAUNCH(ctx, "proj_tail_fwd");
  return 0;
}

extern "C" int ffpn_proj_tail_bwd(ffpn_ctx* ctx, int dtype, int64_t EW, int64_t H, int C, const void* dout, int ostride,
                                  int coff, void* dA, void* stream) {
  const int vec = dtype == FFPN_F32 ? 4 : 8;
  if (C % vec == 0 && ostride % vec == 0 && coff % vec == 0) {
    const int g = grid_of(ctx, EW * (C / vec));
    if (dtype == FFPN_F32) ffpn_launch(proj_tail_bwd_vec_kernel<float>, g, TH, 0, (cudaStream_t)stream, EW, (int)H, C, (const float*)dout, ostride, coff, (float*)dA);
    else ffpn_launch(proj_tail_bwd_vec_kernel<bf16>, g, TH, 0, (cudaStream_t)stream, EW, (int)H, C, (const bf16*)dout, ostride, coff, (bf16*)dA);
  } else if (dtype == FFPN_F32) ffpn_launch(proj_tail_bwd_kernel<float>, grid_of(ctx, EW * H * C), TH, 0, (cudaStream_t)stream, EW, (int)H, C, (const float*)dout, ostride, coff, (float*)dA);
  else ffpn_launch(proj_tail_bwd_kernel<bf16>, grid_of(ctx, EW * H * C), TH, 0, (cudaStream_t)stream, EW, (int)H, C, (const bf16*)dout, ostride, coff, (bf16*)dA);
  FFPN_CHECK_LAUNCH(ctx, "proj_tail_bwd");
  return 0;
}

extern "C" int ffpn_resize2d_fwd(ffpn_ctx* ctx, int dtype, int mode, int64_t B, int64_t Si, int64_t Wi, int64_t So,
                                 int64_t Wo, int C, const void* x, void* out, int ostride, int coff, int32_t* idx,
                                 void* stream) {
  if (mode < 0 || mode > 2) FFPN_FAIL(ctx, "resize2d: unknown mode %d", mode);
  if (mode == 0 && (Si != So || Wi != Wo)) FFPN_FAIL(ctx, "resize2d: copy mode needs equal sizes (%lld,%lld)!=(%lld,%lld)", (long long)Si, (long long)Wi, (long long)So, (long long)Wo);
  const int g = grid_of(ctx, B * So * Wo * C);
  if (dtype == FFPN_F32) ffpn_launch(resize2d_fwd_kernel<float>, g, TH, 0, (cudaStream_t)stream, mode, (int)B, (int)Si, (int)Wi, (int)So, (int)Wo, C, (const float*)x, (float*)out, ostride, coff, idx);
  else ffpn_launch(resize2d_fwd_kernel<bf16>, g, TH, 0, (cudaStream_t)stream, mode, (int)B, (int)Si, (int)Wi, (int)So, (int)Wo, C, (const bf16*)x, (bf16*)out, ostride, coff, idx);
  FFPN_CHECK_LAUNCH(ctx, "resize2d_fwd");
  return 0;
}

extern "C" int ffpn_resize2d_bwd(ffpn_ctx* ctx, int dtype, int mode, int64_t B, int64_t Si, int64_t Wi, int64_t So,
                                 int64_t Wo, int C, const void* dout, int ostride, int coff, const int32_t* idx, void* dx,
                                 void* stream) {
  if (mode < 0 || mode > 2) FFPN_FAIL(ctx, "resize2d: unknown mode %d", mode);
  if (mode == 1 && idx == nullptr) FFPN_FAIL(ctx, "resize2d_bwd: adaptive max needs the forward argmax");
  if (B * Si * Wi * C >= (1ll << 31)) FFPN_FAIL(ctx, "resize2d_bwd: more than 2^31 elements");
  {
    const int vec = dtype == FFPN_F32 ? 4 : 8;
    if (mode == 1 && Si % So == 0 && Wi % Wo == 0 && C % vec == 0 && ostride % vec == 0 && coff % vec == 0) {
      const int gv = grid_of(ctx, B * Si * Wi * (C / vec));
      if (dtype == FFPN_F32) ffpn_launch(resize2d_bwd_max_tiled_kernel<float>, gv, TH, 0, (cudaStream_t)stream, (int)B, (int)Si, (int)Wi, (int)So, (int)Wo, C, (const float*)dout, ostride, coff, idx, (float*)dx);
      else ffpn_launch(resize2d_bwd_max_tiled_kernel<bf16>, gv, TH, 0, (cudaStream_t)stream, (int)B, (int)Si, (int)Wi, (int)So, (int)Wo, C, (const bf16*)dout, ostride, coff, idx, (bf16*)dx);
      FFPN_CHECK_LAUNCH(ctx, "resize2d_bwd");
      return 0;
    }
  }
  const int g = grid_of(ctx, B * Si * Wi * C);
  if (dtype == FFPN_F32) ffpn_launch(resize2d_bwd_kernel<float>, g, TH, 0, (cudaStream_t)stream, mode, (int)B, (int)Si, (int)Wi, (int)So, (int)Wo, C, (const float*)dout, ostride, coff, idx, (float*)dx);
  else ffpn_launch(resize2d_bwd_kernel<bf16>, g, TH, 0, (cudaStream_t)stream, mode, (int)B, (int)Si, (int)Wi, (int)So, (int)Wo, C, (const bf16*)dout, ostride, coff, idx, (bf16*)dx);
  FFPN_CHECK_LAUNCH(ctx, "resize2d_bwd");
  return 0;
}

extern "C" int ffpn_upsample_fwd(ffpn_ctx* ctx, int dtype, int64_t B, int64_t Si, int64_t Wi, int fS, int fW, int C,
                                 const void* x, void* out, int ostride, int coff, void* stream) {
  if (fS < 1 || fW < 1) FFPN_FAIL(ctx, "upsample: integer factors >= 1 required");
  const int g = grid_of(ctx, B * Si * fS * Wi * fW * C);
  if (dtype == FFPN_F32) ffpn_launch(upsample_fwd_kernel<float>, g, TH, 0, (cudaStream_t)stream, (int)B, (int)Si, (int)Wi, fS, fW, C, (const float*)x, (float*)out, ostride, coff);
  else ffpn_launch(upsample_fwd_kernel<bf16>, g, TH, 0, (cudaStream_t)stream, (int)B, (int)Si, (int)Wi, fS, fW, C, (const bf16*)x, (bf16*)out, ostride, coff);
  FFPN_CHECK_LAUNCH(ctx, "upsample_fwd");
  return 0;
}

extern "C" int ffpn_upsample_bwd(ffpn_ctx* ctx, int dtype, int64_t B, int64_t Si, int64_t Wi, int fS, int fW, int C,
                                 const void* dout, int ostride, int coff, void* dx, void* stream) {
  if (fS < 1 || fW < 1) FFPN_FAIL(ctx, "upsample: integer factors >= 1 required");
  const int g = grid_of(ctx, B * Si * Wi * C);
  if (dtype == FFPN_F32) ffpn_launch(upsample_bwd_kernel<float>, g, TH, 0, (cudaStream_t)stream, (int)B, (int)Si, (int)Wi, fS, fW, C, (const float*)dout, ostride, coff, (float*)dx);
  else ffpn_launch(upsample_bwd_kernel<bf16>, g, TH, 0, (cudaStream_t)stream, (int)B, (int)Si, (int)Wi, fS, fW, C, (const bf16*)dout, ostride, coff, (bf16*)dx);
  FFPN_CHECK_LAUNCH(ctx, "upsample_bwd");
  return 0;
}

extern "C" int ffpn_slice_copy(ffpn_ctx* ctx, int dtype, int64_t P, int C, const void* src, int sstride, int soff,
                               void* dst, int dstride, int doff, void* stream) {
  const int g = grid_of(ctx, P * C);
  if (dtype == FFPN_F32) ffpn_launch(slice_copy_kernel<float>, g, TH, 0, (cudaStream_t)stream, P, C, (const float*)src, sstride, soff, (float*)dst, dstride, doff);
  else ffpn_launch(slice_copy_kernel<bf16>, g, TH, 0, (cudaStream_t)stream, P, C, (const bf16*)src, sstride, soff, (bf16*)dst, dstride, doff);
  FFPN_CHECK_LAUNCH(ctx, "slice_copy");
  return 0;
}

extern "C" int ffpn_head_fwd(ffpn_ctx* ctx, int dtype, int64_t B, int64_t EW, int C, int n, int act, const void* x, const float* w,
                             const float* bias, float* out, void* stream) {
  if (act != 0 && act != 1) FFPN_FAIL(ctx, "head_fwd: unknown activation %d", act);
  const int g = grid_of(ctx, B * EW);
  if (dtype == FFPN_F32) ffpn_launch(head_fwd_kernel<float>, g, TH, 0, (cudaStream_t)stream, (int)B, EW, C, n, act, (const float*)x, w, bias, out);
  else ffpn_launch(head_fwd_kernel<bf16>, g, TH, 0, (cudaStream_t)stream, (int)B, EW, C, n, act, (const bf16*)x, w, bias, out);
  FFPN_CHECK_LAUNCH(ctx, "head_fwd");
  return 0;
}

extern "C" size_t ffpn_head_bwd_workspace_bytes(int C, int n) { return (size_t)HB_BLOCKS * n * (C + 1) * sizeof(float); }

extern "C" int ffpn_head_bwd(ffpn_ctx* ctx, int dtype, int64_t B, int64_t EW, int C, int n, const void* x, const float* w,
                             const float* dout, const float* pred, void* dx, float* dw, float* dbias, void* ws, size_t ws_bytes,
                             void* stream) {
  if (C + 1 > TH) FFPN_FAIL(ctx, "head_bwd: more than %d input channels", TH - 1);
  if (ws == nullptr || ws_bytes < ffpn_head_bwd_workspace_bytes(C, n)) FFPN_FAIL(ctx, "head_bwd: workspace too small");
  const int g = grid_of(ctx, B * EW * C);
  if (dx != nullptr) {
    if (dtype == FFPN_F32) ffpn_launch(head_bwd_dx_kernel<float>, g, TH, 0, (cudaStream_t)stream, (int)B, EW, C, n, w, dout, pred, (float*)dx);
    else ffpn_launch(head_bwd_dx_kernel<bf16>, g, TH, 0, (cudaStream_t)stream, (int)B, EW, C, n, w, dout, pred, (bf16*)dx);
    FFPN_CHECK_LAUNCH(ctx, "head_bwd_dx");
  }
  const int lanes = TH / (C + 1);
  int nb = (int)((B * EW + lanes - 1) / lanes);
  if (nb > HB_BLOCKS) nb = HB_BLOCKS;
  if (dtype == FFPN_F32) ffpn_launch(head_bwd_dw_partial_kernel<float>, nb, TH, 0, (cudaStream_t)stream, (int)B, EW, C, n, (const float*)x, dout, pred, (float*)ws);
  else ffpn_launch(head_bwd_dw_partial_kernel<bf16>, nb, TH, 0, (cudaStream_t)stream, (int)B, EW, C, n, (const bf16*)x, dout, pred, (float*)ws);
  FFPN_CHECK_LAUNCH(ctx, "head_bwd_dw_partial");
  ffpn_launch(head_bwd_dw_final_kernel, 1, TH, 0, (cudaStream_t)stream, nb, C, n, (const float*)ws, dw, dbias);
  FFPN_CHECK_LAUNCH(ctx, "head_bwd_dw_final");
  return 0;
}

extern "C" int ffpn_pack_volume(ffpn_ctx* ctx, int dtype, int64_t R, int64_t H, int64_t W, const float* src, void* dst,
                                void* stream) {
  const int64_t ntile = R * ((W + 31) / 32) * ((H + 31) / 32);
  const int g = (int)(ntile < (int64_t)ctx->num_sms * 16 ? ntile : (int64_t)ctx->num_sms * 16);
  if (dtype == FFPN_F32) ffpn_launch(pack_volume_kernel<float>, g, TH, 0, (cudaStream_t)stream, R, (int)H, (int)W, src, (float*)dst);
  else ffpn_launch(pack_volume_kernel<bf16>, g, TH, 0, (cudaStream_t)stream, R, (int)H, (int)W, src, (bf16*)dst);
  FFPN_CHECK_LAUNCH(ctx, "pack_volume");
  return 0;
}

extern "C" int ffpn_cast(ffpn_ctx* ctx, int dtype, int64_t n, const float* src, void* dst, void* stream) {
  if (dtype == FFPN_F32) ffpn_launch(cast_kernel<float>, grid_of(ctx, n), TH, 0, (cudaStream_t)stream, n, src, (float*)dst);
  else ffpn_launch(cast_kernel<bf16>, grid_of(ctx, n), TH, 0, (cudaStream_t)stream, n, src, (bf16*)dst);
  FFPN_CHECK_LAUNCH(ctx, "cast");
  return 0;
}

extern "C" int ffpn_sgd_step(ffpn_ctx* ctx, int64_t n, float* p, const float* g, float* mom, float lr, float momentum,
                             float weight_decay, float grad_scale, int first_step, void* stream) {
  ffpn_launch(sgd_kernel, grid_of(ctx, n), TH, 0, (cudaStream_t)stream, n, p, g, mom, lr, momentum, weight_decay, grad_scale, first_step);
  FFPN_CHECK_LAUNCH(ctx, "sgd_step");
  return 0;
}
