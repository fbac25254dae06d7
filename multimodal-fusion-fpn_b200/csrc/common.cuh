// Shared device helpers and the ctx object of libfusionfpn.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/ffpn.h"

struct ffpn_ctx {
  int device;
  int num_sms;
  int64_t launches;
  char err[512];
};

#define FFPN_FAIL(ctx, ...)                                   \
  do {                                                        \
    if (ctx) snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__); \
    return 1;                                                 \
  } while (0)

#define FFPN_CHECK_LAUNCH(ctx, name)                                                   \
  do {                                                                                 \
    cudaError_t e__ = cudaGetLastError();                                              \
    if (e__ != cudaSuccess) FFPN_FAIL(ctx, "%s: launch failed: %s", name, cudaGetErrorString(e__)); \
    (ctx)->launches++;                                                                 \
  } while (0)

typedef __nv_bfloat16 bf16;

// ---- element access: VEC consecutive channels as floats ---------------------------------------------
template <typename T> struct Elem;
template <> struct Elem<float> {
  static constexpr int VEC = 4;  // 16 bytes
  __device__ __forceinline__ static void load(const float* p, float (&v)[4]) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ __forceinline__ static void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
  __device__ __forceinline__ static float ld1(const float* p) { return *p; }
  __device__ __forceinline__ static void st1(float* p, float v) { *p = v; }
  __device__ __forceinline__ static float rnd(float v) { return v; }
};
template <> struct Elem<bf16> {
  static constexpr int VEC = 8;  // 16 bytes
  __device__ __forceinline__ static void load(const bf16* p, float (&v)[8]) {
    uint4 t = *reinterpret_cast<const uint4*>(p);
    const uint32_t u[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
      v[2 * i] = __uint_as_float(u[i] << 16);
      v[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
    }
  }
  __device__ __forceinline__ static void store(bf16* p, const float (&v)[8]) {
    uint32_t u[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      u[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(u[0], u[1], u[2], u[3]);
  }
  __device__ __forceinline__ static float ld1(const bf16* p) { return __bfloat162float(*p); }
  __device__ __forceinline__ static void st1(bf16* p, float v) { *p = __float2bfloat16_rn(v); }
  __device__ __forceinline__ static float rnd(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// PyTorch max-pool comparison: candidate replaces best when (cand > best) || isnan(cand)
// (aten/src/ATen/native/cuda/DilatedMaxPool3d.cu semantics restated in SURVEY.md App. B).
__device__ __forceinline__ bool pool_better(float cand, float best) { return (cand > best) || (cand != cand); }

static inline int ffpn_grid_for(int64_t work_items, int per_block, int max_blocks) {
  int64_t g = (work_items + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return (int)g;
}
