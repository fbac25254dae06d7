// Helpers shared by the warp-specialised tcgen05 kernels (conv_ws.cu, conv_wgrad_ws.cu).
#pragma once
#include "tc_common.cuh"

static __device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
static __device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// K-major swizzled operand: rows `pitch` bytes apart (= swizzle width), 8-row groups at SBO = 8 * pitch.  Returns the
// high word; the low word is (addr >> 4) | (1 << 16).
static __device__ __forceinline__ uint32_t desc_sw_hi(uint32_t pitch) {
  const uint32_t ltype = pitch == 128 ? 2u : pitch == 64 ? 4u : 6u;
  return ((8u * pitch) >> 4) | (1u << 14) | (ltype << 29);
}
static __device__ __forceinline__ uint64_t desc64(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));   // pure register pairing, no arithmetic
  return d;
}

// Sum v[0..15] over the 32 lanes of a warp with a transpose-reduce (16 shuffles): afterwards lanes with even index hold,
// in v[0], the total of channel ch = 8*b4 + 4*b3 + 2*b2 + b1 (bN = bit N of the lane index).
static __device__ __forceinline__ float warp_transpose_sum16(float (&v)[16], int lane) {
#pragma unroll
  for (int step = 0; step < 4; step++) {
    const int half = 8 >> step, bit = 16 >> step;
    const bool upper = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < half; i++) {
      const float send = upper ? v[i] : v[i + half];
      const float keep = upper ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}
static __device__ __forceinline__ int transpose_sum_channel(int lane) {
  return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}

static __device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}

