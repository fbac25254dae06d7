// Hardware probe for the layout questions behind the conv kernels (sm_100a):
//  1. K-major SWIZZLE_32B/64B/128B A operand whose descriptor start address is shifted by s rows (not a multiple of
//     the 8-row swizzle atom): are the right rows read, and which base_offset does the descriptor need?
//  2. MN-major SWIZZLE_32B A operand whose "slabs" (LBO) overlap: slab m = the same tile shifted by m*Ls rows.
//  3. Issue cost (cycles per tcgen05.mma) of the small-N shapes the narrow layers use.
// Build + run on the GPU box:  nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/umma_probe tools/umma_probe.cu && /tmp/umma_probe
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
               "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct Probe {
  int mode;          // 0: K-major swizzled A with row shift; 1: MN-major SW32 A with overlapping slabs; 2: timing
  int sw;            // swizzle bytes 32/64/128 (row pitch of the tile)
  int shift;         // start-address shift in rows
  int k0;            // first channel of the K=16 slice (mode 0)
  int bo_mode;       // base_offset: 0 -> 0, 1 -> (start >> 7) & 7
  int Ls;            // slab stride in rows (mode 1)
  int M, N;          // MMA shape
  int nmma;          // timing: MMAs per commit
  int nacc, run;     // timing: cycle over nacc accumulator tiles, switching every `run` MMAs
  int tshape;        // timing operand layout: 0 no-swizzle planar K-major, 1 SW K-major, 2 MN-major SW32 both
};

__host__ __device__ inline float aval(int r, int c) { return (float)(((r * 5 + c * 3) % 61) - 30); }

// 128 threads.  out[m][n] fp32 (M x 16 or M x N for mode 1), cyc[0] = cycles
__global__ void __launch_bounds__(128) probe_kernel(Probe p, float* out, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* a_s = smem;                       // up to 96 KB
  uint8_t* b_s = smem + 96 * 1024;           // 16 KB
  const uint32_t abase = smem_u32(a_s), bbase = smem_u32(b_s);
  // ---- fill A: row-major tile [R][sw/2 channels], swizzled by absolute address bits ----
  const int mask = p.sw == 128 ? 7 : p.sw == 64 ? 3 : p.sw == 32 ? 1 : 0;
  const int R = 96 * 1024 / (p.sw ? p.sw : 32);
  const int C = (p.sw ? p.sw : 32) / 2;
  if (p.mode != 2 || p.tshape != 0) {
    for (int i = tid; i < R * C; i += 128) {
      const int r = i / C, c = i - r * C;
      uint32_t off = (uint32_t)r * (C * 2) + c * 2;
      uint32_t a = abase + off;
      a ^= ((a >> 7) & mask) << 4;
      *reinterpret_cast<bf16*>(a_s + (a - abase)) = __float2bfloat16(aval(r, c));
    }
  } else {
    for (int i = tid; i < 96 * 1024 / 2; i += 128) reinterpret_cast<bf16*>(a_s)[i] = __float2bfloat16(1.f);
  }
  // ---- fill B: K-major no-swizzle selection matrix B[n][k] = (n == k), N x 16: addr = (n/8)*128 + kchunk*(N*16) + (n%8)*16 ----
  for (int i = tid; i < 16 * 1024 / 2; i += 128) reinterpret_cast<bf16*>(b_s)[i] = __float2bfloat16(0.f);
  __syncthreads();
  if (p.mode != 2) {
    for (int n = tid; n < p.N; n += 128) {
      const int k = n % 16;
      const uint32_t off = (n / 8) * 128 + (k / 8) * (p.N * 16) + (n % 8) * 16 + (k % 8) * 2;
      *reinterpret_cast<bf16*>(b_s + off) = __float2bfloat16(1.f);
    }
  }
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const uint64_t ltype = p.sw == 128 ? 2ull : p.sw == 64 ? 4ull : p.sw == 32 ? 6ull : 0ull;
  const uint64_t bdesc = (uint64_t)((bbase & 0x3FFFFu) >> 4) | ((uint64_t)((p.N * 16) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
  if (tid == 0) {
    if (p.mode == 0) {
      const uint32_t start = abase + (uint32_t)p.shift * p.sw + (uint32_t)p.k0 * 2;
      const uint64_t bo = p.bo_mode ? (uint64_t)((start >> 7) & 7) : 0ull;
      const uint64_t adesc = (uint64_t)((start & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)((8 * p.sw) >> 4) << 32) | (1ull << 46) |
                             (bo << 49) | (ltype << 61);
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(p.M >> 4) << 24);
      umma_bf16(tmem, adesc, bdesc, idesc, 0u);
      umma_commit(smem_u32(&bar));
    } else if (p.mode == 1) {
      // MN-major SW32: LBO = slab stride (bytes), SBO = 8 K-rows = 256 B
      const uint32_t start = abase + (uint32_t)p.shift * 32u;
      const uint64_t bo = p.bo_mode ? (uint64_t)((start >> 7) & 7) : 0ull;
      const uint64_t adesc = (uint64_t)((start & 0x3FFFFu) >> 4) | ((uint64_t)(((uint32_t)p.Ls * 32u) >> 4) << 16) |
                             ((uint64_t)(256 >> 4) << 32) | (1ull << 46) | (bo << 49) | (6ull << 61);
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(p.M >> 4) << 24);
      umma_bf16(tmem, adesc, bdesc, idesc, 0u);
      umma_commit(smem_u32(&bar));
    } else {
      uint64_t adesc, bd = bdesc;
      uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(p.M >> 4) << 24);
      if (p.tshape == 0) {          // planar no-swizzle K-major: LBO = plane (rows*16), SBO = 128
        adesc = (uint64_t)((abase & 0x3FFFFu) >> 4) | ((uint64_t)((1024 * 16) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
      } else if (p.tshape == 1) {   // swizzled K-major
        adesc = (uint64_t)((abase & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)((8 * p.sw) >> 4) << 32) | (1ull << 46) | (ltype << 61);
      } else {                      // MN-major SW32 for A (slabs Ls rows apart) and B (N/16 slabs, LBO = Ls rows too)
        adesc = (uint64_t)((abase & 0x3FFFFu) >> 4) | ((uint64_t)(((uint32_t)p.Ls * 32u) >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) |
                (1ull << 46) | (6ull << 61);
        bd = (uint64_t)(((abase + 32768u) & 0x3FFFFu) >> 4) | ((uint64_t)(((uint32_t)p.Ls * 32u) >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) |
             (1ull << 46) | (6ull << 61);
        idesc |= (1u << 15) | (1u << 16);
      }
      const long long t0 = clock64();
      const int nacc = p.nacc > 0 ? p.nacc : 1, run = p.run > 0 ? p.run : 1;
      for (int i = 0; i < p.nmma; i++) {
        const int a = (i / run) % nacc;
        umma_bf16(tmem + (uint32_t)(a * 32), adesc + (uint64_t)((i & 7) * 2), bd, idesc, i >= nacc * run ? 1u : 0u);
      }
      umma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), 0);
      const long long t1 = clock64();
      cyc[0] = t1 - t0;
    }
  }
  mbar_wait(smem_u32(&bar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (p.mode != 2) {
    const int nblk = p.M / 32;     // lanes: M=128 -> 4 warps; M=64 -> warps 0,1 hold lanes 0-31, 32-63 (layout D for M=64: lanes 0..63)
    if (warp < 4) {
      for (int ch = 0; ch < p.N / 16; ch++) {
        uint32_t raw[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + ch * 16, raw);
        (void)nblk;
          for (int q = 0; q < 16; q++) out[(size_t)(warp * 32 + lane) * p.N + ch * 16 + q] = __uint_as_float(raw[q]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u));
  }
}

// Mode 3: hardware MMA rate with a lean, warp-uniform issue loop (8 MMAs unrolled, compile-time accumulator rotation).
template <int NACC>
__global__ void __launch_bounds__(128) rate_kernel(int sw, int M, int N, int nmma, int astep, long long* cyc, int mn = 0, int lboA = 0, int lboB = 0, int swB = 0) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 112 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3f803f80u;   // bf16 1.0
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    const uint32_t abase = smem_u32(smem), bbase = abase + 96 * 1024;
    const uint64_t ltype = sw == 128 ? 2ull : sw == 64 ? 4ull : sw == 32 ? 6ull : 0ull;
    const uint64_t adesc0 = sw ? ((uint64_t)((abase & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)((8 * sw) >> 4) << 32) | (1ull << 46) | (ltype << 61))
                               : ((uint64_t)((abase & 0x3FFFFu) >> 4) | ((uint64_t)((1024 * 16) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46));
    const uint64_t bdesc = (uint64_t)((bbase & 0x3FFFFu) >> 4) | ((uint64_t)((N * 16) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
    uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t cs = N < 32 ? 32 : N;
    uint64_t adesc0m = adesc0, bdescm = bdesc;
    if (mn) {   // both operands MN-major swizzled: LBO = slab stride, SBO = 8 rows
      const uint64_t ltB = swB == 128 ? 2ull : swB == 64 ? 4ull : 6ull;
      adesc0m = (uint64_t)((abase & 0x3FFFFu) >> 4) | ((uint64_t)((uint32_t)lboA >> 4) << 16) | ((uint64_t)((8 * sw) >> 4) << 32) | (1ull << 46) | (ltype << 61);
      bdescm = (uint64_t)(((abase + 49152u) & 0x3FFFFu) >> 4) | ((uint64_t)((uint32_t)lboB >> 4) << 16) | ((uint64_t)((8 * swB) >> 4) << 32) | (1ull << 46) | (ltB << 61);
      idesc |= (1u << 15) | (1u << 16);
    }
    const long long t0 = clock64();
    for (int i = 0; i < nmma; i += 8) {
#pragma unroll
      for (int j = 0; j < 8; j++) {
        if (pred) umma_bf16(tmem + (uint32_t)((j % NACC) * cs), adesc0m + (uint64_t)(j * astep), bdescm + (uint64_t)(mn ? j * astep : 0), idesc, i ? 1u : 0u);
      }
    }
    if (pred) umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t1 = clock64();
    if (pred) cyc[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
  }
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 128 * 256 * 4);
  cudaMalloc(&cyc, 8);
  const size_t smem = 96 * 1024 + 16 * 1024 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  static float h[128 * 256];
  const bool fast = getenv("PROBE_FAST") != nullptr;
  printf("== mode 0: K-major swizzled A, start shifted by s rows (M=128, N=16, K=16) ==\n");
  for (int sw : {32, 64, 128}) {
    if (fast) break;
    for (int bo = 0; bo < 2; bo++) {
      for (int k0 = 0; k0 < sw / 2; k0 += 16) {
        printf("sw %3d base_offset_mode %d k0 %2d: ", sw, bo, k0);
        for (int s = 0; s < 20; s++) {
          Probe p; memset(&p, 0, sizeof(p));
          p.mode = 0; p.sw = sw; p.shift = s; p.k0 = k0; p.bo_mode = bo; p.M = 128; p.N = 16;
          cudaMemset(out, 0, 128 * 256 * 4);
          probe_kernel<<<1, 128, smem>>>(p, out, cyc);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
          cudaMemcpy(h, out, 128 * 16 * 4, cudaMemcpyDeviceToHost);
          int bad = 0;
          for (int m = 0; m < 128; m++) for (int n = 0; n < 16; n++) if (h[m * 16 + n] != aval(m + s, k0 + n)) bad++;
          printf("%s", bad ? "x" : ".");
        }
        printf("\n");
      }
    }
  }
  printf("== mode 1: MN-major SW32 A (tile [pos][16ch]), M = 16ch x (M/16) slabs at LBO = Ls rows, start shifted by s rows, K = 16 positions ==\n");
  for (int M : {128, 64}) {
    if (fast) break;
    for (int Ls : {1, 2, 3, 8, 18, 33, 130}) {
      for (int bo = 0; bo < 2; bo++) {
        printf("M %3d Ls %3d base_offset_mode %d: ", M, Ls, bo);
        for (int s = 0; s < 12; s++) {
          Probe p; memset(&p, 0, sizeof(p));
          p.mode = 1; p.sw = 32; p.shift = s; p.bo_mode = bo; p.Ls = Ls; p.M = M; p.N = 16;
          cudaMemset(out, 0, 128 * 256 * 4);
          probe_kernel<<<1, 128, smem>>>(p, out, cyc);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
          cudaMemcpy(h, out, 128 * 16 * 4, cudaMemcpyDeviceToHost);
          int bad = 0;
          // D[(slab*16 + c)][n] = X[n + slab*Ls + s][c]
          int bad2 = 0;
          for (int mm = 0; mm < M; mm++) for (int n = 0; n < 16; n++) {
            const int slab = mm / 16, c = mm % 16;
            if (h[mm * 16 + n] != aval(n + slab * Ls + s, c)) bad++;
            const int lane2 = (mm / 16) * 32 + (mm % 16);       // alternative M=64 lane mapping
            if (h[lane2 * 16 + n] != aval(n + slab * Ls + s, c)) bad2++;
          }
          printf("%s", !bad ? "." : (M == 64 && !bad2) ? "o" : "x");
        }
        printf("\n");
      }
    }
  }
  printf("== mode 2: cycles per MMA (1 CTA, back-to-back issue, K=16) ==\n");
  struct T { int M, N, tshape, sw, Ls; const char* name; };
  const T ts[] = {{128, 16, 0, 0, 0, "M128 N16  planar no-swizzle K-major"}, {128, 16, 1, 32, 0, "M128 N16  SW32 K-major"},
                  {128, 32, 1, 64, 0, "M128 N32  SW64 K-major"}, {128, 64, 1, 128, 0, "M128 N64  SW128 K-major"},
                  {128, 128, 1, 128, 0, "M128 N128 SW128 K-major"}, {128, 256, 1, 128, 0, "M128 N256 SW128 K-major"},
                  {64, 48, 2, 32, 18, "M64  N48  MN-major SW32 both"}, {128, 48, 2, 32, 18, "M128 N48  MN-major SW32 both"},
                  {64, 16, 2, 32, 18, "M64  N16  MN-major SW32 both"}, {64, 96, 2, 32, 18, "M64  N96  MN-major SW32 both"},
                  {64, 16, 1, 32, 0, "M64  N16  SW32 K-major"}, {64, 256, 1, 128, 0, "M64 N256 SW128 K-major"}};
  for (const T& t : ts) {
    if (fast) break;
    for (int nm : {256, 2048}) {
      Probe p; memset(&p, 0, sizeof(p));
      p.mode = 2; p.sw = t.sw; p.M = t.M; p.N = t.N; p.nmma = nm; p.tshape = t.tshape; p.Ls = t.Ls;
      probe_kernel<<<1, 128, smem>>>(p, out, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
      long long c;
      cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("%-40s nmma %5d: %8lld cycles = %.1f cyc/MMA\n", t.name, nm, c, (double)c / nm);
    }
  }
  printf("== mode 2b: accumulator switching (M128 N16 SW32 K-major, 2048 MMAs) ==\n");
  for (int nacc : {1, 2, 4, 8}) {
    if (fast) break;
    for (int run : {1, 3, 9, 36}) {
      Probe p; memset(&p, 0, sizeof(p));
      p.mode = 2; p.sw = 32; p.M = 128; p.N = 16; p.nmma = 2048; p.tshape = 1; p.nacc = nacc; p.run = run;
      probe_kernel<<<1, 128, smem>>>(p, out, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
      long long c;
      cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("nacc %d run %2d: %.1f cyc/MMA\n", nacc, run, (double)c / 2048);
    }
  }
  printf("== mode 3: MMA rate, lean uniform issue (cycles per MMA; astep = A start shift per MMA in 16-byte units) ==\n");
  cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(rate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(rate_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  struct R { int sw, M, N, astep; const char* name; };
  const R rs[] = {{32, 128, 16, 0, "M128 N16 SW32 same rows"}, {32, 128, 16, 2, "M128 N16 SW32 shift 1 row/MMA"}, {32, 128, 16, 260, "M128 N16 SW32 shift 130 rows/MMA"},
                  {0, 128, 16, 0, "M128 N16 no-swizzle planar"}, {0, 128, 16, 1, "M128 N16 no-swizzle planar shift 1 row"},
                  {64, 128, 32, 0, "M128 N32 SW64"}, {64, 128, 32, 4, "M128 N32 SW64 shift 1 row"},
                  {128, 128, 64, 0, "M128 N64 SW128"}, {128, 128, 64, 8, "M128 N64 SW128 shift 1 row"}, {128, 128, 64, 2, "M128 N64 SW128 K-advance 32B"},
                  {128, 128, 128, 8, "M128 N128 SW128 shift 1 row"}, {128, 128, 256, 8, "M128 N256 SW128 shift 1 row"},
                  {32, 64, 16, 2, "M64 N16 SW32 shift 1 row"}, {128, 64, 64, 8, "M64 N64 SW128 shift 1 row"}};
  for (const R& r : rs) {
    printf("%-36s:", r.name);
    for (int nacc : {1, 2, 8}) {
      if (nacc * (r.N < 32 ? 32 : r.N) > 512) { printf("  nacc %d: -", nacc); continue; }
      if (nacc == 1) rate_kernel<1><<<1, 128, smem>>>(r.sw, r.M, r.N, 4096, r.astep, cyc);
      else if (nacc == 2) rate_kernel<2><<<1, 128, smem>>>(r.sw, r.M, r.N, 4096, r.astep, cyc);
      else rate_kernel<8><<<1, 128, smem>>>(r.sw, r.M, r.N, 4096, r.astep, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
      long long c;
      cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("  nacc %d: %6.1f", nacc, (double)c / 4096);
    }
    printf("\n");
  }
  printf("== mode 3b: MN-major (wgrad) shapes, cycles per MMA ==\n");
  struct Q { int sw, M, N, astep, lboA, lboB, swB; const char* name; };
  const Q qs[] = {{32, 64, 48, 32, 32, 130 * 32, 32, "M64 N48 SW32/SW32 (L1: 4 dx slabs x 3 line slabs)"},
                  {32, 64, 16, 32, 32, 130 * 32, 32, "M64 N16 SW32/SW32"},
                  {32, 128, 48, 32, 32, 130 * 32, 32, "M128 N48 SW32/SW32"},
                  {64, 128, 96, 64, 64, 66 * 64, 64, "M128 N96 SW64/SW64 (L2)"},
                  {128, 128, 192, 128, 128, 34 * 128, 128, "M128 N192 SW128/SW128 (L3)"},
                  {128, 128, 64, 128, 128, 34 * 128, 128, "M128 N64 SW128/SW128"},
                  {128, 128, 128, 128, 16384, 16384, 128, "M128 N128 SW128 sub-tile slabs (L4)"},
                  {128, 128, 256, 128, 16384, 8192, 128, "M128 N256 SW128 sub-tile slabs (L5)"}};
  for (const Q& q : qs) {
    rate_kernel<1><<<1, 128, smem>>>(q.sw, q.M, q.N, 4096, q.astep, cyc, 1, q.lboA, q.lboB, q.swB);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-52s: %6.1f\n", q.name, (double)c / 4096);
  }
  return 0;
}
