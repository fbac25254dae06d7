// tcgen05 / TMEM implicit-GEMM convolution for sm_100a (forward and, through flipped/transposed packed
// weights, dgrad) for every stride-1 kernel shape on the path.
//
// Formulation.  The conv is canonicalised to (slices D with kD taps) x (inner plane Y x X with kY x kX taps,
// X contiguous).  Positions of the zero-padded inner plane are flattened, i = y*Xp + x, so that for an output
// row m the input row of tap (dd,dy,dx) is simply m + dd*Lr + dy*Xp + dx: every tap's A operand is the SAME
// shared-memory tile viewed through a shifted UMMA descriptor start address.  The tile is staged once per
// K-group by all threads with 16-byte global loads; the producer's BatchNorm scale/shift (+ReLU) and the zero
// padding are applied in registers on the way in, so the normalised tensor never exists in HBM.  A (and B)
// use the canonical K-major, no-swizzle UMMA layout: 8-channel (16-byte) chunks in planes, rows 16 bytes apart
// (SBO = 128 B, LBO = plane stride), which is what makes arbitrary row shifts legal.  Weights are packed per
// call to bf16 in exactly the smem image ([kgroup][tap][kchunk][cout][8]) and fetched with one
// cp.async.bulk (UBLKCP) per K-group, completion on an mbarrier.  One elected thread issues
// tcgen05.mma.cta_group::1.kind::f16 (M=128, N=Cout, K=16) into TMEM accumulators (one 128-row block per
// column group); tcgen05.commit signals an mbarrier; 8 warps read the accumulators back with tcgen05.ld,
// round, store channels-last rows and reduce the BatchNorm partial sums.  Two CTAs per SM overlap one CTA's
// staging/epilogue with the other's MMAs.


#include "tc_common.cuh"

namespace {

// Walks the rows r0, r0+RP, ... of a staged tile, tracking (slice j, padded-flat inner index -> yp, xp)
// incrementally so that no per-element division is needed.
struct RowWalk {
  int r, j, ii, yp, xp;
  __device__ __forceinline__ void init(int r0, int i0, const TcParams& p) {
    r = r0; j = r / p.Lr; ii = r - j * p.Lr;
    yp = (i0 + ii) / p.Xp; xp = (i0 + ii) - yp * p.Xp;
  }
  __device__ __forceinline__ void advance(int RP, int i0, const TcParams& p) {
    r += RP; ii += RP; xp += RP;
    if (ii >= p.Lr) {
      while (ii >= p.Lr) { ii -= p.Lr; j++; }
      yp = (i0 + ii) / p.Xp; xp = (i0 + ii) - yp * p.Xp;
    } else {
      while (xp >= p.Xp) { xp -= p.Xp; yp++; }
    }
  }
  __device__ __forceinline__ bool source(int d0, int res, const TcParams& p, long long& pos) const {
    const int d = d0 + j - p.pD, yy = yp - p.pY;
    const int xx = p.sX == 1 ? xp - p.hl : p.sX * (xp - p.hl) + res;
    pos = (long long)d * p.inD + (long long)yy * p.inY + xx;
    return d >= 0 && d < p.D && yy >= 0 && yy < p.Y && xx >= 0 && xx < p.X;
  }
};

struct TileCoord {
  int nb, d0, i0, tD_t, L_t, M_t, nmb;
  __device__ __forceinline__ void set(int tile, const TcParams& p) {
    int t = tile;
    const int it = t % p.nI; t /= p.nI;
    const int dt = t % p.nD;
    nb = t / p.nD;
    d0 = dt * p.tD; i0 = it * p.L;
    tD_t = min(p.tD, p.oD - d0);
    L_t = min(p.L, p.Qout - i0);
    M_t = (tD_t - 1) * p.Lr + L_t;
    nmb = (M_t + 127) >> 7;
  }
};

// Issue the asynchronous copies of one (tile, K-group) straight into the planar A layout (padding zero-filled).
__device__ __forceinline__ void stage_issue(const TcParams& p, const TileCoord& tc, int kg, uint32_t a_dst, uint32_t plane,
                                            int tid) {
  const int nkc = p.KG >> 3, kc = tid % nkc, RP = TC_THREADS / nkc;
  const bf16* xb = p.x + (long long)tc.nb * p.inNB * p.Cin + kg * p.KG + kc * 8;
  for (int set = 0; set < p.nsets; set++) {
    const int res = p.nsets == 1 ? (p.sX == 1 ? 0 : ((-p.pX) % p.sX + p.sX) % p.sX) : set;
    const uint32_t dst = a_dst + (uint32_t)(set * nkc + kc) * plane;
    RowWalk w;
    w.init(tid / nkc, tc.i0, p);
    while (w.r < p.region_rows) {
      long long pos;
      const bool ok = w.source(tc.d0, res, p, pos);
      cp_async16(dst + (uint32_t)w.r * 16u, ok ? (const void*)(xb + pos * p.Cin) : (const void*)p.x, ok);
      w.advance(RP, tc.i0, p);
    }
  }
}

// In-place BN scale/shift + ReLU of the valid elements of a staged tile (padding stays zero).
__device__ __forceinline__ void transform_inplace(const TcParams& p, const TileCoord& tc, int kg, uint8_t* a_buf, uint32_t plane,
                                                  int tid) {
  const int nkc = p.KG >> 3, kc = tid % nkc, RP = TC_THREADS / nkc;
  const int cofs = kg * p.KG + kc * 8;
  float s[8], h[8];
  {
    const float4 s0 = *reinterpret_cast<const float4*>(p.sc + cofs), s1 = *reinterpret_cast<const float4*>(p.sc + cofs + 4);
    const float4 h0 = *reinterpret_cast<const float4*>(p.sh + cofs), h1 = *reinterpret_cast<const float4*>(p.sh + cofs + 4);
    s[0] = s0.x; s[1] = s0.y; s[2] = s0.z; s[3] = s0.w; s[4] = s1.x; s[5] = s1.y; s[6] = s1.z; s[7] = s1.w;
    h[0] = h0.x; h[1] = h0.y; h[2] = h0.z; h[3] = h0.w; h[4] = h1.x; h[5] = h1.y; h[6] = h1.z; h[7] = h1.w;
  }
  for (int set = 0; set < p.nsets; set++) {
    const int res = p.nsets == 1 ? (p.sX == 1 ? 0 : ((-p.pX) % p.sX + p.sX) % p.sX) : set;
    uint8_t* base = a_buf + (size_t)(set * nkc + kc) * plane;
    RowWalk w;
    w.init(tid / nkc, tc.i0, p);
    while (w.r < p.region_rows) {
      long long pos;
      if (w.source(tc.d0, res, p, pos)) {
        uint4* q = reinterpret_cast<uint4*>(base + (size_t)w.r * 16);
        *q = bn_relu_bf16x8(*q, s, h, p.relu);
      }
      w.advance(RP, tc.i0, p);
    }
  }
}

// TMA staging of one (tile, K-group): one 4-D box per 8-channel chunk lands as one plane of the A layout; halo /
// out-of-image elements are filled by the TMA unit (zero, or NaN when a BN+ReLU transform follows: relu(NaN) = 0).
__device__ __forceinline__ void stage_issue_tma(const TcParams& p, const CUtensorMap* tmap, const TileCoord& tc, int kg,
                                                uint32_t a_dst, uint32_t plane, uint32_t bar) {
  const int nkc = p.KG >> 3;
  int c1, c2, c3;
  if (p.tma_mode == 0) { c1 = -p.hl; c2 = tc.i0 / p.Xp - p.pY; c3 = tc.d0; }
  else if (p.tma_mode == 1) { c1 = tc.i0; c2 = tc.d0 - p.pD; c3 = tc.nb; }
  else { c1 = 0; c2 = tc.i0 >> 8; c3 = 0; }
  mbar_expect_tx(bar, p.tma_bytes);
  for (int kc = 0; kc < nkc; kc++) tma_load_4d(a_dst + (uint32_t)kc * plane, tmap, kg * p.KG + kc * 8, c1, c2, c3, bar);
}

// In-place BN scale/shift + ReLU over a TMA-staged tile: a flat loop, no coordinates (NaN-filled halo -> 0).
__device__ __forceinline__ void transform_flat(const TcParams& p, int kg, uint8_t* a_buf, uint32_t plane, int tid) {
  const int nkc = p.KG >> 3, kc = tid % nkc, RP = TC_THREADS / nkc;
  const int cofs = kg * p.KG + kc * 8;
  float s[8], h[8];
  {
    const float4 s0 = *reinterpret_cast<const float4*>(p.sc + cofs), s1 = *reinterpret_cast<const float4*>(p.sc + cofs + 4);
    const float4 h0 = *reinterpret_cast<const float4*>(p.sh + cofs), h1 = *reinterpret_cast<const float4*>(p.sh + cofs + 4);
    s[0] = s0.x; s[1] = s0.y; s[2] = s0.z; s[3] = s0.w; s[4] = s1.x; s[5] = s1.y; s[6] = s1.z; s[7] = s1.w;
    h[0] = h0.x; h[1] = h0.y; h[2] = h0.z; h[3] = h0.w; h[4] = h1.x; h[5] = h1.y; h[6] = h1.z; h[7] = h1.w;
  }
  uint4* base = reinterpret_cast<uint4*>(a_buf + (size_t)kc * plane);
  for (int r = tid / nkc; r < p.region_rows; r += RP) base[r] = bn_relu_bf16x8(base[r], s, h, 1);
}

// TMEM accumulators of one tile -> bf16 rows in global memory (+ residual addend, + BatchNorm partial sums).
__device__ __forceinline__ void epilogue_tile(const TcParams& p, const TileCoord& tc, uint32_t tmem_tile, int n0, float* myscr,
                                              float* stat_s, int warp, int lane) {
  const int nchunks = p.Npad >> 4;
  for (int mb = warp >> 2; mb < tc.nmb; mb += 2) {
    const int m = mb * 128 + (warp & 3) * 32 + lane;
    const int j = m / p.Lr, ii = m - j * p.Lr;
    const int i = tc.i0 + ii;
    const int oy = i / p.Xp, ox = i - oy * p.Xp;
    const bool valid = (m < tc.M_t) && (j < tc.tD_t) && (ii < tc.L_t) && (oy < p.oY) && (ox < p.oX);
    const long long opos = (long long)tc.nb * p.outNB + (long long)(tc.d0 + j) * p.outD + (long long)oy * p.outY + ox;
    for (int ch = 0; ch < nchunks; ch++) {
      uint32_t raw[16];
      tmem_ld16(tmem_tile + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(mb * p.colstride + ch * 16), raw);
      float v[16];
#pragma unroll
      for (int q = 0; q < 16; q++) v[q] = __uint_as_float(raw[q]);
      const int cbase = n0 + ch * 16;                 // global output channel of this 16-wide chunk
      if (p.has_add && valid) {
        const uint4* ap = reinterpret_cast<const uint4*>(p.addend + opos * p.Cout + cbase);
#pragma unroll
        for (int h2 = 0; h2 < 2; h2++) {
          if (cbase + h2 * 8 < p.Cout) {
            const uint4 a4 = ap[h2];
            const uint32_t u[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
            for (int q = 0; q < 4; q++) {
              v[h2 * 8 + 2 * q] += __uint_as_float(u[q] << 16);
              v[h2 * 8 + 2 * q + 1] += __uint_as_float(u[q] & 0xffff0000u);
            }
          }
        }
      }
      uint32_t packed[8];
#pragma unroll
      for (int q = 0; q < 8; q++) {
        __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * q], v[2 * q + 1]);
        packed[q] = *reinterpret_cast<uint32_t*>(&hh);
        v[2 * q] = __uint_as_float(packed[q] << 16);           // statistics of the stored (rounded) values
        v[2 * q + 1] = __uint_as_float(packed[q] & 0xffff0000u);
      }
      if (valid) {
        uint4* yp = reinterpret_cast<uint4*>(p.y + opos * p.Cout + cbase);
        if (cbase < p.Cout) yp[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        if (cbase + 8 < p.Cout) yp[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
      }
      if (p.has_stats) {
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 16; q++) myscr[lane * SCR_STRIDE + q] = valid ? v[q] : 0.f;
        __syncwarp();
        const int c = lane & 15, half = lane >> 4;
        float sm = 0.f, s2 = 0.f;
#pragma unroll
        for (int rr = 0; rr < 16; rr++) {
          const float f = myscr[(half * 16 + rr) * SCR_STRIDE + c];
          sm += f;
          s2 = fmaf(f, f, s2);
        }
        sm += __shfl_xor_sync(0xffffffffu, sm, 16);
        s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
        if (lane < 16) {
          float* slot = stat_s + (warp & 7) * 2 * p.Npad;     // one slot per warp: no atomics, fixed summation order
          slot[ch * 16 + c] += sm;
          slot[p.Npad + ch * 16 + c] += s2;
        }
      }
    }
  }
}

// Software pipeline over units u = (tile, K-group) of a persistent CTA:
//   loads of unit u+1 (cp.async, straight into the other A buffer)  ||  transform + MMAs of unit u  ||
//   epilogue of the previous tile (other TMEM buffer).
__global__ void __launch_bounds__(TC_THREADS) conv_tc_kernel(const __grid_constant__ TcParams p,
                                                             const __grid_constant__ CUtensorMap tmap) {
  pdl_prologue();
  extern __shared__ __align__(128) uint8_t smem[];
  // header: [0],[8] weight barriers, [16] mma barrier, [24] tmem base, [32] tap offsets, [144],[152] TMA barriers, [160..] stats
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 24);
  int* tapoff_s = reinterpret_cast<int*>(smem + HDR_TAPS);
  float* stat_s = reinterpret_cast<float*>(smem + HDR_STATS);
  const uint32_t hdr = (HDR_STATS + 16 * p.Npad * 4 + 127) & ~127u;
  float* scr = reinterpret_cast<float*>(smem + hdr);                            // [8 warps][32][SCR_STRIDE]
  const uint32_t a_off = hdr + 8 * 32 * SCR_STRIDE * 4;
  uint8_t* a_s[2] = {smem + a_off, smem + a_off + p.a_bytes};
  uint8_t* b_s[2] = {smem + a_off + 2 * (size_t)p.a_bytes, smem + a_off + 2 * (size_t)p.a_bytes + (p.nkg > 1 ? p.b_bytes : 0)};
  const uint32_t bar_w[2] = {smem_u32(&bars[0]), smem_u32(&bars[1])}, bar_m = smem_u32(&bars[2]);
  const uint32_t bar_a[2] = {smem_u32(smem + 144), smem_u32(smem + 152)};
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ntaps = p.kD * p.kY * p.kX;
  const int nkc = p.KG >> 3;
  const uint32_t plane = (uint32_t)p.rows_alloc * 16u;
  const int n0 = blockIdx.y * p.Npad;                 // first output channel of this CTA's N-chunk
  const uint8_t* wp = reinterpret_cast<const uint8_t*>(p.wp) + (size_t)blockIdx.y * p.nkg * p.b_bytes;

  if (tid == 0) {
    mbar_init(bar_w[0], 1);
    mbar_init(bar_w[1], 1);
    mbar_init(bar_m, N_ISSUE);
    mbar_init(bar_a[0], 1);
    mbar_init(bar_a[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < ntaps) {
    // tap -> (residue set, row offset): input x = sX*ox + dx - pX = sX*(ox + q) + res, slot = ox + q + hl
    const int dx = tid % p.kX, dy = (tid / p.kX) % p.kY, dd = tid / (p.kX * p.kY);
    const int e = dx - p.pX;
    const int q = e >= 0 ? e / p.sX : -((-e + p.sX - 1) / p.sX);
    const int res = e - q * p.sX;
    const int set = p.nsets == 1 ? 0 : res;
    tapoff_s[tid] = set * nkc * p.rows_alloc + dd * p.Lr + dy * p.Xp + q + p.hl;
  }
  for (int i = tid; i < 16 * p.Npad; i += TC_THREADS) stat_s[i] = 0.f;   // [8 warps][2][Npad]
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)p.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_half = (uint32_t)(p.tmem_cols >> 1);     // two accumulator buffers (tile parity)

  // instruction descriptor: D=f32, A=B=bf16, K-major both, N=Npad, M=128
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.Npad >> 3) << 17) | (8u << 24);
  const uint64_t a_desc[2] = {make_desc(smem_u32(a_s[0]), plane, 128u), make_desc(smem_u32(a_s[1]), plane, 128u)};
  const uint64_t b_desc[2] = {make_desc(smem_u32(b_s[0]), (uint32_t)p.Npad * 16u, 128u),
                              make_desc(smem_u32(b_s[1]), (uint32_t)p.Npad * 16u, 128u)};
  float* myscr = scr + warp * 32 * SCR_STRIDE;

  const int ntiles = p.NB * p.nD * p.nI;
  const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int nunits = my_tiles * p.nkg;
  TileCoord tc_cur, tc_prev, tc_next;

  // prologue: unit 0
  tc_cur.set(blockIdx.x, p);
  if (p.use_tma) {
    if (tid == 0) {
      if (p.dbg & 4) { mbar_expect_tx(bar_a[0], 0); }
      else stage_issue_tma(p, &tmap, tc_cur, 0, smem_u32(a_s[0]), plane, bar_a[0]);
    }
  } else {
    stage_issue(p, tc_cur, 0, smem_u32(a_s[0]), plane, tid);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  if (tid == 0) {
    mbar_expect_tx(bar_w[0], p.b_bytes);
    bulk_g2s(smem_u32(b_s[0]), wp, p.b_bytes, bar_w[0]);
  }
  tc_prev = tc_cur;
  for (int u = 0; u < nunits; u++) {
    const int tl = u / p.nkg, kg = u - tl * p.nkg;
    const int buf = u & 1;
    if (u >= 1) {
      mbar_wait(bar_m, (u - 1) & 1);        // MMAs of unit u-1 done: the other A/B buffers and its TMEM tile are ready
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (u + 1 < nunits) {
      const int tl1 = (u + 1) / p.nkg, kg1 = (u + 1) - tl1 * p.nkg;
      tc_next.set(blockIdx.x + tl1 * gridDim.x, p);
      if (p.use_tma) {
        if (tid == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic accesses to this buffer
          if (p.dbg & 4) { mbar_expect_tx(bar_a[buf ^ 1], 0); }
          else stage_issue_tma(p, &tmap, tc_next, kg1, smem_u32(a_s[buf ^ 1]), plane, bar_a[buf ^ 1]);
        }
      } else {
        stage_issue(p, tc_next, kg1, smem_u32(a_s[buf ^ 1]), plane, tid);
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
      if (p.nkg > 1 && tid == 0) {
        mbar_expect_tx(bar_w[buf ^ 1], p.b_bytes);
        bulk_g2s(smem_u32(b_s[buf ^ 1]), wp + (size_t)kg1 * p.b_bytes, p.b_bytes, bar_w[buf ^ 1]);
      }
      if (!p.use_tma) asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      if (!p.use_tma) asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    if (p.use_tma) mbar_wait(bar_a[buf], (u >> 1) & 1);   // unit u's boxes have landed
    else __syncthreads();                                  // unit u's cp.async data visible to every thread
    if (p.has_aff) {
      if (p.use_tma) transform_flat(p, kg, a_s[buf], plane, tid);
      else transform_inplace(p, tc_cur, kg, a_s[buf], plane, tid);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to UMMA
    __syncthreads();
    // ---- MMA issue: lane 0 of warps 0..3, each for its own accumulator blocks ----
    if (warp < N_ISSUE && lane == 0) {
      if (p.nkg > 1) mbar_wait(bar_w[buf], (u >> 1) & 1);
      else if (u == 0) mbar_wait(bar_w[0], 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t kstep_a = 2u * (plane >> 4), kstep_b = 2u * (uint32_t)p.Npad;
      const uint64_t bd0 = b_desc[p.nkg > 1 ? buf : 0];
      for (int mb = warp; mb < ((p.dbg & 1) ? 0 : tc_cur.nmb); mb += N_ISSUE) {
        const uint32_t d_tmem = tmem_base + (uint32_t)(tl & 1) * tmem_half + (uint32_t)(mb * p.colstride);
        uint32_t acc = kg != 0 ? 1u : 0u;
        for (int tap = 0; tap < ntaps; tap++) {
          uint64_t ad = a_desc[buf] + (uint64_t)(uint32_t)(mb * 128 + tapoff_s[tap]);
          uint64_t bd = bd0 + (uint64_t)(uint32_t)(tap * nkc * p.Npad);
          for (int ks = 0; ks < nkc / 2; ks++) {
            umma_bf16(d_tmem, ad, bd, idesc, acc);
            acc = 1u;
            ad += kstep_a;
            bd += kstep_b;
          }
        }
      }
      umma_commit(bar_m);
    }
    // ---- epilogue of the previous tile (its last K-group was unit u-1), overlapping this unit's MMAs ----
    if (u >= 1 && kg == 0 && !(p.dbg & 2)) {
      epilogue_tile(p, tc_prev, tmem_base + (uint32_t)((tl - 1) & 1) * tmem_half, n0, myscr, stat_s, warp, lane);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    if (kg == p.nkg - 1) tc_prev = tc_cur;
    if (u + 1 < nunits) tc_cur = tc_next;
  }
  // drain: last tile
  mbar_wait(bar_m, (nunits - 1) & 1);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (!(p.dbg & 2)) epilogue_tile(p, tc_prev, tmem_base + (uint32_t)((my_tiles - 1) & 1) * tmem_half, n0, myscr, stat_s, warp, lane);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (p.has_stats) {
    for (int i = tid; i < 2 * p.Npad; i += TC_THREADS) {
      const int which = i / p.Npad, c = i - which * p.Npad;
      if (n0 + c < p.Cout) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; w++) t += stat_s[(w * 2 + which) * p.Npad + c];
        p.stat[((size_t)blockIdx.x * 2 + which) * p.Cout + n0 + c] = t;
      }
    }
  }
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols));
  }
}

// Single-buffer variant for narrow layers: TMA (or cp.async) -> transform -> MMAs -> epilogue, strictly in sequence inside
// a CTA, but small enough (<= 56 KB smem, 128 TMEM columns, 64 registers) that FOUR CTAs share an SM and overlap
// each other's phases.  Measured on B200 this beats the two-stage pipeline above when N <= 32.
__global__ void __launch_bounds__(TC_THREADS, 4) conv_tc_simple_kernel(const __grid_constant__ TcParams p,
                                                                       const __grid_constant__ CUtensorMap tmap) {
  pdl_prologue();
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 24);
  int* tapoff_s = reinterpret_cast<int*>(smem + HDR_TAPS);
  float* stat_s = reinterpret_cast<float*>(smem + HDR_STATS);
  const uint32_t hdr = (HDR_STATS + 16 * p.Npad * 4 + 127) & ~127u;
  float* scr = reinterpret_cast<float*>(smem + hdr);
  const uint32_t a_off = hdr + 8 * 32 * SCR_STRIDE * 4;
  uint8_t* a_s = smem + a_off;
  uint8_t* b_s = a_s + p.a_bytes;
  const uint32_t bar_w = smem_u32(&bars[0]), bar_m = smem_u32(&bars[2]), bar_a = smem_u32(smem + 144);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ntaps = p.kD * p.kY * p.kX;
  const int nkc = p.KG >> 3;
  const uint32_t plane = (uint32_t)p.rows_alloc * 16u;
  const int n0 = blockIdx.y * p.Npad;
  const uint8_t* wp = reinterpret_cast<const uint8_t*>(p.wp) + (size_t)blockIdx.y * p.nkg * p.b_bytes;
  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_m, N_ISSUE);
    mbar_init(bar_a, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < ntaps) {
    const int dx = tid % p.kX, dy = (tid / p.kX) % p.kY, dd = tid / (p.kX * p.kY);
    const int e = dx - p.pX;
    const int q = e >= 0 ? e / p.sX : -((-e + p.sX - 1) / p.sX);
    const int res = e - q * p.sX;
    const int set = p.nsets == 1 ? 0 : res;
    tapoff_s[tid] = set * nkc * p.rows_alloc + dd * p.Lr + dy * p.Xp + q + p.hl;
  }
  for (int i = tid; i < 16 * p.Npad; i += TC_THREADS) stat_s[i] = 0.f;   // [8 warps][2][Npad]
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)p.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.Npad >> 3) << 17) | (8u << 24);
  const uint64_t a_desc0 = make_desc(smem_u32(a_s), plane, 128u);
  const uint64_t b_desc0 = make_desc(smem_u32(b_s), (uint32_t)p.Npad * 16u, 128u);
  float* myscr = scr + warp * 32 * SCR_STRIDE;
  if (tid == 0) {                                      // weights: resident for the whole CTA (nkg == 1 in this variant)
    mbar_expect_tx(bar_w, p.b_bytes);
    bulk_g2s(smem_u32(b_s), wp, p.b_bytes, bar_w);
  }
  const int ntiles = p.NB * p.nD * p.nI;
  uint32_t n = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, n++) {
    TileCoord tc;
    tc.set(tile, p);
    if (p.use_tma) {
      if (tid == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        stage_issue_tma(p, &tmap, tc, 0, smem_u32(a_s), plane, bar_a);
      }
      mbar_wait(bar_a, n & 1);
    } else {
      stage_issue(p, tc, 0, smem_u32(a_s), plane, tid);
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
    }
    if (p.has_aff) {
      if (p.use_tma) transform_flat(p, 0, a_s, plane, tid);
      else transform_inplace(p, tc, 0, a_s, plane, tid);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (warp < N_ISSUE && lane == 0) {
      if (n == 0) mbar_wait(bar_w, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t kstep_a = 2u * (plane >> 4), kstep_b = 2u * (uint32_t)p.Npad;
      for (int mb = warp; mb < tc.nmb; mb += N_ISSUE) {
        const uint32_t d_tmem = tmem_base + (uint32_t)(mb * p.colstride);
        uint32_t acc = 0u;
        for (int tap = 0; tap < ntaps; tap++) {
          uint64_t ad = a_desc0 + (uint64_t)(uint32_t)(mb * 128 + tapoff_s[tap]);
          uint64_t bd = b_desc0 + (uint64_t)(uint32_t)(tap * nkc * p.Npad);
          for (int ks = 0; ks < nkc / 2; ks++) {
            umma_bf16(d_tmem, ad, bd, idesc, acc);
            acc = 1u;
            ad += kstep_a;
            bd += kstep_b;
          }
        }
      }
      umma_commit(bar_m);
    }
    mbar_wait(bar_m, n & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    epilogue_tile(p, tc, tmem_base, n0, myscr, stat_s, warp, lane);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();                                   // TMEM drained, A free
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (p.has_stats) {
    for (int i = tid; i < 2 * p.Npad; i += TC_THREADS) {
      const int which = i / p.Npad, c = i - which * p.Npad;
      if (n0 + c < p.Cout) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; w++) t += stat_s[(w * 2 + which) * p.Npad + c];
        p.stat[((size_t)blockIdx.x * 2 + which) * p.Cout + n0 + c] = t;
      }
    }
  }
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols));
  }
}

// Pack fp32 master weights [Cout][Cin][taps] to the bf16 smem image [nchunk][kg][tap][kc][n (Npad)][8].
// transposed: the operator applied is the dgrad conv: n <-> ci, k <-> co, taps flipped.
// Kc = reduction channels of the packed operator, Nc = its output channels, Npad = channels per N-chunk
__device__ __forceinline__ float pack_value(const float* __restrict__ w, int64_t i, int Cout, int Cin, int ntaps, int Kc, int Nc,
                                            int Npad, int KG, int mode, int sH, int pH, int kH, int tmin) {
  const uint32_t nkc = (uint32_t)KG >> 3;
  const uint32_t nkg = (uint32_t)(Kc / KG);
  uint32_t r = (uint32_t)i;                   // one image has < 2^31 elements: 32-bit index arithmetic
  const int e = (int)(r & 7u); r >>= 3;
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

__global__ void pack_weights_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cout, int Cin, int ntaps,
                                    int Kc, int Nc, int Npad, int KG, int nchunks, int mode, int sH, int pH, int kH,
                                    int tmin) {
  pdl_prologue();
  const int64_t total = (int64_t)nchunks * ntaps * Kc * Npad;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(pack_value(w, i, Cout, Cin, ntaps, Kc, Nc, Npad, KG, mode, sH, pH, kH, tmin));
}

// All images of the packed-weight arena in one launch: element -> job by binary search on the prefix sums.
__global__ void pack_all_kernel(const ffpn_pack_job* __restrict__ jobs, int njobs, long long total, bf16* __restrict__ arena) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int lo = 0, hi = njobs - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (jobs[mid].prefix <= i) lo = mid; else hi = mid - 1;
    }
    const ffpn_pack_job& j = jobs[lo];
    const long long li = i - j.prefix;
    arena[j.dst + li] = __float2bfloat16_rn(pack_value(j.w, li, j.Cout, j.Cin, j.ntaps, j.Kc, j.Nc, j.Npad, j.KG, j.mode, j.sH, j.pH, j.kH, j.tmin));
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
  const int ntaps = p.kD * p.kY * p.kX;
  const int maxinner = (p.kY - 1) * p.Xp + hr;
  p.colstride = p.Npad < 32 ? 32 : p.Npad;
  // K-group: largest of Cin(<=64)/64/32/16 whose weight image fits 72 KB
  int KG = 64;
  while (KG > 16 && (Cin % KG != 0 || (size_t)ntaps * KG * p.Npad * 2 > 72 * 1024)) KG >>= 1;
  if (Cin % KG != 0) return pl;
  if ((size_t)ntaps * KG * p.Npad * 2 > 200 * 1024) return pl;
  p.KG = KG; p.nkg = Cin / KG;
  p.b_bytes = (unsigned)((size_t)ntaps * KG * p.Npad * 2);
  const uint32_t hdr = (HDR_STATS + 16 * p.Npad * 4 + 127) & ~127u;
  const size_t fixed = hdr + 8 * 32 * SCR_STRIDE * 4 + (size_t)(p.nkg > 1 ? 2 : 1) * p.b_bytes;
  // Tile size: bounded by the TMEM columns and shared memory of the target occupancy.  Several CTAs per SM is
  // what overlaps one CTA's staging / epilogue with another's MMAs, so small-N layers aim for 4 CTAs per SM.
  if (ntaps > 27) return pl;
  // {CTAs per SM, TMEM columns per CTA}: two accumulator buffers (tile parity) share the columns
  // budgets: {CTAs per SM, TMEM columns per CTA, buffers}.  First the single-buffer 4-CTA/SM variant (narrow layers with
  // resident weights), then the two-stage pipeline with two accumulator buffers sharing the columns.
  const int budgets[3][3] = {{4, 128, 1}, {2, 256, 2}, {1, 512, 2}};
  for (int bi = 0; bi < 3; bi++) {
    const int per_sm = budgets[bi][0], nbuf = budgets[bi][2];
    if (nbuf == 1 && (p.nkg != 1 || p.colstride > 32)) continue;
    int nmb_max = budgets[bi][1] / nbuf / p.colstride;
    if (nmb_max > 8) nmb_max = 8;
    if (nmb_max < (bi == 2 ? 1 : 2)) continue;
    const size_t smem_cap = (size_t)(227 * 1024) / per_sm - 1024;
    for (; nmb_max >= (bi == 2 ? 1 : 2); nmb_max--) {
      const int max_rows = nmb_max * 128;
      int tD = 1, L = 0, Lr = 0, region = 0, tY = 0, tma_mode = -1;
      bool tma = false;
      // ---- tiles that are TMA boxes: whole lines (mode 0), slice runs (mode 1), 256-row blocks (mode 2) ----
      if (p.sX == 1) {
        if (p.kD > 1) {
          L = p.X < 128 ? p.X : 128; Lr = L;
          tD = max_rows / L; if (tD > p.oD) tD = p.oD; if (tD < 1) tD = 1;
          region = (tD + p.kD - 1) * L;
          tma = tD + p.kD - 1 <= 256; tma_mode = 1;
        } else if (p.Y == 1 && p.D == 1 && p.kY == 1 && p.kX == 1) {
          if (p.X % 256 == 0) {
            int nblk = max_rows / 256; if (nblk < 1) nblk = 1;
            if (nblk * 256 > p.X) nblk = p.X / 256;
            L = nblk * 256; Lr = L; tD = 1; region = L;
            tma = nblk <= 256 && max_rows >= 256; tma_mode = 2;
          }
        } else if (p.Xp <= 256) {
          tY = max_rows / p.Xp;
          if (tY >= 1) {
            if (tY >= p.oY) {
              tY = p.oY;
              Lr = (tY + p.kY - 1) * p.Xp;
              tD = (max_rows - tY * p.Xp) / Lr + 1; if (tD > p.oD) tD = p.oD; if (tD < 1) tD = 1;
            } else {
              Lr = (tY + p.kY - 1) * p.Xp; tD = 1;
            }
            L = tY * p.Xp; region = tD * Lr;
            tma = tY + p.kY - 1 <= 256 && tD <= 256; tma_mode = 0;
          }
        }
      }
      if (!tma) {
        tma_mode = -1; tY = 0;
        if (p.kD > 1) {
          L = p.X < 128 ? p.X : 128; Lr = L;
          tD = max_rows / L; if (tD > p.oD) tD = p.oD; if (tD < 1) tD = 1;
        } else if (p.Qout >= max_rows) {
          const int nt = (p.Qout + max_rows - 1) / max_rows;
          L = (((p.Qout + nt - 1) / nt) + 127) & ~127; if (L > max_rows) L = max_rows;
          Lr = L + maxinner; tD = 1;
        } else {
          L = p.Qout; Lr = L + maxinner;
          tD = (max_rows - L) / Lr + 1; if (tD > p.oD) tD = p.oD;
        }
        region = (tD + p.kD - 1) * Lr;
      }
      const int M_total = (tD - 1) * Lr + L;
      const int nmb = (M_total + 127) / 128;
      const int maxoff = (p.kD - 1) * Lr + maxinner;
      int rows_alloc;
      size_t a_bytes;
      if (tma) {
        rows_alloc = (region + 7) & ~7;                           // each plane is its own TMA box; 128-byte aligned
        const int over = nmb * 128 + maxoff - rows_alloc;         // MMA rows read past the last plane
        a_bytes = (size_t)(KG / 8) * rows_alloc * 16 + (size_t)(over > 0 ? over : 0) * 16;
      } else {
        rows_alloc = nmb * 128 + maxoff;
        if (rows_alloc < region) rows_alloc = region;
        a_bytes = (size_t)p.nsets * rows_alloc * KG * 2;
      }
      a_bytes = (a_bytes + 127) & ~(size_t)127;
      const size_t smem = fixed + (size_t)nbuf * a_bytes;
      if (smem > smem_cap) continue;
      pl.simple = nbuf == 1;
      p.tD = tD; p.L = L; p.Lr = Lr; p.tY = tY;
      p.use_tma = tma ? 1 : 0; p.tma_mode = tma_mode;
      p.rows_alloc = rows_alloc; p.region_rows = region;
      p.tma_bytes = (unsigned)((size_t)(KG / 8) * region * 16);
      p.a_bytes = (unsigned)a_bytes;
      int cols = nmb * p.colstride, tc = 32;
      while (tc < cols) tc <<= 1;
      if (nbuf == 2) tc <<= 1;                        // two accumulator buffers
      p.tmem_cols = tc;
      p.nD = (p.oD + tD - 1) / tD;
      p.nI = (p.Qout + L - 1) / L;
      pl.smem = smem;
      const int ntiles = p.NB * p.nD * p.nI;
      pl.grid = ntiles < per_sm * num_sms ? ntiles : per_sm * num_sms;
      pl.ok = ((uint64_t)p.nsets * (KG / 8) * rows_alloc * 16 < (1u << 18)) && tc <= 512;
      return pl;
    }
  }
  return pl;
}


namespace {
// =====================================================================================================
// wgrad on tcgen05:  dW[tap][ci][co] = sum_rows  f(x)[row + off(tap)][ci] * dy[row][co]
// Both operands are staged in the same planar layout as the forward A tile ([8-channel chunk][row][16 B]) but are
// read as MN-major operands (positions = K): A = x planes (M = channel chunks at SBO = plane stride, 8-row
// K groups at LBO = 128 B), shifted by the tap's row offset; B = dy planes.  For narrow layers the kX row
// shifts are materialised as extra planes so that one MMA (M = 128 lanes = (shift, channel)) serves kX taps.
// Accumulators stay in TMEM across ALL tiles of a CTA (no per-tile epilogue); smem is double buffered so the
// staging of tile i+1 overlaps the MMAs of tile i; the final TMEM tile is atomically added into fp32 dW.
struct WgParams {
  int NB, D, Y, X, oD, oY, oX, kD, kY, kX, pD, pY, pX;
  long long inNB, inD, inY, outNB, outD, outY;
  int Cin, Cout, Xp, Qout, tD, L, Lr, nD, nI;
  int ci_t, co_t, nci, nco, nkcx, nkcy;
  int nshift, ngroups, ntaps;
  int sX, nsets, hl;
  int goff[27];
  int Kpad, rows_x, colstride, tmem_cols;
  unsigned xbuf_bytes, ybuf_bytes;
  int relu, has_aff;
  const bf16* x;
  const bf16* dy;
  const float* sc;
  const float* sh;
  float* dw;
  long long dw_slice;    // > 0: CTA column blockIdx.x adds into its own zeroed slice dw + blockIdx.x * dw_slice (summed in a fixed order afterwards)
};

constexpr int WG_THREADS = 512;

__global__ void __launch_bounds__(WG_THREADS) conv_wgrad_tc_kernel(const __grid_constant__ WgParams p) {
  pdl_prologue();
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);          // [0], [8]: one per smem buffer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 16);
  uint8_t* xs[2] = {smem + 128, smem + 128 + p.xbuf_bytes};
  uint8_t* ys[2] = {smem + 128 + 2 * (size_t)p.xbuf_bytes, smem + 128 + 2 * (size_t)p.xbuf_bytes + p.ybuf_bytes};
  const uint32_t bar[2] = {smem_u32(&bars[0]), smem_u32(&bars[1])};
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t plane_x = (uint32_t)p.rows_x * 16u, plane_y = (uint32_t)p.Kpad * 16u;
  const int cit = blockIdx.y % p.nci, cot = blockIdx.y / p.nci;
  const int ci0 = cit * p.ci_t, co0 = cot * p.co_t;

  if (tid == 0) {
    mbar_init(bar[0], N_ISSUE);
    mbar_init(bar[1], N_ISSUE);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)p.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  // D=f32, A=B=bf16, both MN-major (bits 15,16), N = co_t, M = 128
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.co_t >> 3) << 17) | (8u << 24);

  const int kcx = tid % p.nkcx, RPX = WG_THREADS / p.nkcx;
  const int kcy = tid % p.nkcy, RPY = WG_THREADS / p.nkcy;
  float s[8], h[8];
  if (p.has_aff) {
    const int cofs = ci0 + kcx * 8;
#pragma unroll
    for (int q = 0; q < 8; q++) { s[q] = p.sc[cofs + q]; h[q] = p.sh[cofs + q]; }
  }
  const int xrows_needed = p.rows_x + p.nshift - 1;
  const int region_rows = (p.tD + p.kD - 1) * p.Lr;

  uint32_t n = 0;
  const int ntiles = p.NB * p.nD * p.nI;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, n++) {
    const int buf = n & 1;
    if (n >= 2) {
      mbar_wait(bar[buf], ((n >> 1) - 1) & 1);       // MMAs that read this buffer two tiles ago are done
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    int t = tile;
    const int it = t % p.nI; t /= p.nI;
    const int dt = t % p.nD;
    const int nb = t / p.nD;
    const int d0 = dt * p.tD, i0 = it * p.L;
    const int tD_t = min(p.tD, p.oD - d0);
    const int L_t = min(p.L, p.Qout - i0);
    const int M_t = (tD_t - 1) * p.Lr + L_t;
    // ---- stage x (with the producer's BN+ReLU, zero padding), kX shifted copies when materialised ----
    for (int set = 0; set < p.nsets; set++) {
      const int res = p.nsets == 1 ? (p.sX == 1 ? 0 : ((-p.pX) % p.sX + p.sX) % p.sX) : set;
      const bf16* xb = p.x + (long long)nb * p.inNB * p.Cin + ci0 + kcx * 8;
      uint8_t* xdst = xs[buf] + (size_t)set * p.nshift * p.nkcx * plane_x;
      int r = tid / p.nkcx;
      int j = r / p.Lr, ii = r - j * p.Lr;
      int yp = (i0 + ii) / p.Xp, xp = (i0 + ii) - yp * p.Xp;
      while (r < xrows_needed) {
        uint4 v[8];
        int rr[8];
        bool ok[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
          rr[u] = r;
          const int d = d0 + j - p.pD, yy = yp - p.pY;
          const int xx = p.sX == 1 ? xp - p.hl : p.sX * (xp - p.hl) + res;
          ok[u] = (r < region_rows) && d >= 0 && d < p.D && yy >= 0 && yy < p.Y && xx >= 0 && xx < p.X;
          v[u] = make_uint4(0u, 0u, 0u, 0u);
          if (ok[u]) v[u] = *reinterpret_cast<const uint4*>(xb + ((long long)d * p.inD + (long long)yy * p.inY + xx) * p.Cin);
          r += RPX; ii += RPX; xp += RPX;
          if (ii >= p.Lr) {
            while (ii >= p.Lr) { ii -= p.Lr; j++; }
            yp = (i0 + ii) / p.Xp; xp = (i0 + ii) - yp * p.Xp;
          } else {
            while (xp >= p.Xp) { xp -= p.Xp; yp++; }
          }
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
          if (rr[u] < xrows_needed) {
            if (p.has_aff && ok[u]) v[u] = bn_relu_bf16x8(v[u], s, h, p.relu);
            for (int sft = 0; sft < p.nshift; sft++) {
              const int rd = rr[u] - sft;
              if (rd >= 0 && rd < p.rows_x)
                *reinterpret_cast<uint4*>(xdst + (size_t)(sft * p.nkcx + kcx) * plane_x + (size_t)rd * 16) = v[u];
            }
          }
        }
      }
    }
    // ---- stage dy rows (same padded-flat row space as the forward output rows; garbage rows are zero) ----
    {
      const bf16* yb = p.dy + co0 + kcy * 8;
      for (int m = tid / p.nkcy; m < p.Kpad; m += RPY) {
        const int j = m / p.Lr, ii = m - j * p.Lr;
        const int i = i0 + ii;
        const int oy = i / p.Xp, ox = i - oy * p.Xp;
        const bool valid = (m < M_t) && (j < tD_t) && (ii < L_t) && (oy < p.oY) && (ox < p.oX);
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (valid) {
          const long long opos = (long long)nb * p.outNB + (long long)(d0 + j) * p.outD + (long long)oy * p.outY + ox;
          v = *reinterpret_cast<const uint4*>(yb + opos * p.Cout);
        }
        *reinterpret_cast<uint4*>(ys[buf] + (size_t)kcy * plane_y + (size_t)m * 16) = v;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (warp < N_ISSUE && lane == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // MN-major, no swizzle: SBO = stride between 8-channel chunks (plane), LBO = stride between 8-row K groups
      const uint64_t a0 = make_desc(smem_u32(xs[buf]), 128u, plane_x);
      const uint64_t b0 = make_desc(smem_u32(ys[buf]), 128u, plane_y);
      const int ksteps = p.Kpad >> 4;
      for (int g = warp; g < p.ngroups; g += N_ISSUE) {
        const uint32_t d_tmem = tmem_base + (uint32_t)(g * p.colstride);
        uint64_t ad = a0 + (uint64_t)(uint32_t)p.goff[g];
        uint64_t bd = b0;
        uint32_t acc = n != 0 ? 1u : 0u;
        for (int ks = 0; ks < ksteps; ks++) {
          umma_bf16(d_tmem, ad, bd, idesc, acc);
          acc = 1u;
          ad += 16; bd += 16;                          // 16 rows = 256 bytes = 16 descriptor units
        }
      }
      umma_commit(bar[buf]);
    }
  }
  // ---- drain: wait for the last MMAs on both buffers, then TMEM -> atomicAdd into dW ----
  if (n >= 1) { const uint32_t uses = (n + 1) >> 1; mbar_wait(bar[0], (uses - 1) & 1); }
  if (n >= 2) { const uint32_t uses = n >> 1; mbar_wait(bar[1], (uses - 1) & 1); }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (n >= 1) {
    const int L128 = (warp & 3) * 32 + lane;           // TMEM lane = (shift, chunk, channel-in-chunk)
    const int q = L128 >> 3, e = L128 & 7;
    const int sft = q / p.nkcx, c = q - sft * p.nkcx;
    const bool lane_ok = q < p.nshift * p.nkcx;
    const int ci = ci0 + c * 8 + e;
    const int nch = p.co_t >> 4;
    for (int w = warp >> 2; w < p.ngroups * nch; w += WG_THREADS / 128) {
      const int g = w / nch, ch = w - g * nch;
      uint32_t raw[16];
      tmem_ld16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(g * p.colstride + ch * 16), raw);
      if (lane_ok && ci < p.Cin) {
        const int tap = g * p.nshift + sft;
        if (tap < p.ntaps) {
#pragma unroll
          for (int k = 0; k < 16; k++) {
            const int co = co0 + ch * 16 + k;
            if (co < p.Cout) atomicAdd(p.dw + (size_t)blockIdx.x * p.dw_slice + ((size_t)co * p.Cin + ci) * p.ntaps + tap, __uint_as_float(raw[k]));
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols));
  }
}

struct WgPlan {
  WgParams p;
  size_t smem;
  dim3 grid;
  bool ok;
};

WgPlan make_wgrad_plan(const ffpn_conv_desc* d, int num_sms) {
  WgPlan w;
  memset(&w, 0, sizeof(w));
  Plan f = ffpn_tc_make_plan(d, false, num_sms);          // canonical geometry (also validates dtype / stride / channels)
  if (!f.ok || d->Cout % 16 != 0) return w;
  WgParams& p = w.p;
  const TcParams& c = f.p;
  p.NB = c.NB; p.D = c.D; p.Y = c.Y; p.X = c.X; p.oD = c.oD; p.oY = c.oY; p.oX = c.oX;
  p.kD = c.kD; p.kY = c.kY; p.kX = c.kX; p.pD = c.pD; p.pY = c.pY; p.pX = c.pX;
  p.inNB = c.inNB; p.inD = c.inD; p.inY = c.inY; p.outNB = c.outNB; p.outD = c.outD; p.outY = c.outY;
  p.Cin = d->Cin; p.Cout = d->Cout; p.Xp = c.Xp; p.Qout = c.Qout;
  p.sX = c.sX; p.nsets = c.nsets; p.hl = c.hl;
  p.ntaps = p.kD * p.kY * p.kX;
  if (p.ntaps > 27) return w;
  p.ci_t = p.Cin < 128 ? p.Cin : 128;
  if (p.Cin % p.ci_t != 0) return w;
  p.nci = p.Cin / p.ci_t;
  p.nkcx = p.ci_t / 8;
  // materialise the kX row shifts as extra planes when they fit the 16 chunks of an M=128 operand
  p.nshift = (p.sX == 1 && p.kX > 1 && p.nkcx * p.kX <= 16) ? p.kX : 1;
  const int ngroups = p.ntaps / p.nshift;
  p.ngroups = ngroups;
  int hr = 0;
  for (int dx = 0; dx < p.kX; dx++) {
    const int e = dx - p.pX;
    const int q = e >= 0 ? e / p.sX : -((-e + p.sX - 1) / p.sX);
    if (q + p.hl > hr) hr = q + p.hl;
  }
  const int maxinner = (p.kY - 1) * p.Xp + hr;
  // output-channel tile: accumulators of all groups must fit 512 TMEM columns
  int co_t = 256;
  while (co_t > 16 && (co_t > p.Cout || ngroups * (co_t < 32 ? 32 : co_t) > 512 || p.Cout % co_t != 0)) co_t >>= 1;
  if (p.Cout % co_t != 0 || ngroups * (co_t < 32 ? 32 : co_t) > 512) return w;
  p.co_t = co_t; p.nco = p.Cout / co_t; p.nkcy = co_t / 8;
  p.colstride = co_t < 32 ? 32 : co_t;
  int tc = 32;
  while (tc < ngroups * p.colstride) tc <<= 1;
  p.tmem_cols = tc;
  for (int Lmax = 512; Lmax >= 128; Lmax -= 128) {
    int tD, L, Lr;
    if (p.kD > 1) {
      L = p.X < 128 ? p.X : 128; Lr = L;
      tD = Lmax / L; if (tD > p.oD) tD = p.oD; if (tD < 1) tD = 1;
    } else if (p.Qout >= Lmax) {
      const int nt = (p.Qout + Lmax - 1) / Lmax;
      L = (((p.Qout + nt - 1) / nt) + 127) & ~127; if (L > Lmax) L = Lmax;
      Lr = L + maxinner; tD = 1;
    } else {
      L = p.Qout; Lr = L + maxinner;
      tD = (Lmax - L) / Lr + 1; if (tD > p.oD) tD = p.oD; if (tD < 1) tD = 1;
    }
    const int M_total = (tD - 1) * Lr + L;
    const int Kpad = (M_total + 15) & ~15;
    int maxg = 0, gset[27];
    for (int g = 0; g < ngroups; g++) {
      const int tap = g * p.nshift;                       // first tap of the group (dx = 0 when materialised)
      const int dx = tap % p.kX, dy = (tap / p.kX) % p.kY, dd = tap / (p.kX * p.kY);
      const int e = dx - p.pX;
      const int q = e >= 0 ? e / p.sX : -((-e + p.sX - 1) / p.sX);
      gset[g] = p.nsets == 1 ? 0 : e - q * p.sX;
      p.goff[g] = dd * Lr + dy * p.Xp + q + p.hl;
      if (p.goff[g] > maxg) maxg = p.goff[g];
    }
    const int rows_x = Kpad + maxg;
    for (int g = 0; g < ngroups; g++) p.goff[g] += gset[g] * p.nshift * p.nkcx * rows_x;   // residue set base (rows)
    const size_t xbuf = (size_t)p.nsets * p.nshift * p.nkcx * rows_x * 16, ybuf = (size_t)p.nkcy * Kpad * 16;
    size_t smem = 128 + 2 * xbuf + 2 * ybuf;
    // an M=128 operand always reads 16 chunk planes: keep the (ignored) extra ones inside the allocation
    const size_t reach = 128 + xbuf + (size_t)(p.nsets - 1) * p.nshift * p.nkcx * rows_x * 16 + (size_t)16 * rows_x * 16;
    if (reach > smem) smem = reach;
    if (smem > 224 * 1024 || (uint64_t)rows_x * 16 >= (1u << 18)) continue;
    p.tD = tD; p.L = L; p.Lr = Lr; p.Kpad = Kpad; p.rows_x = rows_x;
    p.xbuf_bytes = (unsigned)xbuf; p.ybuf_bytes = (unsigned)ybuf;
    p.nD = (p.oD + tD - 1) / tD;
    p.nI = (p.Qout + L - 1) / L;
    const int ntiles = p.NB * p.nD * p.nI;
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm > 512 / tc) per_sm = 512 / tc;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 2) per_sm = 2;
    int gx = per_sm * num_sms / (p.nci * p.nco);
    if (gx < 1) gx = 1;
    if (gx > ntiles) gx = ntiles;
    w.grid = dim3(gx, p.nci * p.nco);
    w.smem = smem;
    w.ok = true;
    return w;
  }
  return w;
}


// ---- wgrad, TMA-staged variant -------------------------------------------------------------------------
// Operand roles: A = dy planes (M = output channels; an M=128 operand always reads 16 chunk planes, the unused ones
// alias the x buffers that follow in smem and land in ignored TMEM lanes), B = x planes (N = kX shifts x input
// channels: the kX shifted copies are simply kX TMA loads with shifted X coordinates), K = positions.
// Both tiles are written by the TMA unit directly in the planar layout; out-of-image elements are zero (dy) or NaN
// (x, turned into 0 by the in-place BN+ReLU pass).  Rows that no box covers are zeroed once at kernel start.
struct WgTmaParams {
  int NB, D, Y, X, oD, oY, oX, kD, kY, kX, pD, pY, pX, hl;
  int Cin, Cout, Xp, tD, tY, L, Lr, nD, nI;
  int ci_t, co_t, nci, nco, nkcx, nkcy, nshift, ngroups, ntaps;
  int goff[27];
  int Kpad, rows_x, xbox_rows, ybox_rows, ncol, colstride, tmem_cols, tma_mode;
  unsigned xbuf_bytes, ybuf_bytes, tma_bytes;
  int has_aff, dbg, nstages;
  const float* sc;
  const float* sh;
  float* dw;
  long long dw_slice;    // > 0: CTA column blockIdx.x adds into its own zeroed slice dw + blockIdx.x * dw_slice (summed in a fixed order afterwards)
};

constexpr int WG_STAGES = 3;          // maximum ring depth (p.nstages = 2 or 3)
constexpr int WG_HDR = 128 + 2 * 256 * 4;  // barriers + tmem slot, then scale/shift of this CTA's input-channel tile

__device__ __forceinline__ void wgrad_tma_issue(const WgTmaParams& p, const CUtensorMap* tmx, const CUtensorMap* tmy, int tile,
                                                uint32_t xd, uint32_t yd, uint32_t plane_x, uint32_t plane_y, int ci0, int co0,
                                                uint32_t bar) {
  int t = tile;
  const int it = t % p.nI; t /= p.nI;
  const int dt = t % p.nD;
  const int nb = t / p.nD;
  const int d0 = dt * p.tD;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  mbar_expect_tx(bar, (p.dbg & 4) ? 0u : p.tma_bytes);
  if (p.dbg & 4) return;
  if (p.tma_mode == 0) {
    const int y0 = it * p.tY;
    for (int kc = 0; kc < p.nkcx; kc++)            // shift 0 only: the transform pass writes the kX shifted copies
      tma_load_4d(xd + (uint32_t)kc * plane_x, tmx, ci0 + kc * 8, -p.hl, y0 - p.pY, d0, bar);
    for (int j = 0; j < p.tD; j++)
      for (int kc = 0; kc < p.nkcy; kc++)
        tma_load_4d(yd + (uint32_t)kc * plane_y + (uint32_t)(j * p.Lr) * 16u, tmy, co0 + kc * 8, 0, y0, d0 + j, bar);
  } else if (p.tma_mode == 1) {
    const int i0 = it * p.L;
    for (int kc = 0; kc < p.nkcx; kc++) tma_load_4d(xd + (uint32_t)kc * plane_x, tmx, ci0 + kc * 8, i0, d0 - p.pD, nb, bar);
    for (int kc = 0; kc < p.nkcy; kc++) tma_load_4d(yd + (uint32_t)kc * plane_y, tmy, co0 + kc * 8, i0, d0, nb, bar);
  } else {
    const int b0 = (it * p.L) >> 8;
    for (int kc = 0; kc < p.nkcx; kc++) tma_load_4d(xd + (uint32_t)kc * plane_x, tmx, ci0 + kc * 8, 0, b0, 0, bar);
    for (int kc = 0; kc < p.nkcy; kc++) tma_load_4d(yd + (uint32_t)kc * plane_y, tmy, co0 + kc * 8, 0, b0, 0, bar);
  }
}

// Three-stage ring: TMA of tile n+1  ||  in-place BN+ReLU of tile n  ||  MMAs of tile n-1 (async, TMEM-resident accumulators).
__global__ void __launch_bounds__(WG_THREADS) conv_wgrad_tma_kernel(const __grid_constant__ WgTmaParams p,
                                                                     const __grid_constant__ CUtensorMap tmx,
                                                                     const __grid_constant__ CUtensorMap tmy) {
  pdl_prologue();
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);          // [0..2] MMA done per stage; [3..5] TMA landed per stage
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 64);
  float* sc_s = reinterpret_cast<float*>(smem + 128);
  float* sh_s = sc_s + 256;
  const int NST = p.nstages;
  uint8_t* ybase = smem + WG_HDR;
  uint8_t* xbase = smem + WG_HDR + NST * (size_t)p.ybuf_bytes;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t plane_x = (uint32_t)p.rows_x * 16u, plane_y = (uint32_t)p.Kpad * 16u;
  const int cit = blockIdx.y % p.nci, cot = blockIdx.y / p.nci;
  const int ci0 = cit * p.ci_t, co0 = cot * p.co_t;
  const int nqx = p.nshift * p.nkcx;

  if (tid == 0) {
    for (int i = 0; i < WG_STAGES; i++) { mbar_init(smem_u32(&bars[i]), N_ISSUE); mbar_init(smem_u32(&bars[WG_STAGES + i]), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (p.has_aff) {
    for (int i = tid; i < p.ci_t; i += WG_THREADS) { sc_s[i] = p.sc[ci0 + i]; sh_s[i] = p.sh[ci0 + i]; }
  }
  {  // zero all stages once: rows outside the TMA boxes must be finite (x) / zero (dy) for the K reduction
    uint4* z = reinterpret_cast<uint4*>(smem + WG_HDR);
    const int n16 = (int)((NST * ((size_t)p.ybuf_bytes + (size_t)p.xbuf_bytes)) >> 4);
    for (int i = tid; i < n16; i += WG_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)p.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  // D=f32, A=B=bf16, both MN-major, N = ncol, M = 128
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.ncol >> 3) << 17) | (8u << 24);

  const int ntiles = p.NB * p.nD * p.nI;
  const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  if (tid == 0)
    wgrad_tma_issue(p, &tmx, &tmy, blockIdx.x, smem_u32(xbase), smem_u32(ybase), plane_x, plane_y, ci0, co0, smem_u32(&bars[WG_STAGES]));
  for (int n = 0; n < my_tiles; n++) {
    const int b = n % NST;
    if (n + 1 < my_tiles) {
      const int b1 = (n + 1) % NST, u1 = (n + 1) / NST;
      if (u1 >= 1) {
        mbar_wait(smem_u32(&bars[b1]), (u1 - 1) & 1);         // MMAs that read stage b1 (tile n-2) are done
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      if (tid == 0)
        wgrad_tma_issue(p, &tmx, &tmy, blockIdx.x + (n + 1) * gridDim.x, smem_u32(xbase + (size_t)b1 * p.xbuf_bytes),
                        smem_u32(ybase + (size_t)b1 * p.ybuf_bytes), plane_x, plane_y, ci0, co0, smem_u32(&bars[WG_STAGES + b1]));
    }
    uint8_t* xs = xbase + (size_t)b * p.xbuf_bytes;
    uint8_t* ys = ybase + (size_t)b * p.ybuf_bytes;
    mbar_wait(smem_u32(&bars[WG_STAGES + b]), (n / NST) & 1);
    if ((p.has_aff || p.nshift > 1) && !(p.dbg & 8)) {
      // chunk plane by chunk plane, consecutive threads on consecutive rows (conflict-free 16-byte accesses):
      // BN+ReLU in place on the shift-0 plane and the kX-1 shifted copies written from the same registers
      for (int kc = 0; kc < p.nkcx; kc++) {
        float s[8], h[8];
        if (p.has_aff) {
          const float4 s0 = *reinterpret_cast<const float4*>(sc_s + kc * 8), s1 = *reinterpret_cast<const float4*>(sc_s + kc * 8 + 4);
          const float4 h0 = *reinterpret_cast<const float4*>(sh_s + kc * 8), h1 = *reinterpret_cast<const float4*>(sh_s + kc * 8 + 4);
          s[0] = s0.x; s[1] = s0.y; s[2] = s0.z; s[3] = s0.w; s[4] = s1.x; s[5] = s1.y; s[6] = s1.z; s[7] = s1.w;
          h[0] = h0.x; h[1] = h0.y; h[2] = h0.z; h[3] = h0.w; h[4] = h1.x; h[5] = h1.y; h[6] = h1.z; h[7] = h1.w;
        }
        uint4* base = reinterpret_cast<uint4*>(xs + (size_t)kc * plane_x);
        for (int r = tid; r < p.xbox_rows; r += WG_THREADS) {
          uint4 v = base[r];
          if (p.has_aff) { v = bn_relu_bf16x8(v, s, h, 1); base[r] = v; }
          for (int sft = 1; sft < p.nshift; sft++)
            if (r >= sft) *reinterpret_cast<uint4*>(xs + (size_t)(sft * p.nkcx + kc) * plane_x + (size_t)(r - sft) * 16) = v;
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (warp < N_ISSUE && lane == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint64_t a0 = make_desc(smem_u32(ys), 128u, plane_y);
      const uint64_t b0 = make_desc(smem_u32(xs), 128u, plane_x);
      const int ksteps = p.Kpad >> 4;
      for (int g = warp; g < ((p.dbg & 1) ? 0 : p.ngroups); g += N_ISSUE) {
        const uint32_t d_tmem = tmem_base + (uint32_t)(g * p.colstride);
        uint64_t ad = a0;
        uint64_t bd = b0 + (uint64_t)(uint32_t)p.goff[g];
        uint32_t acc = n != 0 ? 1u : 0u;
        for (int ks = 0; ks < ksteps; ks++) {
          umma_bf16(d_tmem, ad, bd, idesc, acc);
          acc = 1u;
          ad += 16; bd += 16;
        }
      }
      umma_commit(smem_u32(&bars[b]));
    }
  }
  // drain: the last use of every stage
  for (int b = 0; b < NST; b++) {
    if (my_tiles > b) {
      const int uses = (my_tiles - b + NST - 1) / NST;
      mbar_wait(smem_u32(&bars[b]), (uses - 1) & 1);
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    const int L128 = (warp & 3) * 32 + lane;           // TMEM lane = output channel within the tile
    const int co = co0 + L128;
    const int nch = p.ncol >> 4;
    for (int w = warp >> 2; w < p.ngroups * nch; w += WG_THREADS / 128) {
      const int g = w / nch, ch = w - g * nch;
      uint32_t raw[16];
      tmem_ld16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(g * p.colstride + ch * 16), raw);
      if (L128 < p.co_t && co < p.Cout) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
          const int nn = ch * 16 + k;
          const int sft = nn / p.ci_t, ci = ci0 + nn - sft * p.ci_t;
          const int tap = g * p.nshift + sft;
          if (tap < p.ntaps && ci < p.Cin) atomicAdd(p.dw + (size_t)blockIdx.x * p.dw_slice + ((size_t)co * p.Cin + ci) * p.ntaps + tap, __uint_as_float(raw[k]));
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols));
  }
}

struct WgTmaPlan {
  WgTmaParams p;
  size_t smem;
  dim3 grid;
  bool ok;
  // tensor-map geometry
  int xbox[4], ybox[4];
  long long xdims[4], ydims[4], xstr[3], ystr[3];
};

WgTmaPlan make_wgrad_tma_plan(const ffpn_conv_desc* d, int num_sms) {
  WgTmaPlan w;
  memset(&w, 0, sizeof(w));
  Plan f = ffpn_tc_make_plan(d, false, num_sms);
  if (!f.ok || d->Cout % 16 != 0 || f.p.sX != 1) return w;
  const TcParams& c = f.p;
  WgTmaParams& p = w.p;
  p.NB = c.NB; p.D = c.D; p.Y = c.Y; p.X = c.X; p.oD = c.oD; p.oY = c.oY; p.oX = c.oX;
  p.kD = c.kD; p.kY = c.kY; p.kX = c.kX; p.pD = c.pD; p.pY = c.pY; p.pX = c.pX; p.hl = c.hl;
  p.Cin = d->Cin; p.Cout = d->Cout; p.Xp = c.Xp;
  p.ntaps = p.kD * p.kY * p.kX;
  if (p.ntaps > 27) return w;
  const bool flat = (p.Y == 1 && p.D == 1 && p.kY == 1 && p.kX == 1 && p.kD == 1);
  if (p.kD > 1) p.tma_mode = 1;
  else if (flat) { if (p.X % 256 != 0) return w; p.tma_mode = 2; }
  else { if (p.Xp > 256) return w; p.tma_mode = 0; }
  p.nshift = (p.tma_mode == 0 && p.kX > 1) ? p.kX : 1;
  p.ngroups = p.ntaps / p.nshift;
  p.co_t = p.Cout < 128 ? p.Cout : 128;
  if (p.Cout % p.co_t != 0) return w;
  p.nco = p.Cout / p.co_t; p.nkcy = p.co_t / 8;
  int ci_t = 256;
  while (ci_t > 16 && (ci_t > p.Cin || p.Cin % ci_t != 0 || p.nshift * ci_t > 256 ||
                       p.ngroups * (p.nshift * ci_t < 32 ? 32 : p.nshift * ci_t) > 512)) ci_t >>= 1;
  if (p.Cin % ci_t != 0 || p.nshift * ci_t > 256 || p.ngroups * (p.nshift * ci_t < 32 ? 32 : p.nshift * ci_t) > 512) return w;
  p.ci_t = ci_t; p.nci = p.Cin / ci_t; p.nkcx = ci_t / 8;
  p.ncol = p.nshift * ci_t;
  p.colstride = p.ncol < 32 ? 32 : p.ncol;
  int tc = 32;
  while (tc < p.ngroups * p.colstride) tc <<= 1;
  p.tmem_cols = tc;
  int hr = 0;
  for (int dx = 0; dx < p.kX; dx++) if (dx - p.pX + p.hl > hr) hr = dx - p.pX + p.hl;
  for (int attempt = 0; attempt < 8; attempt++) {
    // prefer deep (3-stage) rings with large tiles; fall back to 2 stages before shrinking the tile below 256 rows
    static const int cand[8][2] = {{512, 3}, {384, 3}, {512, 2}, {256, 3}, {384, 2}, {256, 2}, {128, 3}, {128, 2}};
    const int Lmax = cand[attempt][0], nst = cand[attempt][1];
    int tD = 1, tY = 0, L, Lr, xbox_rows, ybox_rows;
    if (p.tma_mode == 1) {
      L = p.X < 128 ? p.X : 128; Lr = L;
      tD = Lmax / L; if (tD > p.oD) tD = p.oD; if (tD < 1) tD = 1;
      xbox_rows = (tD + p.kD - 1) * L; ybox_rows = tD * L;
    } else if (p.tma_mode == 2) {
      int nblk = Lmax / 256; if (nblk < 1) continue;
      if (nblk * 256 > p.X) nblk = p.X / 256;
      L = nblk * 256; Lr = L; xbox_rows = L; ybox_rows = L;
    } else {
      tY = Lmax / p.Xp; if (tY < 1) continue;
      if (tY >= p.oY) {
        tY = p.oY;
        Lr = (tY + p.kY - 1) * p.Xp;
        tD = (Lmax - tY * p.Xp) / Lr + 1; if (tD > p.oD) tD = p.oD; if (tD < 1) tD = 1;
        if (tD > 1 && (Lr % 8) != 0) tD = 1;             // per-slab dy boxes must start 128-byte aligned
        if (tD > 64) tD = 64;
      } else {
        Lr = (tY + p.kY - 1) * p.Xp; tD = 1;
      }
      L = tY * p.Xp; xbox_rows = tD * Lr; ybox_rows = L;
    }
    const int M_total = (tD - 1) * Lr + L;
    const int Kpad = (M_total + 15) & ~15;
    int maxg = 0;
    for (int g = 0; g < p.ngroups; g++) {
      const int tap = g * p.nshift;
      const int dx = tap % p.kX, dy = (tap / p.kX) % p.kY, dd = tap / (p.kX * p.kY);
      p.goff[g] = dd * Lr + dy * p.Xp + (p.nshift > 1 ? 0 : dx - p.pX + p.hl);
      if (p.goff[g] > maxg) maxg = p.goff[g];
    }
    int rows_x = Kpad + maxg; if (rows_x < xbox_rows) rows_x = xbox_rows;
    rows_x = (rows_x + 7) & ~7;
    const size_t xbuf = (size_t)p.nshift * p.nkcx * rows_x * 16, ybuf = (size_t)p.nkcy * Kpad * 16;
    size_t smem = WG_HDR + nst * (xbuf + ybuf);
    const size_t reach = WG_HDR + (nst - 1) * ybuf + (size_t)16 * Kpad * 16;   // dy operand reads 16 chunk planes from the last Y stage
    p.nstages = nst;
    if (reach > smem) smem = reach;
    if (smem > 224 * 1024 || (uint64_t)rows_x * 16 >= (1u << 18)) continue;
    p.tD = tD; p.tY = tY; p.L = L; p.Lr = Lr; p.Kpad = Kpad; p.rows_x = rows_x;
    p.xbox_rows = xbox_rows; p.ybox_rows = ybox_rows;
    p.xbuf_bytes = (unsigned)xbuf; p.ybuf_bytes = (unsigned)ybuf;
    p.tma_bytes = (unsigned)((size_t)p.nkcx * xbox_rows * 16 + (size_t)p.nkcy * (p.tma_mode == 0 ? tD * ybox_rows : ybox_rows) * 16);
    p.nD = (p.oD + tD - 1) / tD;
    if (p.tma_mode == 0) p.nI = (p.oY + tY - 1) / tY;
    else p.nI = (p.X + L - 1) / L;
    // tensor maps
    const long long cbx = (long long)p.Cin * 2, cby = (long long)p.Cout * 2;
    w.xdims[0] = p.Cin; w.ydims[0] = p.Cout; w.xbox[0] = 8; w.ybox[0] = 8;
    if (p.tma_mode == 0) {
      w.xdims[1] = p.X; w.xdims[2] = p.Y; w.xdims[3] = p.D;
      w.xstr[0] = cbx; w.xstr[1] = (p.Y == 1 ? (long long)p.X : c.inY) * cbx; w.xstr[2] = c.inD * cbx;
      w.xbox[1] = p.Xp; w.xbox[2] = tY + p.kY - 1; w.xbox[3] = tD;
      w.ydims[1] = p.oX; w.ydims[2] = p.oY; w.ydims[3] = p.oD;
      w.ystr[0] = cby; w.ystr[1] = (p.oY == 1 ? (long long)p.oX : c.outY) * cby; w.ystr[2] = c.outD * cby;
      w.ybox[1] = p.Xp; w.ybox[2] = tY; w.ybox[3] = 1;
    } else if (p.tma_mode == 1) {
      w.xdims[1] = p.X; w.xdims[2] = p.D; w.xdims[3] = p.NB;
      w.xstr[0] = cbx; w.xstr[1] = c.inD * cbx; w.xstr[2] = (p.NB > 1 ? c.inNB : c.inD * p.D) * cbx;
      w.xbox[1] = L; w.xbox[2] = tD + p.kD - 1; w.xbox[3] = 1;
      w.ydims[1] = p.oX; w.ydims[2] = p.oD; w.ydims[3] = p.NB;
      w.ystr[0] = cby; w.ystr[1] = c.outD * cby; w.ystr[2] = (p.NB > 1 ? c.outNB : c.outD * p.oD) * cby;
      w.ybox[1] = L; w.ybox[2] = tD; w.ybox[3] = 1;
    } else {
      w.xdims[1] = 256; w.xdims[2] = p.X / 256; w.xdims[3] = 1;
      w.xstr[0] = cbx; w.xstr[1] = 256 * cbx; w.xstr[2] = (long long)p.X * cbx;
      w.xbox[1] = 256; w.xbox[2] = L / 256; w.xbox[3] = 1;
      w.ydims[1] = 256; w.ydims[2] = p.X / 256; w.ydims[3] = 1;
      w.ystr[0] = cby; w.ystr[1] = 256 * cby; w.ystr[2] = (long long)p.X * cby;
      w.ybox[1] = 256; w.ybox[2] = L / 256; w.ybox[3] = 1;
    }
    for (int i = 1; i < 4; i++) if (w.xbox[i] > 256 || w.ybox[i] > 256) return w;
    const int ntiles = p.NB * p.nD * p.nI;
    int gx = num_sms / (p.nci * p.nco);
    if (gx < 1) gx = 1;
    if (gx > ntiles) gx = ntiles;
    w.grid = dim3(gx, p.nci * p.nco);
    w.smem = smem;
    w.ok = true;
    return w;
  }
  return w;
}

bool encode_map4(CUtensorMap* m, const void* base, const long long* dims, const long long* str, const int* box, bool nan_fill) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return false;
  cuuint64_t gd[4], gs[3];
  cuuint32_t bx[4], es[4] = {1, 1, 1, 1};
  for (int i = 0; i < 4; i++) { gd[i] = (cuuint64_t)dims[i]; bx[i] = (cuuint32_t)box[i]; }
  for (int i = 0; i < 3; i++) gs[i] = (cuuint64_t)str[i];
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             nan_fill ? CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA : CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

namespace {
// dw[i] += sum over the K-split slices, in slice order (bitwise reproducible, unlike atomics on dw itself)
__global__ void __launch_bounds__(128) wgrad_slice_reduce_kernel(const float* __restrict__ part, int nslices, long long n, float* __restrict__ dw) {
  pdl_prologue();
  // one block per output: thread t adds slices t, t + 128, ... then a fixed-shape tree (same bits every run)
  __shared__ float red[128];
  for (long long i = blockIdx.x; i < n; i += gridDim.x) {
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
    __syncthreads();
  }
}
constexpr size_t FFPN_WG_SLICE_MAX_BYTES = (size_t)64 << 20;
// K-split slices for the first-generation wgrad kernels: returns the slice length (elements) when the workspace can hold them
long long wgrad_slices(const ffpn_conv_desc* d, unsigned gx, size_t ws_bytes) {
  const long long n = (long long)d->Cout * d->Cin * d->kS * d->kW * d->kH;
  const size_t need = (size_t)gx * (size_t)n * 4;
  return (need <= FFPN_WG_SLICE_MAX_BYTES && need <= ws_bytes) ? n : 0;
}
int wgrad_slices_finish(ffpn_ctx* ctx, const float* part, unsigned gx, long long n, float* dw, cudaStream_t st) {
  const int blocks = (int)(n < 16384 ? n : 16384);
  ffpn_launch(wgrad_slice_reduce_kernel, blocks, 128, 0, st, part, (int)gx, n, dw);
  FFPN_CHECK_LAUNCH(ctx, "wgrad_slice_reduce");
  return 0;
}
}  // namespace

bool ffpn_wgrad_ws_supported(const ffpn_conv_desc* d);
bool ffpn_tc_wgrad_supported(const ffpn_conv_desc* d) { return ffpn_wgrad_ws_supported(d) || make_wgrad_plan(d, 148).ok; }

size_t ffpn_wgrad_ws_workspace_bytes(const ffpn_conv_desc* d);
int ffpn_conv_wgrad_ws(ffpn_ctx*, const ffpn_conv_desc*, const void*, const float*, const float*, int, const void*, float*, void*, size_t,
                       cudaStream_t);

int ffpn_conv_wgrad_tc(ffpn_ctx* ctx, const ffpn_conv_desc* d, const void* x, const float* in_scale, const float* in_shift,
                       int in_relu, const void* dy, float* dw, void* ws, size_t ws_bytes, cudaStream_t st) {
  {
    const int r = ffpn_conv_wgrad_ws(ctx, d, x, in_scale, in_shift, in_relu, dy, dw, ws, ws_bytes, st);
    if (r >= 0) { ctx->routes[FFPN_ROUTE_WS]++; return r; }   // handled (or failed) by the warp-specialised kernel
  }
  ffpn_log_route("conv_wgrad -> first-generation tcgen05 kernel", d);
  ctx->routes[FFPN_ROUTE_GEN1]++;
  if (!(in_scale != nullptr && !in_relu)) {
    WgTmaPlan t = make_wgrad_tma_plan(d, ctx->num_sms);
    CUtensorMap tmx, tmy;
    if (t.ok && encode_map4(&tmx, x, t.xdims, t.xstr, t.xbox, in_scale != nullptr) &&
        encode_map4(&tmy, dy, t.ydims, t.ystr, t.ybox, false)) {
      WgTmaParams& q = t.p;
      q.sc = in_scale; q.sh = in_shift; q.dw = dw; q.has_aff = in_scale != nullptr;
      q.dbg = ffpn_debug_env("FFPN_TC_DEBUG");
      if (!(ctx->attr_mask & FFPN_ATTR_WGRAD_TMA)) {
        cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) FFPN_FAIL(ctx, "conv_wgrad_tma: cannot raise dynamic smem: %s", cudaGetErrorString(e));
        ctx->attr_mask |= FFPN_ATTR_WGRAD_TMA;
      }
      q.dw_slice = ws ? wgrad_slices(d, t.grid.x, ws_bytes) : 0;
      if (q.dw_slice) {
        q.dw = (float*)ws;
        if (cudaMemsetAsync(ws, 0, (size_t)t.grid.x * q.dw_slice * 4, st) != cudaSuccess) FFPN_FAIL(ctx, "conv_wgrad_tma: memset failed");
      }
      ffpn_launch(conv_wgrad_tma_kernel, t.grid, WG_THREADS, t.smem, st, q, tmx, tmy);
      FFPN_CHECK_LAUNCH(ctx, "conv_wgrad_tma");
      return q.dw_slice ? wgrad_slices_finish(ctx, (const float*)ws, t.grid.x, q.dw_slice, dw, st) : 0;
    }
  }
  WgPlan w = make_wgrad_plan(d, ctx->num_sms);
  if (!w.ok) FFPN_FAIL(ctx, "conv_wgrad_tc: geometry not supported");
  WgParams& p = w.p;
  p.x = (const bf16*)x; p.dy = (const bf16*)dy; p.sc = in_scale; p.sh = in_shift; p.dw = dw;
  p.relu = in_relu; p.has_aff = in_scale != nullptr;
  if (!(ctx->attr_mask & FFPN_ATTR_WGRAD_TC)) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) FFPN_FAIL(ctx, "conv_wgrad_tc: cannot raise dynamic smem: %s", cudaGetErrorString(e));
    ctx->attr_mask |= FFPN_ATTR_WGRAD_TC;
  }
  p.dw_slice = ws ? wgrad_slices(d, w.grid.x, ws_bytes) : 0;
  if (p.dw_slice) {
    p.dw = (float*)ws;
    if (cudaMemsetAsync(ws, 0, (size_t)w.grid.x * p.dw_slice * 4, st) != cudaSuccess) FFPN_FAIL(ctx, "conv_wgrad_tc: memset failed");
  }
  ffpn_launch(conv_wgrad_tc_kernel, w.grid, WG_THREADS, w.smem, st, p);
  FFPN_CHECK_LAUNCH(ctx, "conv_wgrad_tc");
  return p.dw_slice ? wgrad_slices_finish(ctx, (const float*)ws, w.grid.x, p.dw_slice, dw, st) : 0;
}

namespace {
}  // namespace

bool ffpn_tc_fwd_supported(const ffpn_conv_desc* d) { return ffpn_tc_make_plan(d, false, 148).ok; }
bool ffpn_tc_dgrad_supported(const ffpn_conv_desc* d) { return ffpn_tc_make_plan(d, true, 148).ok; }

size_t ffpn_wgrad_ws_workspace_bytes(const ffpn_conv_desc* d);
size_t ffpn_tc_workspace_bytes(const ffpn_conv_desc* d) {
  const size_t taps = (size_t)d->kS * d->kW * d->kH;
  const size_t cin = (d->Cin + 63) & ~63, cout = (d->Cout + 63) & ~63;
  const size_t pack = taps * cin * cout * 2 + 65536;                  // packed weights | wgrad partial tiles
  size_t wg = ffpn_wgrad_ws_workspace_bytes(d);
  if (wg == 0 && d->dtype == FFPN_BF16) {                              // first-generation wgrad: one dW slice per CTA column
    const size_t slices = (size_t)296 * d->Cout * d->Cin * taps * 4;
    if (slices <= FFPN_WG_SLICE_MAX_BYTES) wg = slices;
  }
  return pack > wg ? pack : wg;
}

namespace {
// 4-D map over a channels-last activation; one box = one 8-channel plane of a tile.
bool encode_act_map(CUtensorMap* m, const TcParams& p, const void* x, bool nan_fill) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return false;
  const cuuint64_t cb = (cuuint64_t)p.Cin * 2;
  cuuint64_t dims[4], strides[3];
  cuuint32_t box[4], es[4] = {1, 1, 1, 1};
  dims[0] = (cuuint64_t)p.Cin; box[0] = 8;
  if (p.tma_mode == 0) {
    dims[1] = p.X; dims[2] = p.Y; dims[3] = p.D;
    strides[0] = cb; strides[1] = (cuuint64_t)p.inY * cb; strides[2] = (cuuint64_t)p.inD * cb;
    box[1] = p.Xp; box[2] = p.tY + p.kY - 1; box[3] = p.tD;
    if (p.Y == 1) strides[1] = (cuuint64_t)p.X * cb;              // unused dimension: any valid stride
  } else if (p.tma_mode == 1) {
    dims[1] = p.X; dims[2] = p.D; dims[3] = p.NB;
    strides[0] = cb; strides[1] = (cuuint64_t)p.inD * cb; strides[2] = (cuuint64_t)(p.NB > 1 ? p.inNB : (long long)p.inD * p.D) * cb;
    box[1] = p.L; box[2] = p.tD + p.kD - 1; box[3] = 1;
  } else {
    dims[1] = 256; dims[2] = p.X / 256; dims[3] = 1;
    strides[0] = cb; strides[1] = 256 * cb; strides[2] = (cuuint64_t)p.X * cb;
    box[1] = 256; box[2] = p.L / 256; box[3] = 1;
  }
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         nan_fill ? CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA : CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}
}  // namespace

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
  const int g = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
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

int ffpn_conv_fwd_ws(ffpn_ctx*, const ffpn_conv_desc*, bool transposed, const void*, const float*, const float*, int, const float*,
                     const void* addend, void*, float*, int*, void*, size_t, cudaStream_t);

int ffpn_conv_fwd_tc(ffpn_ctx* ctx, const ffpn_conv_desc* d, bool transposed, const void* x, const float* in_scale,
                     const float* in_shift, int in_relu, const float* w, const void* addend, void* y, float* stat_partial,
                     int* stat_rows, void* ws, size_t ws_bytes, cudaStream_t st) {
  {
    const int r = ffpn_conv_fwd_ws(ctx, d, transposed, x, in_scale, in_shift, in_relu, w, addend, y, stat_partial, stat_rows, ws, ws_bytes, st);
    if (r >= 0) { ctx->routes[FFPN_ROUTE_WS]++; return r; }   // handled (or failed) by the warp-specialised kernel
  }
  ffpn_log_route(transposed ? "conv_dgrad -> first-generation tcgen05 kernel" : "conv_fwd -> first-generation tcgen05 kernel", d);
  ctx->routes[FFPN_ROUTE_GEN1]++;
  Plan pl = ffpn_tc_make_plan(d, transposed, ctx->num_sms);
  if (!pl.ok) FFPN_FAIL(ctx, "conv_tc: geometry not supported");
  TcParams& p = pl.p;
  const size_t need = (size_t)pl.nchunks * p.nkg * p.b_bytes;
  if (ws == nullptr || ws_bytes < need) FFPN_FAIL(ctx, "conv_tc: workspace too small (%zu < %zu)", ws_bytes, need);
  if (stat_partial != nullptr && pl.grid > FFPN_STAT_ROWS) FFPN_FAIL(ctx, "conv_tc: %d tiles exceed the statistics buffer", pl.grid);
  bool packed_now = false;
  const void* wimg = ffpn_tc_pack_weights(ctx, w, ws, d, p, pl.nchunks, p.KG, st, &packed_now);
  if (packed_now) FFPN_CHECK_LAUNCH(ctx, "pack_weights");
  p.x = (const bf16*)x; p.sc = in_scale; p.sh = in_shift; p.wp = (const bf16*)wimg;
  p.addend = (const bf16*)addend; p.y = (bf16*)y; p.stat = stat_partial;
  p.relu = in_relu; p.has_aff = in_scale != nullptr; p.has_stats = stat_partial != nullptr; p.has_add = addend != nullptr;
  if (!(ctx->attr_mask & FFPN_ATTR_TC)) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_simple_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) FFPN_FAIL(ctx, "conv_tc: cannot raise dynamic smem: %s", cudaGetErrorString(e));
    ctx->attr_mask |= FFPN_ATTR_TC;
  }
  p.dbg = ffpn_debug_env("FFPN_TC_DEBUG");
  const dim3 grid(pl.grid, pl.nchunks);
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (p.use_tma) {
    // BN+ReLU prologue: halo filled with NaN so that relu(scale*NaN+shift) = 0; without a prologue: zero fill
    if ((p.has_aff && !p.relu) || !encode_act_map(&tmap, p, x, p.has_aff != 0)) p.use_tma = 0;
  }
  if (pl.simple) ffpn_launch(conv_tc_simple_kernel, grid, TC_THREADS, pl.smem, st, p, tmap);
  else ffpn_launch(conv_tc_kernel, grid, TC_THREADS, pl.smem, st, p, tmap);
  FFPN_CHECK_LAUNCH(ctx, transposed ? "conv_dgrad_tc" : "conv_fwd_tc");
  if (stat_rows) *stat_rows = pl.grid;
  return 0;
}

