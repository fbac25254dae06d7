// Warp-specialised tcgen05 weight-gradient kernel:  dW[co][ci][tap] = sum_q dy[q][co] * f(x)[q + off(tap)][ci].
//
// Both operands are the channels-last tiles exactly as TMA delivers them, [position][channels] with a 32/64/128-byte
// row in the matching SWIZZLE layout, read as MN-major UMMA operands (K = positions).  In that layout an operand is a
// set of "slabs" (one swizzle row of channels each) a fixed byte distance (LBO) apart -- and nothing says the slabs may
// not overlap.  With LBO = one row the slabs of the dy operand are the SAME tile shifted by 0,1,2,.. positions (the kX
// taps); with LBO = one padded line the slabs of the x operand are the tile shifted by whole lines (the kY taps, or the
// kD slice taps).  For the narrow layers one MMA per 16 positions (M = 4 shifts x 16 co, N = 3 lines x 16 ci) therefore
// produces all nine taps, with no materialised copies; wide layers use the slabs for the 64-channel groups instead and
// one accumulator per tap.  (tools/umma_probe.cu verifies overlapping slabs and unaligned starts on the hardware.)
// Accumulators stay in TMEM for the whole persistent CTA; every CTA then stores its partial tile to the workspace and
// a second small kernel sums the partials in a fixed order into the state_dict layout: no atomics, bitwise reproducible.
#include <math.h>

#include "ws_common.cuh"

namespace {

constexpr int WG2_THREADS = 512;
constexpr int WG2_NT = 160;             // transform threads: warps 3, 12..15
constexpr int WG2_NEPI = 8;             // epilogue warps 4..11
constexpr int WG2_HDR = 1024;
constexpr int WG2_MAX_STAGES = 4;
constexpr int WG2_MAX_ACC = 32;

struct WgWsParams {
  int NB, D, Y, X, oD, oY, oX, kD, kY, kX, pD, pY, pX, hl;
  int Cin, Cout, ntaps, Xp, tD, tY, L, nD, nI, tma_mode;
  int co_t, ci_t, n_co, n_ci, Cy, Cx, pitch_y, pitch_x, nys, nxs;
  int M, N, colsN, sA, sB, kA, kB, Ls, passesA, passesB, nacc_total, acc_per_cta, npg;
  int region_x, Kpad, xsub_bytes, ysub_bytes, stage_bytes, nstages, tmem_cols;
  unsigned tx_bytes, lboA, lboB;
  unsigned accA[WG2_MAX_ACC], accB[WG2_MAX_ACC];     // per global accumulator: start offsets (descriptor units) of A and B
  int pdl_early;                      // wait for the predecessor grid only after the prologue (FFPN_PDL_EARLY)
  int xseg;                           // lines wider than one TMA box: X cut into NB segments of xseg outputs (0 = off); dy then comes through a 5-D map
  int has_aff, gx, dbg, pair_cin;     // pair_cin: real Cin when the plan runs on the pair view of a depth-strided conv (0 = off)     // dbg (FFPN_TC_DEBUG, timing experiments): 1 no MMA, 4 no TMA, 8 no transform body
  const float* sc;
  const float* sh;
  float* part;                                        // [grid.y][grid.x][acc_per_cta][128][colsN]
};

struct WgTile {
  int it, dt, nb, sit, sdt, snb;
  __device__ __forceinline__ void init(const WgWsParams& p) {
    int t = blockIdx.x;
    it = t % p.nI; t /= p.nI; dt = t % p.nD; nb = t / p.nD;
    t = gridDim.x;
    sit = t % p.nI; t /= p.nI; sdt = t % p.nD; snb = t / p.nD;
  }
  __device__ __forceinline__ bool valid(const WgWsParams& p) const { return nb < p.NB; }
  __device__ __forceinline__ void next(const WgWsParams& p) {
    it += sit;
    if (it >= p.nI) { it -= p.nI; dt++; }
    dt += sdt;
    if (dt >= p.nD) { dt -= p.nD; nb++; }
    nb += snb;
  }
};

__global__ void __launch_bounds__(WG2_THREADS, 1) conv_wgrad_ws_kernel(const __grid_constant__ WgWsParams p,
                                                                       const __grid_constant__ CUtensorMap tmx,
                                                                       const __grid_constant__ CUtensorMap tmy) {
  pdl_trigger();
  if (!p.pdl_early) pdl_wait();
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  const uint32_t bar0 = smem_u32(bars);
  auto FULL = [&](int s) { return bar0 + 8u * (uint32_t)s; };
  auto READY = [&](int s) { return bar0 + 8u * (uint32_t)(4 + s); };
  auto EMPTY = [&](int s) { return bar0 + 8u * (uint32_t)(8 + s); };
  const uint32_t DONE = bar0 + 8u * 12u;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 128);
  uint8_t* stage0 = smem + WG2_HDR;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  // blockIdx.y = (pass group, co tile, ci tile)
  const int cit = blockIdx.y % p.n_ci, cot = (blockIdx.y / p.n_ci) % p.n_co, pg = blockIdx.y / (p.n_ci * p.n_co);
  const int ci0 = cit * p.ci_t, co0 = cot * p.co_t;
  const int acc0 = pg * p.acc_per_cta;
  const int nacc = min(p.acc_per_cta, p.nacc_total - acc0);

  if (tid == 0) {
    for (int s = 0; s < WG2_MAX_STAGES; s++) {
      mbar_init(FULL(s), 1);
      mbar_init(READY(s), WG2_NT / 32);
      mbar_init(EMPTY(s), 1);
    }
    mbar_init(DONE, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {  // rows no TMA box covers (leading / trailing dy rows, x rows past the box) must be zero for the K reduction
    uint4* z = reinterpret_cast<uint4*>(stage0);
    const int n16 = (p.nstages * p.stage_bytes) >> 4;
    for (int i = tid; i < n16; i += WG2_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)p.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (p.pdl_early) pdl_wait();                    // barrier init, smem zeroing and TMEM allocation overlapped the predecessor
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t ybytes = (uint32_t)(p.nys * p.ysub_bytes);

  if (warp == 0) {
    // ================= TMA producer =================
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      bool wrapped = false;
      WgTile tc;
      for (tc.init(p); tc.valid(p); tc.next(p)) {
        if (wrapped) mbar_wait(EMPTY(s), ph ^ 1u);
        const uint32_t dst = smem_u32(stage0) + (uint32_t)s * (uint32_t)p.stage_bytes;
        mbar_expect_tx(FULL(s), (p.dbg & 4) ? 0u : p.tx_bytes);
        int x1, x2, x3, y1, y2, y3;
        if (p.tma_mode == 0) { x1 = -p.hl + tc.nb * p.xseg; x2 = tc.it * p.tY - p.pY; x3 = tc.dt; y1 = 0; y2 = tc.it * p.tY; y3 = tc.dt; }
        else if (p.tma_mode == 1) { x1 = tc.it * p.L; x2 = tc.dt * p.tD - p.pD; x3 = tc.nb; y1 = x1; y2 = tc.dt * p.tD; y3 = tc.nb; }
        else { x1 = tc.it * p.L; x2 = 0; x3 = 0; y1 = x1; y2 = 0; y3 = 0; }
        const int nbox = p.tma_mode == 2 ? (p.L >> 8) : 1;        // flat: one box of 256 positions at a time
        if (!(p.dbg & 4))
        for (int j = 0; j < p.nys; j++)        // dy sub-tiles land 8 rows in: the rows before stay zero (shifted views)
        {
          const uint32_t ydst = dst + (uint32_t)(j * p.ysub_bytes) + 8u * (uint32_t)p.pitch_y;
          if (p.xseg) tma_load_5d(ydst, &tmy, co0 + j * p.Cy, 0, tc.nb, y2, y3, FULL(s));
          else for (int b = 0; b < nbox; b++) tma_load_4d(ydst + (uint32_t)(b * 256 * p.pitch_y), &tmy, co0 + j * p.Cy, y1 + b * 256, y2, y3, FULL(s));
        }
        if (!(p.dbg & 4))
        for (int j = 0; j < p.nxs; j++)
          for (int b = 0; b < nbox; b++)
            tma_load_4d(dst + ybytes + (uint32_t)(j * p.xsub_bytes) + (uint32_t)(b * 256 * p.pitch_x), &tmx, ci0 + j * p.Cx, x1 + b * 256, x2, x3, FULL(s));
        if (++s == p.nstages) { s = 0; ph ^= 1u; wrapped = true; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (uniform datapath, one elected lane) =================
    const bool leader = elect_one();
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.N >> 3) << 17) |
                           ((uint32_t)(p.M >> 4) << 24);
    const uint32_t ltA = p.pitch_y == 128 ? 2u : p.pitch_y == 64 ? 4u : 6u, ltB = p.pitch_x == 128 ? 2u : p.pitch_x == 64 ? 4u : 6u;
    const uint32_t a_hi = ((8u * (uint32_t)p.pitch_y) >> 4) | (1u << 14) | (ltA << 29);
    const uint32_t b_hi = ((8u * (uint32_t)p.pitch_x) >> 4) | (1u << 14) | (ltB << 29);
    const uint32_t a_lbo = ((p.lboA >> 4) & 0x3FFFu) << 16, b_lbo = ((p.lboB >> 4) & 0x3FFFu) << 16;
    const int ksteps = (p.dbg & 1) ? 0 : (p.Kpad >> 4);
    // slot table of the lean issue loop: slot i = (K-step j = i / nacc, accumulator a = i % nacc)
    const int J = nacc <= 4 ? 8 / nacc : 1, nslot = nacc <= 4 ? J * nacc : 0;
    uint32_t sA_[8], sB_[8], sD[8];
    int sJ[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int j = nacc <= 4 ? i / nacc : 0, a = nacc <= 4 ? i - j * nacc : 0;
      const bool on = i < nslot;
      sJ[i] = j;
      sA_[i] = on ? p.accA[acc0 + a] + (uint32_t)(j * p.pitch_y) : 0u;
      sB_[i] = on ? p.accB[acc0 + a] + (uint32_t)(j * p.pitch_x) : 0u;
      sD[i] = tmem_base + (uint32_t)(a * p.colsN);
    }
    int s = 0, n = 0;
    uint32_t ph = 0;
    WgTile tc;
    for (tc.init(p); tc.valid(p); tc.next(p), n++) {
      mbar_wait(p.has_aff ? READY(s) : FULL(s), ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t st = smem_u32(stage0) + (uint32_t)s * (uint32_t)p.stage_bytes;
      const uint32_t a_base = ((st & 0x3FFFFu) >> 4) | a_lbo, b_base = (((st + ybytes) & 0x3FFFFu) >> 4) | b_lbo;
      uint32_t a_ks = a_base, b_ks = b_base;
      int ks = 0;
      if (n == 0 && ksteps > 0) {                            // very first K-step of the CTA overwrites the accumulators
        for (int a = 0; a < nacc; a++)
          if (leader)
            umma_bf16(tmem_base + (uint32_t)(a * p.colsN), desc64(a_ks + p.accA[acc0 + a], a_hi), desc64(b_ks + p.accB[acc0 + a], b_hi), idesc, 0u);
        a_ks += (uint32_t)p.pitch_y; b_ks += (uint32_t)p.pitch_x; ks = 1;
      }
      if (nacc <= 4) {
        // lean path: eight (K-step, accumulator) slots per iteration whose descriptors are INDEPENDENT adds off one base
        // (the uniform datapath has a long latency: dependent chains between MMAs would pace the tensor pipe)
        for (; ks < ksteps; ks += J) {
#pragma unroll
          for (int i = 0; i < 8; i++)
            if (i < nslot && ks + sJ[i] < ksteps && leader)
              umma_bf16(sD[i], desc64(a_ks + sA_[i], a_hi), desc64(b_ks + sB_[i], b_hi), idesc, 1u);
          a_ks += (uint32_t)(J * p.pitch_y); b_ks += (uint32_t)(J * p.pitch_x);
        }
      } else {
        for (; ks < ksteps; ks++) {
          for (int a = 0; a < nacc; a++)
            if (leader)
              umma_bf16(tmem_base + (uint32_t)(a * p.colsN), desc64(a_ks + p.accA[acc0 + a], a_hi), desc64(b_ks + p.accB[acc0 + a], b_hi), idesc, 1u);
          a_ks += (uint32_t)p.pitch_y; b_ks += (uint32_t)p.pitch_x;
        }
      }
      if (leader) umma_commit(EMPTY(s));
      __syncwarp();
      if (++s == p.nstages) { s = 0; ph ^= 1u; }
    }
    if (leader) umma_commit(DONE);
    __syncwarp();
  } else if (warp >= 4 && warp < 4 + WG2_NEPI) {
    // ================= epilogue: TMEM partial tile -> workspace (plain coalesced stores) =================
    const int quad = warp & 3, half = (warp - 4) >> 2;
    mbar_wait(DONE, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int nch = p.colsN >> 4;
    float* dst = p.part + ((size_t)(blockIdx.y * gridDim.x + blockIdx.x) * p.acc_per_cta) * (size_t)(128 * p.colsN) +
                 (size_t)(quad * 32 + lane) * p.colsN;
    for (int w = half; w < nacc * nch; w += 2) {
      const int a = w / nch, ch = w - a * nch;
      uint32_t raw[16];
      tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(a * p.colsN + ch * 16), raw);
      float4* o = reinterpret_cast<float4*>(dst + (size_t)a * (128 * p.colsN) + ch * 16);
#pragma unroll
      for (int q = 0; q < 4; q++)
        o[q] = make_float4(__uint_as_float(raw[4 * q]), __uint_as_float(raw[4 * q + 1]), __uint_as_float(raw[4 * q + 2]),
                           __uint_as_float(raw[4 * q + 3]));
    }
  } else if (warp != 2) {
    // ================= in-place BatchNorm scale/shift + ReLU of the landed x tile (NaN-filled halo -> 0) =================
    if (p.has_aff) {
      const int tix = warp < 4 ? tid - 96 : tid - 384 + 32;        // 0..159
      const int cpu = p.nxs * (p.Cx >> 3);
      const int rstep = WG2_NT / cpu;
      const bool active = tix < rstep * cpu;
      const int c = tix % cpu, r0 = tix / cpu;
      const int xs = c / (p.Cx >> 3), cc = c % (p.Cx >> 3);
      const uint32_t cmask = (uint32_t)(p.Cx >> 3) - 1u;
      float s[8], h[8];
      if (active) {
        int cofs = ci0 + xs * p.Cx + cc * 8;
        if (p.pair_cin) cofs %= p.pair_cin;
        const float4 s0 = *reinterpret_cast<const float4*>(p.sc + cofs), s1 = *reinterpret_cast<const float4*>(p.sc + cofs + 4);
        const float4 h0 = *reinterpret_cast<const float4*>(p.sh + cofs), h1 = *reinterpret_cast<const float4*>(p.sh + cofs + 4);
        s[0] = s0.x; s[1] = s0.y; s[2] = s0.z; s[3] = s0.w; s[4] = s1.x; s[5] = s1.y; s[6] = s1.z; s[7] = s1.w;
        h[0] = h0.x; h[1] = h0.y; h[2] = h0.z; h[3] = h0.w; h[4] = h1.x; h[5] = h1.y; h[6] = h1.z; h[7] = h1.w;
      }
      int st = 0;
      uint32_t ph = 0;
      WgTile tc;
      for (tc.init(p); tc.valid(p); tc.next(p)) {
        mbar_wait(FULL(st), ph);
        if (active && !(p.dbg & 8)) {
          uint8_t* base = stage0 + (size_t)st * p.stage_bytes + ybytes + (size_t)xs * p.xsub_bytes;
          const uint32_t abase = smem_u32(base);
          int rr = r0;
          for (; rr + 3 * rstep < p.region_x; rr += 4 * rstep) {
            uint4* q[4];
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
              const uint32_t off = (uint32_t)(rr + u * rstep) * (uint32_t)p.pitch_x;
              q[u] = reinterpret_cast<uint4*>(base + off + ((((uint32_t)cc ^ ((abase + off) >> 7)) & cmask) << 4));
              v[u] = *q[u];
            }
#pragma unroll
            for (int u = 0; u < 4; u++) *q[u] = bn_relu_bf16x8(v[u], s, h, 1);
          }
          for (; rr < p.region_x; rr += rstep) {
            const uint32_t off = (uint32_t)rr * (uint32_t)p.pitch_x;
            uint4* q = reinterpret_cast<uint4*>(base + off + ((((uint32_t)cc ^ ((abase + off) >> 7)) & cmask) << 4));
            *q = bn_relu_bf16x8(*q, s, h, 1);
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(READY(st));           // one arrival per warp
        if (++st == p.nstages) { st = 0; ph ^= 1u; }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols));
  }
}

// Sum the per-CTA partial tiles (fixed order) and add them to dW in the state_dict layout [Cout][Cin][taps].
template <int LANES>
__global__ void wgrad_reduce_kernel(const WgWsParams p, float* __restrict__ dw) {
  pdl_prologue();
  // LANES = 4 (small weight tensors, latency-bound): four lanes per output, each summing every fourth partial tile (eight loads in flight), combined with two shuffles in
  // a fixed order: the loop is L2-latency bound, so the shorter dependent chains matter more than the coalescing
  const int Cr = p.pair_cin ? p.pair_cin : p.Cin, ntr = p.pair_cin ? 3 : p.ntaps;     // real (state_dict) input channels / taps
  const int64_t total = (int64_t)p.Cout * Cr * ntr;
  // LANES = 1 (large weight tensors, bandwidth-bound): one lane per output, consecutive lanes read consecutive columns
  constexpr int SH = LANES == 4 ? 2 : 0;
  const int sub = threadIdx.x & (LANES - 1);
  const int64_t nthr = ((int64_t)gridDim.x * blockDim.x) >> SH;
  const int64_t rounds = (total + nthr - 1) / nthr;
  int64_t idx = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> SH;
  for (int64_t r = 0; r < rounds; r++, idx += nthr) {                 // uniform trip count: the shuffles stay converged
    const bool live = idx < total;
    float acc = 0.f;
    int cir = 0, co = 0, tapr = 0;
    if (live) {
      cir = (int)(idx % Cr);
      co = (int)((idx / Cr) % p.Cout);
      tapr = (int)(idx / ((int64_t)Cr * p.Cout));
      int ci = cir, tap = tapr;
      if (p.pair_cin) {                     // dx -> (half h, pair tap t'): 0 -> (1,0), 1 -> (0,1), 2 -> (1,1)
        const int hh = tapr == 1 ? 0 : 1;
        tap = tapr == 0 ? 0 : 1;
        ci = hh * Cr + cir;
      }
      const int ta = tap % p.kA, tb = tap / p.kA;
      const int cot = co / p.co_t, cco = co - cot * p.co_t, cit = ci / p.ci_t, cci = ci - cit * p.ci_t;
      int passA, lanei;
      if (p.co_t <= 64) {
        passA = ta / p.sA;
        const int m = p.sA - 1 - (ta - passA * p.sA);
        lanei = p.M == 64 ? m * 32 + cco : m * p.Cy + cco;
      } else {
        passA = ta; lanei = cco;
      }
      const int passB = p.sB > 1 ? 0 : tb;
      const int col = p.sB > 1 ? tb * p.ci_t + cci : cci;
      const int g = passB * p.passesA + passA;
      const int pg = g / p.acc_per_cta, a = g - pg * p.acc_per_cta;
      const int y = (pg * p.n_co + cot) * p.n_ci + cit;
      const float* src = p.part + (((size_t)y * p.gx) * p.acc_per_cta + a) * (size_t)(128 * p.colsN) + (size_t)lanei * p.colsN + col;
      const size_t stride = (size_t)p.acc_per_cta * (128 * p.colsN);
      int k = sub;
      for (; k + 7 * LANES < p.gx; k += 8 * LANES) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = src[(size_t)(k + LANES * u) * stride];
        acc += ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
      }
      for (; k < p.gx; k += LANES) acc += src[(size_t)k * stride];
    }
    if (LANES == 4) {
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    }
    if (live && sub == 0) dw[((size_t)co * Cr + cir) * ntr + tapr] += acc;   // accumulate contract of ffpn_conv_wgrad
  }
}

struct WgWsPlan {
  WgWsParams p;
  Plan base;
  size_t smem, ws_bytes;
  dim3 grid;
  bool ok;
  int xbox[4], ybox[4];
};

bool chan_ok(int c) { return c == 16 || c == 32 || c == 64 || (c > 64 && c % 64 == 0); }
// input channels are tiled (blockIdx.y), so any multiple of 16 works there: 48 = 3 x 16, 96 = 3 x 32 (the two-input decoders)
bool cin_ok(int c) { return c >= 16 && c % 16 == 0; }

WgWsPlan make_wgrad_ws_plan(const ffpn_conv_desc* d, int num_sms) {
  WgWsPlan w;
  memset(&w, 0, sizeof(w));
  w.base = ffpn_tc_make_plan(d, false, num_sms);
  if (!w.base.ok) return w;
  const TcParams& c = w.base.p;
  if (c.sX != 1 || c.nsets != 1) return w;
  if (!cin_ok(d->Cin) || !chan_ok(d->Cout)) return w;
  WgWsParams& p = w.p;
  p.NB = c.NB; p.D = c.D; p.Y = c.Y; p.X = c.X; p.oD = c.oD; p.oY = c.oY; p.oX = c.oX;
  p.kD = c.kD; p.kY = c.kY; p.kX = c.kX; p.pD = c.pD; p.pY = c.pY; p.pX = c.pX; p.hl = c.hl;
  p.Cin = d->Cin; p.Cout = d->Cout; p.Xp = c.Xp;
  p.ntaps = p.kD * p.kY * p.kX;
  if (p.ntaps > 27) return w;
  const bool flat = (p.Y == 1 && p.D == 1 && p.kY == 1 && p.kX == 1 && p.kD == 1);
  if (p.kD > 1) { if (p.kY != 1 || p.kX != 1) return w; p.tma_mode = 1; p.kA = 1; p.kB = p.kD; }
  else if (flat) { p.tma_mode = 2; p.kA = 1; p.kB = 1; }
  else {
    p.tma_mode = 0; p.kA = p.kX; p.kB = p.kY;
    if (p.Xp > 256) {
      // a padded line does not fit one TMA box: cut X into equal segments walked like batch entries.  The x tile of a segment
      // carries real neighbour columns as halo; the dy tile must be ZERO there, which a 5-D map (C, tX, nX, Y, D) gives for
      // free: box columns >= tX are out of bounds of the segment axis and zero-filled.
      const int halo = p.Xp - p.oX;
      int nX = (p.oX + (256 - halo) - 1) / (256 - halo);
      while (nX < 64 && p.oX % nX != 0) nX++;                 // equal segments only (the 5-D dy map has one segment extent)
      if (p.NB != 1 || p.oX % nX != 0 || p.oX / nX < 8) return w;
      p.xseg = p.oX / nX; p.NB = nX; p.Xp = p.xseg + halo;
    }
  }
  // channel tiles and slabs
  p.co_t = p.Cout < 128 ? p.Cout : 128; p.n_co = p.Cout / p.co_t;
  p.Cy = p.co_t < 64 ? p.co_t : 64; p.nys = p.co_t / p.Cy; p.pitch_y = p.Cy * 2;
  p.M = p.co_t == 16 ? 64 : 128;
  p.sA = p.co_t <= 64 ? (p.M / p.co_t < p.kA ? p.M / p.co_t : p.kA) : 1;
  p.passesA = (p.kA + p.sA - 1) / p.sA;
  // input-channel tile: the whole Cin when it is one swizzle row (16 / 32 / 64) or a multiple of 64 up to 256, else 128, 64, 32, 16
  for (int ci_cand = p.Cin < 256 ? p.Cin : 256; ci_cand >= 16; ci_cand = (ci_cand > 128 ? 128 : ci_cand > 64 ? 64 : ci_cand > 32 ? 32 : ci_cand > 16 ? 16 : 0)) {
  if (p.Cin % ci_cand != 0) continue;
  if (ci_cand > 64 ? (ci_cand % 64 != 0) : !(ci_cand == 16 || ci_cand == 32 || ci_cand == 64)) continue;
  p.ci_t = ci_cand;
  p.n_ci = p.Cin / p.ci_t;
  p.Cx = p.ci_t < 64 ? p.ci_t : 64; p.nxs = p.ci_t / p.Cx; p.pitch_x = p.Cx * 2;
  p.sB = (p.ci_t <= 64 && p.kB > 1 && p.ci_t * p.kB <= 256) ? p.kB : 1;
  p.passesB = p.kB / p.sB;
  p.N = p.sB * p.ci_t; p.colsN = p.N;
  p.nacc_total = p.passesA * p.passesB;
  if (p.nacc_total > WG2_MAX_ACC) continue;
  {
    const int cap = 512 / p.colsN;
    p.npg = (p.nacc_total + cap - 1) / cap;
    p.acc_per_cta = (p.nacc_total + p.npg - 1) / p.npg;
    p.npg = (p.nacc_total + p.acc_per_cta - 1) / p.acc_per_cta;
  }
  int tc = 32;
  while (tc < p.acc_per_cta * p.colsN) tc <<= 1;
  p.tmem_cols = tc;
  const int gy = p.npg * p.n_co * p.n_ci;
  const int P = p.sA - 1;
  const size_t budget = 227 * 1024 - WG2_HDR - 1024;
  static const int cand[7] = {1024, 768, 512, 384, 256, 192, 128};
  for (int want = 3; want >= 2; want--) {
    for (int ci = 0; ci < 7; ci++) {
      const int Lmax = cand[ci];
      int tD = 1, tY = 0, L, Ls, region_x;
      if (p.tma_mode == 1) {
        const int Lp = p.X < 128 ? p.X : 128;
        tD = Lmax / Lp; if (tD > p.oD) tD = p.oD; if (tD < 1) tD = 1;
        if (tD + p.kD - 1 > 256) continue;
        L = tD * Lp; Ls = Lp; region_x = (tD + p.kD - 1) * Lp;
        p.L = Lp;
      } else if (p.tma_mode == 2) {
        int nblk = Lmax / 256; if (nblk < 1) continue;
        if (nblk * 256 > p.X) nblk = (p.X + 255) / 256;         // ragged tail: rows past the end are zero-filled by the TMA unit
        L = nblk * 256; Ls = 0; region_x = L;
        p.L = L;
      } else {
        tY = Lmax / p.Xp; if (tY < 1) continue;
        if (tY > p.oY) tY = p.oY;
        if (tY + p.kY - 1 > 256) continue;
        L = tY * p.Xp; Ls = p.Xp; region_x = (tY + p.kY - 1) * p.Xp;
        p.L = L;
      }
      const int Kpad = (L + P + 15) & ~15;
      const int maxoff = (p.kB - 1) * Ls + (p.passesA - 1) * p.sA;
      int rows_x = Kpad + maxoff; if (rows_x < region_x) rows_x = region_x;
      const int rows_y = Kpad + 16;
      const size_t xsub = ((size_t)rows_x * p.pitch_x + 1023) & ~(size_t)1023;
      const size_t ysub = ((size_t)rows_y * p.pitch_y + 1023) & ~(size_t)1023;
      const size_t stage = xsub * p.nxs + ysub * p.nys;
      int nst = (int)(budget / stage);
      if (nst > WG2_MAX_STAGES) nst = WG2_MAX_STAGES;
      if (nst < want) continue;
      if (xsub >= (1u << 18) || ysub >= (1u << 18)) continue;
      p.tD = tD; p.tY = tY; p.Ls = Ls; p.region_x = region_x; p.Kpad = Kpad;
      p.xsub_bytes = (int)xsub; p.ysub_bytes = (int)ysub; p.stage_bytes = (int)stage; p.nstages = nst;
      p.tx_bytes = (unsigned)((size_t)p.nxs * region_x * p.pitch_x + (size_t)p.nys * L * p.pitch_y);
      p.lboA = p.co_t <= 64 ? (unsigned)p.pitch_y : (unsigned)ysub;
      p.lboB = p.sB > 1 ? (unsigned)(Ls * p.pitch_x) : (unsigned)xsub;
      for (int g = 0; g < p.nacc_total; g++) {
        const int passA = g % p.passesA, passB = g / p.passesA;
        p.accA[g] = (unsigned)((8 - P) * p.pitch_y) >> 4;
        p.accB[g] = (unsigned)((passA * p.sA + (p.sB > 1 ? 0 : passB * Ls)) * p.pitch_x) >> 4;
      }
      if (p.tma_mode == 1) { p.nD = (p.oD + tD - 1) / tD; p.nI = (p.X + p.L - 1) / p.L; }
      else if (p.tma_mode == 2) { p.nD = 1; p.nI = (p.X + L - 1) / L; }
      else { p.nD = p.oD; p.nI = (p.oY + tY - 1) / tY; }
      const int ntiles = p.NB * p.nD * p.nI;
      int gx = num_sms / gy; if (gx < 1) gx = 1;
      if (gx > ntiles) gx = ntiles;
      p.gx = gx;
      w.grid = dim3(gx, gy);
      w.smem = WG2_HDR + (size_t)nst * stage;
      w.ws_bytes = (size_t)gy * gx * p.acc_per_cta * 128 * p.colsN * sizeof(float);
      // tensor-map boxes
      w.xbox[0] = p.Cx; w.ybox[0] = p.Cy;
      if (p.tma_mode == 0) { w.xbox[1] = p.Xp; w.xbox[2] = tY + p.kY - 1; w.xbox[3] = 1; w.ybox[1] = p.Xp; w.ybox[2] = tY; w.ybox[3] = 1; }
      else if (p.tma_mode == 1) { w.xbox[1] = p.L; w.xbox[2] = tD + p.kD - 1; w.xbox[3] = 1; w.ybox[1] = p.L; w.ybox[2] = tD; w.ybox[3] = 1; }
      else { w.xbox[1] = 256; w.xbox[2] = 1; w.xbox[3] = 1; w.ybox[1] = 256; w.ybox[2] = 1; w.ybox[3] = 1; }
      w.ok = true;
      return w;
    }
  }
  }
  return w;
}

bool encode_wg_map(CUtensorMap* m, const WgWsPlan& w, const void* base, bool is_x, bool nan_fill, int in_mult = 1) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return false;
  const WgWsParams& p = w.p;
  const TcParams& c = w.base.p;
  const int C = is_x ? p.Cin : p.Cout;
  if (in_mult != 1 && (p.tma_mode != 2 || !is_x)) return false;          // strided 1x1x1 shortcut: input rows in_mult * Cin apart
  const cuuint64_t cb = (cuuint64_t)C * 2 * (cuuint64_t)in_mult;
  const long long sY = is_x ? c.inY : c.outY, sD = is_x ? c.inD : c.outD, sNB = is_x ? c.inNB : c.outNB;
  const int eX = is_x ? p.X : p.oX, eY = is_x ? p.Y : p.oY, eD = is_x ? p.D : p.oD;
  cuuint64_t dims[5], strides[4];
  cuuint32_t box[5], es[5] = {1, 1, 1, 1, 1};
  dims[0] = (cuuint64_t)C;
  if (p.tma_mode == 0 && p.xseg && !is_x) {
    // dy of a segmented line: (C, tX, nX, Y, D); the box is one segment wide plus the (out-of-bounds, zero) halo columns
    dims[1] = p.xseg; dims[2] = eX / p.xseg; dims[3] = eY; dims[4] = eD;
    strides[0] = cb; strides[1] = (cuuint64_t)p.xseg * cb; strides[2] = (cuuint64_t)(eY == 1 ? eX : sY) * cb; strides[3] = (cuuint64_t)sD * cb;
    box[0] = (cuuint32_t)w.ybox[0]; box[1] = (cuuint32_t)w.ybox[1]; box[2] = 1; box[3] = (cuuint32_t)w.ybox[2]; box[4] = 1;
    const int pitch = p.pitch_y;
    const CUtensorMapSwizzle sw = pitch == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : pitch == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  }
  if (p.tma_mode == 0) {
    dims[1] = eX; dims[2] = eY; dims[3] = eD;
    strides[0] = cb; strides[1] = (cuuint64_t)(eY == 1 ? eX : sY) * cb; strides[2] = (cuuint64_t)sD * cb;
  } else if (p.tma_mode == 1) {
    dims[1] = eX; dims[2] = eD; dims[3] = p.NB;
    strides[0] = cb; strides[1] = (cuuint64_t)sD * cb; strides[2] = (cuuint64_t)(p.NB > 1 ? sNB : sD * eD) * cb;
  } else {
    dims[1] = (cuuint64_t)eX; dims[2] = 1; dims[3] = 1;      // flat: positions are one dimension, loaded 256 at a time
    strides[0] = cb; strides[1] = (cuuint64_t)eX * cb; strides[2] = (cuuint64_t)eX * cb;
  }
  for (int i = 0; i < 4; i++) box[i] = (cuuint32_t)(is_x ? w.xbox[i] : w.ybox[i]);
  const int pitch = is_x ? p.pitch_x : p.pitch_y;
  const CUtensorMapSwizzle sw = pitch == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : pitch == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             nan_fill ? CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA : CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool wgws_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("FFPN_WS"); v = (e && atoi(e) == 0) ? 0 : 1; }
  return v != 0;
}

}  // namespace

bool ffpn_wgrad_ws_supported(const ffpn_conv_desc* d) {
  ffpn_conv_desc dp;
  int mult = 1;
  const bool view = ffpn_make_strided111_desc(d, &dp, &mult) || ffpn_make_pair_desc(d, &dp);
  return d->dtype == FFPN_BF16 && wgws_enabled() && make_wgrad_ws_plan(view ? &dp : d, 148).ok;
}

size_t ffpn_wgrad_ws_workspace_bytes(const ffpn_conv_desc* d) {
  if (d->dtype != FFPN_BF16) return 0;
  ffpn_conv_desc dp;
  int mult = 1;
  const bool view = ffpn_make_strided111_desc(d, &dp, &mult) || ffpn_make_pair_desc(d, &dp);
  WgWsPlan pl = make_wgrad_ws_plan(view ? &dp : d, 148);
  return pl.ok ? pl.ws_bytes : 0;
}

// Returns 0 = launched, 1 = error, -1 = geometry not handled here (caller falls back to conv_tc.cu).
int ffpn_conv_wgrad_ws(ffpn_ctx* ctx, const ffpn_conv_desc* d, const void* x, const float* in_scale, const float* in_shift,
                       int in_relu, const void* dy, float* dw, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!wgws_enabled()) return -1;
  if (in_scale != nullptr && !in_relu) return -1;
  ffpn_conv_desc dp;
  int in_mult = 1;
  const bool strided111 = ffpn_make_strided111_desc(d, &dp, &in_mult);
  const bool pair = !strided111 && ffpn_make_pair_desc(d, &dp);
  WgWsPlan pl = make_wgrad_ws_plan((pair || strided111) ? &dp : d, ctx->num_sms);
  if (!pl.ok || ws == nullptr || ws_bytes < pl.ws_bytes) return -1;
  CUtensorMap tmx, tmy;
  if (!encode_wg_map(&tmx, pl, x, true, in_scale != nullptr, in_mult) || !encode_wg_map(&tmy, pl, dy, false, false)) return -1;
  WgWsParams& p = pl.p;
  p.sc = in_scale; p.sh = in_shift; p.has_aff = in_scale != nullptr; p.part = (float*)ws;
  p.pair_cin = pair ? d->Cin : 0;
  { const char* e = getenv("FFPN_PDL_EARLY"); p.pdl_early = (e && atoi(e) == 0) ? 0 : 1; }
  p.dbg = ffpn_debug_env("FFPN_TC_DEBUG");
  if (!(ctx->attr_mask & FFPN_ATTR_WGRAD_WS)) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) FFPN_FAIL(ctx, "conv_wgrad_ws: cannot raise dynamic smem: %s", cudaGetErrorString(e));
    ctx->attr_mask |= FFPN_ATTR_WGRAD_WS;
  }
  {
    static int verbose = -1;
    if (verbose < 0) { const char* e = getenv("FFPN_WS_VERBOSE"); verbose = e ? atoi(e) : 0; }
    if (verbose)
      fprintf(stderr, "wgrad_ws: Cin %d Cout %d taps %dx%dx%d mode %d co_t %d ci_t %d M %d N %d sA %d sB %d passes %dx%d acc/cta %d npg %d "
              "tY %d tD %d L %d Kpad %d stages %d stage_bytes %d smem %zu tmem %d grid (%u,%u) ws %zu\n", p.Cin, p.Cout, p.kD, p.kY, p.kX,
              p.tma_mode, p.co_t, p.ci_t, p.M, p.N, p.sA, p.sB, p.passesA, p.passesB, p.acc_per_cta, p.npg, p.tY, p.tD, p.L, p.Kpad,
              p.nstages, p.stage_bytes, pl.smem, p.tmem_cols, pl.grid.x, pl.grid.y, pl.ws_bytes);
  }
  ffpn_launch(conv_wgrad_ws_kernel, pl.grid, WG2_THREADS, pl.smem, st, p, tmx, tmy);
  FFPN_CHECK_LAUNCH(ctx, "conv_wgrad_ws");
  const int64_t total = pair ? (int64_t)d->Cout * d->Cin * 3 : (int64_t)p.Cout * p.Cin * p.ntaps;
  if (total <= 65536) {
    const int blocks = (int)((total * 4 + 255) / 256 < 1184 ? (total * 4 + 255) / 256 : 1184);
    ffpn_launch(wgrad_reduce_kernel<4>, blocks, 256, 0, st, p, dw);
  } else {
    const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
    ffpn_launch(wgrad_reduce_kernel<1>, blocks, 256, 0, st, p, dw);
  }
  FFPN_CHECK_LAUNCH(ctx, "wgrad_reduce");
  return 0;
}
