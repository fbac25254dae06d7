// Warp-specialised tcgen05 implicit-GEMM convolution (forward and, through transposed packed weights, dgrad) for the
// stride-1 shapes of the path.  Same formulation as conv_tc.cu (padded-flat positions, every tap = the same smem tile
// seen through a shifted UMMA descriptor) but:
//   * the activation tile is kept ROW-MAJOR in shared memory, [position][channels] with a 32/64/128-byte row pitch in
//     the matching SWIZZLE_32B/64B/128B K-major layout, so ONE TMA box (inner extent = the whole channel run) stages a
//     tile.  The hardware swizzle is a function of the absolute smem address, so a descriptor whose start address is
//     shifted by any number of rows still reads the right data (tools/umma_probe.cu, base_offset 0);
//   * roles are split over warps and decoupled by mbarriers: warp 0 issues TMA into a ring of stages, four warps (one
//     per SM sub-partition) apply the producer's BatchNorm scale/shift + ReLU in place (halo = NaN fill -> 0), warps 1-2
//     issue the MMAs (alternating tiles; tap-outer, accumulator-block-inner so consecutive MMAs hit different TMEM
//     tiles), eight warps drain the TMEM tile buffers, store channels-last bf16 rows with one 256-bit store per lane
//     (+ residual addend) and accumulate the BatchNorm partial sums in registers (N <= 32) or by warp-shuffle transposes;
//   * lines wider than one TMA box are cut into X segments walked like batch entries (xseg); narrow 3-tap layers can run
//     on the pair view of input and output (pair2, off by default).
#include "ws_common.cuh"

namespace {

constexpr int WS_THREADS = 512;
constexpr int WS_MAX_STAGES = 6;
constexpr int WS_NT = 128;            // transform threads: warps 12..15, one per SM sub-partition
// Register split between the warpgroups (setmaxnreg; 128 threads each, 512 x 128 registers in all): the TMA / MMA issuers and the
// transform warps need few, the epilogue warps keep two batches of accumulator rows, the addend / BatchNorm-backward operands
// and the per-channel sums in registers.
constexpr int WS_REGS_ISSUE = 64, WS_REGS_XFORM = 80, WS_REGS_EPI = 184;
static_assert(WS_REGS_ISSUE + WS_REGS_XFORM + 2 * WS_REGS_EPI == 4 * 128, "the four warpgroups share 512 x 128 registers");
constexpr int WS_NEPI = 8;            // epilogue warps 4..11
constexpr int WS_STAT_BYTES = 8 * 2 * 256 * 4;   // one statistics slot per epilogue warp (fixed-order sum: deterministic)
constexpr int WS_TAB_BYTES = 1024 * 8;           // row table of the epilogue: 8 accumulator blocks x 128 rows x (offset, packed j|y|x)
constexpr int WS_HDR = 1024 + WS_STAT_BYTES + WS_TAB_BYTES;   // barriers first

struct WsParams {
  int NB, D, Y, X, oD, oY, oX, kD, kY, kX, pD, pY, pX, hl;
  long long outNB, outD, outY;
  int Cin, Cout, Npad;
  int Xp, tD, tY, L, Lr, nD, nI, Qout, tma_mode;
  int Kc, nkg, pitch, kgu, upt;       // channels per sub-tile (= swizzle row), K-groups, row pitch, K-groups / unit, units / tile
  int region_rows, sub_bytes, a_unit_bytes, stage_bytes, nstages, niss, nst_ring0, nst_ring1, nbuf;   // issuer warps, stages of ring 0 / 1, TMEM tile buffers
  int w_resident;
  unsigned b_unit_bytes, b_total_bytes, tx_bytes;
  int colstride, tmem_cols;
  int fshift;                         // flat (1x1x1) mode: log2 of the TMA box row unit (256 or 128 positions)
  int aff_mod;                        // BatchNorm vectors are indexed modulo this (pair view: two positions share them); 0 = off
  int relu, has_aff, has_stats, has_add, dbg;   // dbg (FFPN_TC_DEBUG, timing experiments): 1 no MMA, 2 no epilogue body, 4 no TMA, 8 no transform body, 16 no output stores, 32 no statistics
  const float* sc;
  const float* sh;
  const bf16* wp;
  const bf16* addend;                 // MODE 1: tensor added to the output; MODE 2: raw output y of the BatchNorm whose backward sums are taken
  const float* bsc;                   // MODE 2: scale / shift of that BatchNorm (the ReLU mask is scale * y + shift > 0)
  const float* bsh;
  unsigned zero;                      // always 0, but only the host knows: ties the issue of an epilogue load to the arrival of the previous one (see consume)
  int bn_mod;                         // MODE 2 on the pair view of dx: the BatchNorm vectors are indexed modulo the real channel count (0 = off)
  int fold;                           // statistics columns c and c + fold are the same channel (pair views): folded when the CTA writes its row (0 = off)
  bf16* y;
  float* stat;
  unsigned tapdesc[27];               // per tap: row offset of the A view in descriptor units ((rows * pitch) >> 4)
  unsigned tap_dx, tap_dy, tap_dd;    // its increments per dx / dy / dd step (the issue loop only loads tapdesc[0])
  int pdl_early;                      // wait for the predecessor grid only after the prologue (FFPN_PDL_EARLY)
  int pair2;                          // stride-1 conv on the pair view of input and output: real Cout (statistics are folded to it), 0 = off
  int xseg, oXtot;                    // lines wider than one TMA box: X is cut into segments of xseg outputs that take the place of the batch axis (0 = off)
  int fin_on;                         // BatchNorm finalize by the last CTA to finish (ffpn_conv_fwd_bn)
  ffpn_bn_fin fin;
  unsigned* fin_counter;
  long long* trace;                   // FFPN_WS_TRACE: per-tile clock64 stamps of CTA 0 (debug)
};

#define WS_TRACE(slot, tl) do { if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && (tl) < 64) p.trace[(tl) * 16 + (slot)] = clock64(); } while (0)

// Tile walk of a persistent CTA without divisions: tile = blockIdx.x + n * gridDim.x decomposed as (nb, dt, it).
struct WsTile {
  int it, dt, nb, sit, sdt, snb;
  int d0, i0, tD_t, L_t, M_t, nmb;
  __device__ __forceinline__ void init(const WsParams& p) {
    int t = blockIdx.x;
    it = t % p.nI; t /= p.nI; dt = t % p.nD; nb = t / p.nD;
    t = gridDim.x;
    sit = t % p.nI; t /= p.nI; sdt = t % p.nD; snb = t / p.nD;
    derive(p);
  }
  __device__ __forceinline__ bool valid(const WsParams& p) const { return nb < p.NB; }
  __device__ __forceinline__ void next(const WsParams& p) {
    it += sit;
    if (it >= p.nI) { it -= p.nI; dt++; }
    dt += sdt;
    if (dt >= p.nD) { dt -= p.nD; nb++; }
    nb += snb;
    derive(p);
  }
  __device__ __forceinline__ void derive(const WsParams& p) {
    d0 = dt * p.tD; i0 = it * p.L;
    tD_t = min(p.tD, p.oD - d0);
    L_t = min(p.L, p.Qout - i0);
    M_t = (tD_t - 1) * p.Lr + L_t;
    nmb = (M_t + 127) >> 7;
  }
};

// Position of one unit in the stage rings.  With two MMA issuers the stages are split into two rings (tile parity), so
// that each issuer only ever waits on barriers whose previous phase it consumed itself.
struct WsRing {
  int i0, i1;
  uint32_t ph0, ph1;
  __device__ __forceinline__ void init() { i0 = i1 = 0; ph0 = ph1 = 0; }
  __device__ __forceinline__ int stage(int r, const WsParams& p) const { return r + p.niss * (r ? i1 : i0); }
  __device__ __forceinline__ uint32_t phase(int r) const { return r ? ph1 : ph0; }
  __device__ __forceinline__ bool advance(int r, const WsParams& p) {     // true when the ring wrapped
    if (r) { if (++i1 == p.nst_ring1) { i1 = 0; ph1 ^= 1u; return true; } }
    else { if (++i0 == p.nst_ring0) { i0 = 0; ph0 ^= 1u; return true; } }
    return false;
  }
};

template <int N> static __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> static __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// NREG: 16-column chunks whose statistics are accumulated in registers (Npad == 16 * NREG); 0 = shuffle per chunk.
// MODE: 0 plain; 1 the epilogue adds a tensor (residual-branch gradient of a dgrad); 2 dgrad fused with pass 1 of the backward of
// the BatchNorm + ReLU that produced the conv's input: the epilogue loads that BatchNorm's raw input y at the output position,
// masks the gradient with the ReLU (scale * y + shift > 0) and accumulates the per-channel sums of G and G * y.
template <int NREG, int MODE, bool SPLITC>
__global__ void __launch_bounds__(WS_THREADS, 1) conv_ws_kernel(const __grid_constant__ WsParams p,
                                                                const __grid_constant__ CUtensorMap tmap) {
  pdl_trigger();
  if (!p.pdl_early) pdl_wait();
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  const uint32_t bar0 = smem_u32(bars);
  // full[s] @ 0, ready[s] @ 6, empty[s] @ 12, tfull[b] @ 18, tempty[b] @ 22, wbar @ 26
  auto FULL = [&](int s) { return bar0 + 8u * (uint32_t)s; };
  auto READY = [&](int s) { return bar0 + 8u * (uint32_t)(6 + s); };
  auto EMPTY = [&](int s) { return bar0 + 8u * (uint32_t)(12 + s); };
  auto TFULL = [&](int b) { return bar0 + 8u * (uint32_t)(18 + b); };
  auto TEMPTY = [&](int b) { return bar0 + 8u * (uint32_t)(22 + b); };
  const uint32_t WBAR = bar0 + 8u * 26u;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 240);
  float* stat_s = reinterpret_cast<float*>(smem + 1024);
  uint2* row_tab = reinterpret_cast<uint2*>(smem + 1024 + WS_STAT_BYTES);
  uint8_t* w_s = smem + WS_HDR;
  uint8_t* stage0 = w_s + (p.w_resident ? ((p.b_total_bytes + 1023u) & ~1023u) : 0u);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // provably warp-uniform: role branches become uniform branches
  const int ntaps = p.kD * p.kY * p.kX;
  const int n0 = blockIdx.y * p.Npad;
  const uint8_t* wp = reinterpret_cast<const uint8_t*>(p.wp) + (size_t)blockIdx.y * p.b_total_bytes;

  if (tid == 0) {
    for (int s = 0; s < WS_MAX_STAGES; s++) {
      mbar_init(FULL(s), 1);
      mbar_init(READY(s), (uint32_t)(WS_NT >> 5));
      mbar_init(EMPTY(s), 1);
    }
    for (int b = 0; b < 4; b++) { mbar_init(TFULL(b), 1); mbar_init(TEMPTY(b), SPLITC ? WS_NEPI : WS_NEPI / 2); }
    mbar_init(WBAR, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < WS_NEPI * 2 * p.Npad; i += WS_THREADS) stat_s[i] = 0.f;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)p.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (p.pdl_early) pdl_wait();                    // nothing above touched global memory: the prologue overlapped the predecessor
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t buf_cols = (uint32_t)(p.tmem_cols / p.nbuf);

  // Every warpgroup sets its register budget at the top of its own branch (the budget must dominate the role code).  Nobody
  // returns to the launch allocation afterwards: an early finisher asking for its registers back (the spare warp 3, the
  // transform warps of a conv without prologue) would race the epilogue warps for the pool and starve them.  The common tail
  // is therefore compiled for the smallest budget.
  if (warp < 4) {
  reg_dec<WS_REGS_ISSUE>();
  if (warp == 0) {
    // ================= TMA producer (one elected lane) =================
    if (elect_one()) {
      if (p.w_resident) {
        mbar_expect_tx(WBAR, p.b_total_bytes);
        bulk_g2s(smem_u32(w_s), wp, p.b_total_bytes, WBAR);
      }
      WsRing ring;
      ring.init();
      bool wrapped0 = false, wrapped1 = false;            // first pass over a ring: its stages are free
      WsTile tc;
      int ptl = 0;
      for (tc.init(p); tc.valid(p); tc.next(p), ptl++) {
        const int r = p.niss == 2 ? (ptl & 1) : 0;
        int c1, c2, c3;
        if (p.tma_mode == 0) { c1 = -p.hl + tc.nb * p.xseg; c2 = tc.it * p.tY - p.pY; c3 = tc.d0; }
        else if (p.tma_mode == 1) { c1 = tc.i0; c2 = tc.d0 - p.pD; c3 = tc.nb; }
        else { c1 = tc.i0; c2 = 0; c3 = 0; }
        for (int uk = 0; uk < p.upt; uk++) {
          const int s = ring.stage(r, p);
          if (r ? wrapped1 : wrapped0) mbar_wait(EMPTY(s), ring.phase(r) ^ 1u);
          if (uk == 0) WS_TRACE(0, ptl);
          const uint32_t dst = smem_u32(stage0) + (uint32_t)s * (uint32_t)p.stage_bytes;
          mbar_expect_tx(FULL(s), (p.dbg & 4) ? 0u : p.tx_bytes);
          if (!(p.dbg & 4)) {
            for (int kgi = 0; kgi < p.kgu; kgi++) {
              if (p.tma_mode == 2) {                       // flat: one box of 2^fshift positions at a time (rows past the end are filled)
                for (int b = 0; b < (p.L >> p.fshift); b++)
                  tma_load_4d(dst + (uint32_t)(kgi * p.sub_bytes) + (uint32_t)((b << p.fshift) * p.pitch), &tmap, (uk * p.kgu + kgi) * p.Kc,
                              c1 + (b << p.fshift), 0, 0, FULL(s));
              } else {
                tma_load_4d(dst + (uint32_t)(kgi * p.sub_bytes), &tmap, (uk * p.kgu + kgi) * p.Kc, c1, c2, c3, FULL(s));
              }
            }
            if (!p.w_resident) bulk_g2s(dst + (uint32_t)p.a_unit_bytes, wp + (size_t)uk * p.b_unit_bytes, p.b_unit_bytes, FULL(s));
          }
          if (ring.advance(r, p)) { if (r) wrapped1 = true; else wrapped0 = true; }
        }
        WS_TRACE(1, ptl);
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ================= MMA issuers: warp-uniform control flow (uniform datapath), one elected lane issues.  With two
    // issuers warp 1 takes the even tiles and warp 2 the odd ones: one warp's issue overhead hides behind the other's MMAs
    const int me = warp - 1;
    if (me < p.niss) {
      const bool leader = elect_one();
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.Npad >> 3) << 17) | (8u << 24);
      const uint32_t bstep = (2u * (uint32_t)p.Npad * 16u) >> 4;           // one K=16 step of the packed weights
      const uint32_t btap = ((uint32_t)p.Kc * (uint32_t)p.Npad * 2u) >> 4;  // one (K-group, tap) block
      const uint32_t mbstep = (128u * (uint32_t)p.pitch) >> 4;             // one 128-row accumulator block of A
      const uint32_t a_hi = desc_sw_hi((uint32_t)p.pitch);
      const uint32_t b_hi = (128u >> 4) | (1u << 14);                       // SBO = 128 B, version 1, no swizzle
      const uint32_t b_lbo = (((uint32_t)p.Npad * 16u) >> 4) << 16;
      const int nks = p.Kc >> 4;
      if (p.w_resident) mbar_wait(WBAR, 0);
      int tl = 0, si = 0;
      uint32_t ph = 0;
      WsTile tc;
      for (tc.init(p); tc.valid(p); tc.next(p), tl++) {
        if (p.niss == 2 && (tl & 1) != me) continue;
        const int buf = tl % p.nbuf, use = tl / p.nbuf;
        if (leader) WS_TRACE(2, tl);
        if (use >= 1) mbar_wait(TEMPTY(buf), (uint32_t)(use - 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (leader) WS_TRACE(3, tl);
        const uint32_t d_tile = tmem_base + (uint32_t)buf * buf_cols;
        const int nmb = (p.dbg & 1) ? 0 : tc.nmb;
        for (int uk = 0; uk < p.upt; uk++) {
          const int s = me + p.niss * si;
          mbar_wait(p.has_aff ? READY(s) : FULL(s), ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (leader && uk == 0) WS_TRACE(4, tl);
          const uint32_t a_stage = smem_u32(stage0) + (uint32_t)s * (uint32_t)p.stage_bytes;
          const uint32_t b_lo0 = (((p.w_resident ? smem_u32(w_s) + (uint32_t)uk * p.b_unit_bytes : a_stage + (uint32_t)p.a_unit_bytes) & 0x3FFFFu) >> 4) | b_lbo;
          uint32_t started = 0u;                         // the first MMA of a tile overwrites the accumulators
          for (int kgi = 0; kgi < p.kgu; kgi++) {
            const uint32_t a_lo0 = (((a_stage + (uint32_t)(kgi * p.sub_bytes)) & 0x3FFFFu) >> 4) | (1u << 16);
            const uint32_t first_k = (uint32_t)(uk | kgi);
            // taps walked as (dd, dy, dx) with running descriptor offsets: an indexed constant load per tap (p.tapdesc[tap]) sits
            // ~190 cycles on the issue path, which starves the tensor pipe on small tiles (FFPN_WS_TRACE, everything-off mode)
            int tx = 0, ty = 0;
            uint32_t t_off = p.tapdesc[0], t_row = p.tapdesc[0], t_slice = p.tapdesc[0];
            uint32_t b_lo1 = b_lo0 + (uint32_t)(kgi * ntaps) * btap;
            for (int tap = 0; tap < ntaps; tap++, b_lo1 += btap) {
              const uint32_t a_lo1 = a_lo0 + t_off;
              // pair view: the outer pair taps only read one element of the pair -> half of their K steps are all-zero weights
              const int t3 = p.pair2 ? tx : 1;
              const int ks_lo = t3 == 0 ? (nks >> 1) : 0, ks_hi = t3 == 2 ? (nks >> 1) : nks;
              // next tap's offset: dx fastest, then dy (one padded line), then dd (one slice region)
              if (++tx == p.kX) {
                tx = 0;
                if (++ty == p.kY) { ty = 0; t_slice += p.tap_dd; t_row = t_slice; } else { t_row += p.tap_dy; }
                t_off = t_row;
              } else {
                t_off += p.tap_dx;
              }
              for (int ks = ks_lo; ks < ks_hi; ks++) {
                const uint64_t bd = desc64(b_lo1 + (uint32_t)ks * bstep, b_hi);
                const uint32_t acc = (first_k | started) != 0 ? 1u : 0u;
                started = 1u;
                uint32_t a_lo = a_lo1 + (uint32_t)ks * 2u, dcol = d_tile;
#pragma unroll 8
                for (int mb = 0; mb < nmb; mb++) {
                  if (leader) umma_bf16(dcol, desc64(a_lo, a_hi), bd, idesc, acc);
                  a_lo += mbstep; dcol += (uint32_t)p.colstride;
                }
              }
            }
          }
          if (leader) umma_commit(EMPTY(s));            // stage s may be refilled once these MMAs have read it
          __syncwarp();
          if (++si == (me ? p.nst_ring1 : p.nst_ring0)) { si = 0; ph ^= 1u; }
        }
        if (leader) { WS_TRACE(5, tl); umma_commit(TFULL(buf)); WS_TRACE(6, tl); }   // accumulators of this tile complete
        __syncwarp();
      }
    }
  }
  } else if (warp < 4 + WS_NEPI) {
    reg_inc<WS_REGS_EPI>();
    // ================= epilogue =================
    // Two warps per TMEM lane quadrant.  SPLITC = false: they take ALTERNATE TILES (tile parity), so the two warps that share an
    // SM sub-partition are in different phases of their tcgen05.ld -> convert -> store chains and hide each other's latency.
    // SPLITC = true (N = 64 with statistics): both work on every tile and take half of the 16-column chunks each, which keeps
    // the per-thread statistics of 32 channels in registers instead of 2 x 16 warp shuffles per item.
    constexpr bool ADD = MODE == 1, BNR = MODE == 2, LD = MODE != 0;
    const int quad = warp & 3, half = (warp - 4) >> 2;
    const int etid = tid - 128;                                    // 0..255 over the epilogue warps
    // ---- row table: tile-local output position of every accumulator row, once per CTA (the decomposition of a row index into
    // slice / line / column does not depend on the tile; only the clipping against the tensor edges does) ----
    {
      const int nrows_tab = (int)(buf_cols / (uint32_t)p.colstride) * 128;
      for (int m = etid; m < nrows_tab; m += WS_NEPI * 32) {
        const int j = m / p.Lr, ii = m - j * p.Lr;
        const int oyl = p.tma_mode == 0 ? ii / p.Xp : 0, oxl = p.tma_mode == 0 ? ii - oyl * p.Xp : ii;
        const bool ok = ii < p.L && j < 256 && oyl < 4096 && oxl < 4096;       // ii >= L: rows between two slices of a tile
        const long long relo = (long long)j * p.outD + (long long)oyl * p.outY + oxl;
        row_tab[m] = ok ? make_uint2((uint32_t)relo, ((uint32_t)j << 24) | ((uint32_t)oyl << 12) | (uint32_t)oxl) : make_uint2(0xffffffffu, 0u);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(WS_NEPI * 32) : "memory");         // epilogue warps only
    }
    float* my_stat = stat_s + (warp - 4) * 2 * p.Npad;            // this warp's own slot: plain read-modify-write, no atomics
    float ssum[NREG > 0 ? 16 * NREG : 1], ssq[NREG > 0 ? 16 * NREG : 1];
#pragma unroll
    for (int i = 0; i < (NREG > 0 ? 16 * NREG : 1); i++) { ssum[i] = 0.f; ssq[i] = 0.f; }
    float bsc_r[BNR && NREG == 1 ? 16 : 1], bsh_r[BNR && NREG == 1 ? 16 : 1];   // MODE 2, N = 16: the BatchNorm coefficients stay in registers
    if (BNR && NREG == 1) {
#pragma unroll
      for (int q = 0; q < 16; q++) { bsc_r[BNR && NREG == 1 ? q : 0] = p.bsc[q]; bsh_r[BNR && NREG == 1 ? q : 0] = p.bsh[q]; }
    }
    const int nch_all = p.Npad >> 4;
    const int nch = NREG > 0 ? NREG : (SPLITC ? (nch_all >> 1) : nch_all);      // 16-column chunks this warp takes per accumulator block
    const int ch0 = SPLITC ? half * nch : 0;
    // Items (one 32-row x 16-column accumulator block each) are processed in half-batches of HB with TWO half-batches of TMEM
    // loads (+ their addend / y loads) in flight: while one is converted and stored the next one is already on its way.
    #ifndef WS_HB
#define WS_HB 2
#endif
    // Measured (B200, level-1 forward / dgrad): HB = 2 beats 3 and 4 (80.1 vs 85.8 us, 70.6 vs 78.8 us) although the registers would
    // allow them -- short dependent chains interleave better between the two warps of a scheduler.  MODE 2 also keeps y and the
    // BatchNorm coefficients: one item per half.
#ifndef WS_HB_BNR
#define WS_HB_BNR 1
#endif
    constexpr int HB = BNR ? WS_HB_BNR : WS_HB;
    int tl = 0;
    WsTile tc;
    for (tc.init(p); tc.valid(p); tc.next(p), tl++) {
      if (!SPLITC && (tl & 1) != half) continue;
      const int buf = tl % p.nbuf;
      if (warp == 4 && lane == 0) WS_TRACE(7, tl);
      const int nitems = ((p.dbg & 2) ? 0 : tc.nmb) * nch;
      // clipping of this tile against the tensor (slices, lines, columns, and the line end when X is cut into segments)
      const int ox_base = p.tma_mode == 0 ? 0 : tc.i0, oy_base = p.tma_mode == 0 ? tc.it * p.tY : 0;
      const int remD = tc.tD_t, remY = p.oY - oy_base, remX = min(p.oX, p.oXtot - tc.nb * p.xseg) - ox_base;
      const long long tbase = (long long)tc.nb * p.outNB + (long long)tc.d0 * p.outD + (long long)oy_base * p.outY + ox_base;
      const uint32_t t_tile = tmem_base + (uint32_t)buf * buf_cols + ((uint32_t)(quad * 32) << 16);
      if (LD) {
        // the addend / y rows of the NEXT tile this warp takes are pulled into L2 now: the per-item register loads (one half-batch
        // ahead, i.e. a few hundred ns) then find them there instead of paying the DRAM latency item by item
        WsTile nx = tc;
        nx.next(p);
        if (!SPLITC) nx.next(p);
        if (nx.valid(p)) {
          const int nox = p.tma_mode == 0 ? 0 : nx.i0, noy = p.tma_mode == 0 ? nx.it * p.tY : 0;
          const int nremD = nx.tD_t, nremY = p.oY - noy, nremX = min(p.oX, p.oXtot - nx.nb * p.xseg) - nox;
          const long long nbase = (long long)nx.nb * p.outNB + (long long)nx.d0 * p.outD + (long long)noy * p.outY + nox;
          const int nit = nx.nmb * nch;
          for (int idx = 0; idx < nit; idx++) {
            const int mbi = NREG > 0 ? idx / NREG : idx / nch, ch = ch0 + (NREG > 0 ? idx % NREG : idx - mbi * nch);
            const uint2 e = row_tab[mbi * 128 + quad * 32 + lane];
            const int j = (int)(e.y >> 24), oyl = (int)((e.y >> 12) & 0xfffu), oxl = (int)(e.y & 0xfffu);
            const int cbase = n0 + ch * 16;
            if ((e.x != 0xffffffffu) && (j < nremD) && (oyl < nremY) && (oxl < nremX) && cbase < p.Cout)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(p.addend + (nbase + (long long)e.x) * p.Cout + cbase));
          }
        }
      }
      uint32_t raw[2][HB][16];
      long long oposv[2][HB];
      bool validv[2][HB];
      uint4 addv[LD ? 2 : 1][LD ? HB : 1][2];
      float fa[LD ? HB : 1][LD ? 16 : 1];                 // the current half-batch's addend / y, unpacked (see consume)
      // output position of every item of a half-batch (+ its addend / y loads, issued long before they are used)
      auto prep = [&](const int hh, const int it0, const uint32_t dep) {
#pragma unroll
        for (int u = 0; u < HB; u++) {
          const int idx = it0 + u;
          validv[hh][u] = false; oposv[hh][u] = 0;
          if (LD) addv[LD ? hh : 0][LD ? u : 0][0] = addv[LD ? hh : 0][LD ? u : 0][1] = make_uint4(0u, 0u, 0u, 0u);
          if (idx < nitems) {
            const int mbi = NREG > 0 ? idx / NREG : idx / nch, ch = ch0 + (NREG > 0 ? (hh * HB + u) % NREG : idx - mbi * nch);
            const uint2 e = row_tab[mbi * 128 + quad * 32 + lane];
            const int j = (int)(e.y >> 24), oyl = (int)((e.y >> 12) & 0xfffu), oxl = (int)(e.y & 0xfffu);
            validv[hh][u] = (e.x != 0xffffffffu) && (j < remD) && (oyl < remY) && (oxl < remX);
            oposv[hh][u] = tbase + (long long)e.x;
            if (LD && validv[hh][u]) {
              const int cbase = n0 + ch * 16;
              const uint4* ap = reinterpret_cast<const uint4*>(p.addend + oposv[hh][u] * p.Cout + cbase + dep);
              if (cbase + 16 <= p.Cout) {
                uint32_t t8[8];
                ld_global_v8(ap, t8);
                addv[LD ? hh : 0][LD ? u : 0][0] = make_uint4(t8[0], t8[1], t8[2], t8[3]);
                addv[LD ? hh : 0][LD ? u : 0][1] = make_uint4(t8[4], t8[5], t8[6], t8[7]);
              } else if (cbase < p.Cout) {
                addv[LD ? hh : 0][LD ? u : 0][0] = ap[0];
              }
            }
          }
        }
      };
      auto issue = [&](const int hh, const int it0) {
#pragma unroll
        for (int u = 0; u < HB; u++) {
          const int idx = it0 + u;
          if (idx < nitems) {
            const int mbi = NREG > 0 ? idx / NREG : idx / nch, ch = ch0 + (NREG > 0 ? (hh * HB + u) % NREG : idx - mbi * nch);
            tmem_ld16_nowait(t_tile + (uint32_t)(mbi * p.colstride + ch * 16), raw[hh][u]);
          }
        }
      };
      // All addend / y loads of this warp share one hardware scoreboard, and waiting on a scoreboard waits for EVERY load
      // outstanding on it: if the registers of half-batch k were first read inside process(k), i.e. after the loads of k + 1
      // have been issued, each half-batch would wait for the loads it has only just issued and nothing would overlap (measured:
      // the addend dgrad at level 1 took 130 us against 70 us without addend, and MORE loads in flight made it slower).  So the
      // loaded words are read BEFORE the next loads go out, one whole process() after they were issued themselves.  Neither the
      // compiler nor ptxas would keep that order on their own (the reads sink to their use, the loads float up), so the address
      // of the next loads is made to depend on the words just read: they are folded, ANDed with a kernel parameter that is
      // always zero but unknown at compile time, and added to the address.
      auto consume = [&](const int hh) -> uint32_t {
        uint32_t fold = 0u;
#pragma unroll
        for (int u = 0; u < HB; u++) {
#pragma unroll
          for (int h2 = 0; h2 < 2; h2++) {
            const uint4 a4 = addv[LD ? hh : 0][LD ? u : 0][h2];
            const uint32_t w4[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
            for (int q = 0; q < 4; q++) {
              // the UNPACKED words feed the dependency and are then made opaque to the compiler: nothing after this point can be
              // derived from the load's destination registers again (a later read of those would wait on the scoreboard anew)
              uint32_t lo = w4[q] << 16, hi = w4[q] & 0xffff0000u;
              asm volatile("" : "+r"(lo), "+r"(hi));
              fold ^= lo ^ hi;
              fa[LD ? u : 0][LD ? h2 * 8 + 2 * q : 0] = __uint_as_float(lo);
              fa[LD ? u : 0][LD ? h2 * 8 + 2 * q + 1 : 0] = __uint_as_float(hi);
            }
          }
        }
        return fold & p.zero;                                         // 0 -- but a true data dependency of the next loads' addresses
      };
      auto process = [&](const int hh, const int it0) {
#pragma unroll
        for (int u = 0; u < HB; u++) {
          const int idx = it0 + u;
          if (idx < nitems) {
            const int mbi = NREG > 0 ? idx / NREG : idx / nch, ch = ch0 + (NREG > 0 ? (hh * HB + u) % NREG : idx - mbi * nch);
            const bool valid = validv[hh][u];
            const long long opos = oposv[hh][u];
            float v[16];
#pragma unroll
            for (int q = 0; q < 16; q++) v[q] = __uint_as_float(raw[hh][u][q]);
            const int cbase = n0 + ch * 16;
            if (ADD && valid) {                                      // (words that were not loaded are zero: prep)
#pragma unroll
              for (int q = 0; q < 16; q++) v[q] += fa[LD ? u : 0][LD ? q : 0];
            }
            uint32_t packed[8];
#pragma unroll
            for (int q = 0; q < 8; q++) {
              __nv_bfloat162 hh2 = __floats2bfloat162_rn(v[2 * q], v[2 * q + 1]);
              packed[q] = *reinterpret_cast<uint32_t*>(&hh2);
            }
            if (valid && !(p.dbg & 16)) {
              uint4* yp = reinterpret_cast<uint4*>(p.y + opos * p.Cout + cbase);
              if (cbase + 16 <= p.Cout) st_global_v8(yp, packed);                  // one full sector per lane
              else if (cbase < p.Cout) yp[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            }
            if (p.has_stats && !(p.dbg & 32)) {
              // statistics of the stored (rounded) values: sum v and sum v * v (BatchNorm forward); MODE 2: sum G and sum G * y with
              // G = v under the ReLU mask of the producer BatchNorm (same expression as bn_bwd_reduce / bn_bwd_apply -- the mask only
              // enters the sums, the stored gradient is the plain dgrad).  Rows outside the tensor are zeroed word by word.
              float w2[16];
              float bs[BNR ? 16 : 1], bh[BNR ? 16 : 1];
              if (BNR) {
                if (NREG == 1) {
#pragma unroll
                  for (int q = 0; q < 16; q++) { bs[BNR ? q : 0] = bsc_r[BNR && NREG == 1 ? q : 0]; bh[BNR ? q : 0] = bsh_r[BNR && NREG == 1 ? q : 0]; }
                } else {
                  const int cb = p.bn_mod ? cbase % p.bn_mod : cbase;
                  const float4* scp = reinterpret_cast<const float4*>(p.bsc + cb);
                  const float4* shp = reinterpret_cast<const float4*>(p.bsh + cb);
#pragma unroll
                  for (int q = 0; q < 4; q++) {
                    const float4 s4 = __ldg(scp + q), h4 = __ldg(shp + q);
                    bs[BNR ? 4 * q : 0] = s4.x; bs[BNR ? 4 * q + 1 : 0] = s4.y; bs[BNR ? 4 * q + 2 : 0] = s4.z; bs[BNR ? 4 * q + 3 : 0] = s4.w;
                    bh[BNR ? 4 * q : 0] = h4.x; bh[BNR ? 4 * q + 1 : 0] = h4.y; bh[BNR ? 4 * q + 2 : 0] = h4.z; bh[BNR ? 4 * q + 3 : 0] = h4.w;
                  }
                }
              }
#pragma unroll
              for (int q = 0; q < 8; q++) {
                const uint32_t pk = valid ? packed[q] : 0u;
                float g0 = __uint_as_float(pk << 16), g1 = __uint_as_float(pk & 0xffff0000u);
                if (BNR) {
                  const float y0 = fa[LD ? u : 0][LD ? 2 * q : 0], y1 = fa[LD ? u : 0][LD ? 2 * q + 1 : 0];
                  g0 = fmaf(y0, bs[BNR ? 2 * q : 0], bh[BNR ? 2 * q : 0]) > 0.f ? g0 : 0.f;
                  g1 = fmaf(y1, bs[BNR ? 2 * q + 1 : 0], bh[BNR ? 2 * q + 1 : 0]) > 0.f ? g1 : 0.f;
                  w2[2 * q] = g0 * y0; w2[2 * q + 1] = g1 * y1;
                } else {
                  w2[2 * q] = g0 * g0; w2[2 * q + 1] = g1 * g1;
                }
                v[2 * q] = g0; v[2 * q + 1] = g1;
              }
              if (NREG > 0) {
                constexpr int CH = NREG > 0 ? NREG : 1;
#pragma unroll
                for (int q = 0; q < 16; q++) {
                  ssum[16 * ((hh * HB + u) % CH) + q] += v[q];
                  if (BNR) ssq[16 * ((hh * HB + u) % CH) + q] += w2[q];
                  else ssq[16 * ((hh * HB + u) % CH) + q] = fmaf(v[q], v[q], ssq[16 * ((hh * HB + u) % CH) + q]);
                }
              } else {
                const float a = warp_transpose_sum16(v, lane);
                const float b = warp_transpose_sum16(w2, lane);
                if ((lane & 1) == 0) {
                  const int c = ch * 16 + transpose_sum_channel(lane);
                  my_stat[c] += a;
                  my_stat[p.Npad + c] += b;
                }
              }
            }
          }
        }
      };
      // the first half-batch's positions and addend / y loads go out BEFORE the wait for the accumulators: their DRAM latency hides
      // behind the MMAs
      prep(0, 0, 0u);
      mbar_wait(TFULL(buf), (uint32_t)(tl / p.nbuf) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (warp == 4 && lane == 0) WS_TRACE(8, tl);
      issue(0, 0);
      for (int it0 = 0; it0 < nitems; it0 += 2 * HB) {
#pragma unroll
        for (int hh = 0; hh < 2; hh++) {
          const int cur = it0 + hh * HB;
          if (cur < nitems) {                                         // warp-uniform
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");   // half-batch hh has landed (nothing else is outstanding)
            uint32_t dep = 0u;
            if (LD) dep = consume(hh);
            if (cur + HB < nitems) { prep(hh ^ 1, cur + HB, dep); issue(hh ^ 1, cur + HB); }
            process(hh, cur);
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (warp == 4 && lane == 0) WS_TRACE(9, tl);
      if (lane == 0) mbar_arrive(TEMPTY(buf));
    }
    if (NREG > 0 && p.has_stats) {
#pragma unroll
      for (int chl = 0; chl < (NREG > 0 ? NREG : 1); chl++) {
        float a[16], b[16];
#pragma unroll
        for (int q = 0; q < 16; q++) { a[q] = ssum[16 * chl + q]; b[q] = ssq[16 * chl + q]; }
        const float ta = warp_transpose_sum16(a, lane);
        const float tb = warp_transpose_sum16(b, lane);
        if ((lane & 1) == 0) {
          const int c = (ch0 + chl) * 16 + transpose_sum_channel(lane);
          my_stat[c] += ta;
          my_stat[p.Npad + c] += tb;
        }
      }
    }
  } else {
    reg_dec<WS_REGS_XFORM>();
    // ================= in-place BatchNorm scale/shift + ReLU of landed tiles =================
    if (p.has_aff) {
      const int tix = tid - 384;                                   // 0..WS_NT-1
      const int cpu = p.kgu * (p.Kc >> 3);                         // 16-byte channel chunks per row of a unit
      const int rstep = WS_NT / cpu;
      const bool active = tix < rstep * cpu;
      const int c = tix % cpu, r0 = tix / cpu;
      const int kgi = c / (p.Kc >> 3), cc = c % (p.Kc >> 3);
      const uint32_t cmask = (uint32_t)(p.Kc >> 3) - 1u;
      float s[8], h[8];
      WsRing ring;
      ring.init();
      bool loaded = false;
      WsTile tc;
      int ttl = 0;
      for (tc.init(p); tc.valid(p); tc.next(p), ttl++) {
        const int r = p.niss == 2 ? (ttl & 1) : 0;
        for (int uk = 0; uk < p.upt; uk++) {
          if (active && (!loaded || p.upt > 1)) {
            int cofs = (uk * p.kgu + kgi) * p.Kc + cc * 8;
            if (p.aff_mod) cofs %= p.aff_mod;
            const float4 s0 = *reinterpret_cast<const float4*>(p.sc + cofs), s1 = *reinterpret_cast<const float4*>(p.sc + cofs + 4);
            const float4 h0 = *reinterpret_cast<const float4*>(p.sh + cofs), h1 = *reinterpret_cast<const float4*>(p.sh + cofs + 4);
            s[0] = s0.x; s[1] = s0.y; s[2] = s0.z; s[3] = s0.w; s[4] = s1.x; s[5] = s1.y; s[6] = s1.z; s[7] = s1.w;
            h[0] = h0.x; h[1] = h0.y; h[2] = h0.z; h[3] = h0.w; h[4] = h1.x; h[5] = h1.y; h[6] = h1.z; h[7] = h1.w;
            loaded = true;
          }
          const int st = ring.stage(r, p);
          mbar_wait(FULL(st), ring.phase(r));
          if (tix == 0 && uk == 0) WS_TRACE(10, ttl);
          if (active && !(p.dbg & 8)) {
            uint8_t* base = stage0 + (size_t)st * p.stage_bytes + (size_t)kgi * p.sub_bytes;
            const uint32_t abase = smem_u32(base);                       // the swizzle is a function of the absolute address
            int rr = r0;
            if (((rstep * p.pitch) & 1023) == 0) {
              // a thread's rows are a whole number of swizzle periods (1 KiB) apart: its 16-byte chunk sits at the same swizzled
              // offset in every one of them, so the loop is a pointer walk -- 8 independent chunks in flight per thread (the four
              // transform warps have a scheduler each: the loop is bound by dependent-issue latency, not by issue slots)
              const uint32_t off0 = (uint32_t)r0 * (uint32_t)p.pitch;
              uint4* q0 = reinterpret_cast<uint4*>(base + off0 + ((((uint32_t)cc ^ ((abase + off0) >> 7)) & cmask) << 4));
              const int qstep = (rstep * p.pitch) >> 4;
              const int n = (p.region_rows - r0 + rstep - 1) / rstep;
              int i = 0;
              for (; i + 8 <= n; i += 8, q0 += 8 * qstep) {
                uint4 v[8];
#pragma unroll
                for (int u = 0; u < 8; u++) v[u] = q0[u * qstep];
#pragma unroll
                for (int u = 0; u < 8; u++) q0[u * qstep] = bn_relu_bf16x8(v[u], s, h, 1);
              }
              for (; i < n; i++, q0 += qstep) *q0 = bn_relu_bf16x8(*q0, s, h, 1);
              rr = p.region_rows;
            }
            for (; rr + 3 * rstep < p.region_rows; rr += 4 * rstep) {
              uint4* q[4];
              uint4 v[4];
#pragma unroll
              for (int u = 0; u < 4; u++) {
                const uint32_t off = (uint32_t)(rr + u * rstep) * (uint32_t)p.pitch;
                q[u] = reinterpret_cast<uint4*>(base + off + ((((uint32_t)cc ^ ((abase + off) >> 7)) & cmask) << 4));
                v[u] = *q[u];
              }
#pragma unroll
              for (int u = 0; u < 4; u++) *q[u] = bn_relu_bf16x8(v[u], s, h, 1);
            }
            for (; rr < p.region_rows; rr += rstep) {
              const uint32_t off = (uint32_t)rr * (uint32_t)p.pitch;
              uint4* q = reinterpret_cast<uint4*>(base + off + ((((uint32_t)cc ^ ((abase + off) >> 7)) & cmask) << 4));
              *q = bn_relu_bf16x8(*q, s, h, 1);
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(READY(st));         // one arrival per warp: 128 arrivals on one barrier serialise in shared memory
          if (tix == 0 && uk == p.upt - 1) WS_TRACE(11, ttl);
          ring.advance(r, p);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (p.has_stats && p.fold) {
    // pair view: columns c and c + fold are the same channel (even / odd position) -> one statistics column
    for (int i = tid; i < 2 * p.fold; i += WS_THREADS) {
      const int which = i / p.fold, c = i - which * p.fold;
      float t = 0.f;
#pragma unroll
      for (int e = 0; e < WS_NEPI; e++) t += stat_s[e * 2 * p.Npad + which * p.Npad + c] + stat_s[e * 2 * p.Npad + which * p.Npad + c + p.fold];
      p.stat[((size_t)blockIdx.x * 2 + which) * p.fold + c] = t;
    }
  } else if (p.has_stats) {
    for (int i = tid; i < 2 * p.Npad; i += WS_THREADS) {
      const int which = i / p.Npad, c = i - which * p.Npad;
      float t = 0.f;
#pragma unroll
      for (int e = 0; e < WS_NEPI; e++) t += stat_s[e * 2 * p.Npad + i];
      if (n0 + c < p.Cout) p.stat[((size_t)blockIdx.x * 2 + which) * p.Cout + n0 + c] = t;
    }
  }
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols));
  }
  if (p.fin_on) {
    // ---- fused BatchNorm finalize: the last CTA to arrive combines all partial rows (fp64, fixed order) ----
    int* flag = reinterpret_cast<int*>(smem + 248);
    __threadfence();
    __syncthreads();
    if (tid == 0) *flag = (atomicAdd(p.fin_counter, 1u) == gridDim.x * gridDim.y - 1u) ? 1 : 0;
    __syncthreads();
    if (*flag) {
      __threadfence();
      double* red = reinterpret_cast<double*>(stat_s);                   // [16 row lanes][32 channels][2]
      const int rows = (int)gridDim.x, rl = tid >> 5, cl = tid & 31;
      for (int c0 = 0; c0 < p.Cout; c0 += 32) {
        const int c = c0 + cl;
        double a = 0.0, b = 0.0;
        if (c < p.Cout) {
          // grid.x <= 160 rows -> at most 10 per row lane: issue all loads first (one L2 round trip), then add in order
          float va[10], vb[10];
#pragma unroll
          for (int u = 0; u < 10; u++) {
            const int r = rl + 16 * u;
            va[u] = r < rows ? __ldcg(p.stat + ((size_t)r * 2 + 0) * p.Cout + c) : 0.f;
            vb[u] = r < rows ? __ldcg(p.stat + ((size_t)r * 2 + 1) * p.Cout + c) : 0.f;
          }
#pragma unroll
          for (int u = 0; u < 10; u++) { a += (double)va[u]; b += (double)vb[u]; }
          for (int r = rl + 160; r < rows; r += 16) {
            a += (double)__ldcg(p.stat + ((size_t)r * 2 + 0) * p.Cout + c);
            b += (double)__ldcg(p.stat + ((size_t)r * 2 + 1) * p.Cout + c);
          }
        }
        __syncthreads();
        red[(rl * 32 + cl) * 2 + 0] = a;
        red[(rl * 32 + cl) * 2 + 1] = b;
        __syncthreads();
        if (rl == 0 && c < p.Cout) {
          double s1 = 0.0, s2 = 0.0;
          for (int l = 0; l < 16; l++) { s1 += red[(l * 32 + cl) * 2 + 0]; s2 += red[(l * 32 + cl) * 2 + 1]; }
          const double count = p.fin.count;
          const double mean = s1 / count;
          double var = s2 / count - mean * mean;
          if (var < 0.0) var = 0.0;
          const double invstd = 1.0 / sqrt(var + (double)p.fin.eps);
          const double gm = (double)p.fin.gamma[c];
          p.fin.scale[c] = (float)(gm * invstd);
          p.fin.shift[c] = (float)((double)p.fin.beta[c] - mean * gm * invstd);
          if (p.fin.save_mean) p.fin.save_mean[c] = (float)mean;
          if (p.fin.save_invstd) p.fin.save_invstd[c] = (float)invstd;
          if (p.fin.running_mean != nullptr) {
            const double mom = (double)p.fin.momentum;
            const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
            p.fin.running_mean[c] = (float)((1.0 - mom) * (double)p.fin.running_mean[c] + mom * mean);
            p.fin.running_var[c] = (float)((1.0 - mom) * (double)p.fin.running_var[c] + mom * unb);
          }
        }
      }
      if (tid == 0) *p.fin_counter = 0u;                                 // self-resetting for the next launch
    }
  }
}

struct WsPlan {
  WsParams p;
  Plan base;
  size_t smem;
  dim3 grid;
  int nchunks;
  bool ok;
  int box[4];
};

WsPlan make_ws_plan_with(const ffpn_conv_desc* d, bool transposed, int num_sms, int npad_override, int nmb_limit) {
  WsPlan w;
  memset(&w, 0, sizeof(w));
  w.base = ffpn_tc_make_plan(d, transposed, num_sms);
  if (!w.base.ok) return w;
  const TcParams& c = w.base.p;
  if (c.sX != 1 || c.nsets != 1) return w;
  WsParams& p = w.p;
  p.NB = c.NB; p.D = c.D; p.Y = c.Y; p.X = c.X; p.oD = c.oD; p.oY = c.oY; p.oX = c.oX;
  p.kD = c.kD; p.kY = c.kY; p.kX = c.kX; p.pD = c.pD; p.pY = c.pY; p.pX = c.pX; p.hl = c.hl;
  p.outNB = c.outNB; p.outD = c.outD; p.outY = c.outY;
  p.Cin = c.Cin; p.Cout = c.Cout; p.Npad = c.Npad; p.Xp = c.Xp; p.Qout = c.Qout;
  w.nchunks = w.base.nchunks;
  {  // wide outputs: split N over blockIdx.y so that each CTA streams half the weights and more SMs take part
    static int nsplit = -1;
    if (nsplit < 0) { const char* e = getenv("FFPN_WS_NSPLIT"); nsplit = e ? atoi(e) : 128; }
    if (nsplit > 0 && p.Cout >= 2 * nsplit && p.Cout % nsplit == 0) { p.Npad = nsplit; w.nchunks = p.Cout / nsplit; }
    if (npad_override > 0 && npad_override < p.Npad && p.Cout % npad_override == 0) { p.Npad = npad_override; w.nchunks = p.Cout / npad_override; }
  }
  const int ntaps = p.kD * p.kY * p.kX;
  if (ntaps > 27 || p.Cin % 16 != 0) return w;
  p.xseg = 0; p.oXtot = p.oX;
  if (p.kD == 1 && p.Xp > 256 && p.NB == 1 && !(p.Y == 1 && p.D == 1 && p.kY == 1 && p.kX == 1)) {
    // a padded line does not fit one TMA box (256 positions): cut X into nX segments.  A segment is a tile column with
    // its own halo (real neighbour data inside the image, TMA fill outside) and is walked like a batch entry: the TMA x
    // coordinate and the output position both advance by nb * xseg.
    const int halo = p.Xp - p.oX;
    const int nX = (p.oX + (256 - halo) - 1) / (256 - halo);
    const int tX = (p.oX + nX - 1) / nX;
    p.xseg = tX; p.NB = nX; p.outNB = tX; p.oX = tX; p.Xp = tX + halo; p.Qout = (p.oY - 1) * p.Xp + p.oX;
  }
  const int hr = p.Xp - p.oX;
  const int maxinner = (p.kY - 1) * p.Xp + hr;
  static int cs16 = -1;
  if (cs16 < 0) { const char* e = getenv("FFPN_WS_CS32"); cs16 = (e && atoi(e)) ? 0 : 1; }
  p.colstride = (p.Npad < 32 && !cs16) ? 32 : p.Npad;
  p.nbuf = p.colstride <= 64 ? 4 : 2;                    // TMEM tile buffers: deep enough that the MMAs never wait for the epilogue
  { const char* e = getenv("FFPN_WS_NBUF"); if (e && (atoi(e) == 2 || atoi(e) == 4) && (512 / atoi(e)) >= p.colstride) p.nbuf = atoi(e); }   // tuning
  int nmb_cap = (512 / p.nbuf) / p.colstride;
  if (nmb_cap > 8) nmb_cap = 8;
  { const char* e = getenv("FFPN_WS_NMBCAP"); if (e && atoi(e) >= 1 && atoi(e) < nmb_cap) nmb_cap = atoi(e); }                                // tuning
  if (nmb_limit > 0 && nmb_limit < nmb_cap) nmb_cap = nmb_limit;
  if (nmb_cap < 1) return w;
  const size_t budget = 227 * 1024 - WS_HDR - 1024;
  const size_t w_total = (size_t)ntaps * p.Cin * p.Npad * 2;
  // option list: resident weights with the widest swizzle row, then streamed weights with 64/32/16-channel K-groups
  struct Opt { int resident, Kc; };
  Opt opts[4];
  int nopts = 0;
  {
    int kc = 64;
    while (kc > 16 && p.Cin % kc != 0) kc >>= 1;
    if (w_total <= 100 * 1024) opts[nopts++] = {1, kc};
    for (int k2 = 64; k2 >= 16; k2 >>= 1)
      if (p.Cin % k2 == 0) opts[nopts++] = {0, k2};
  }
  for (int oi = 0; oi < nopts; oi++) {
    for (int want_stages = 3; want_stages >= 2; want_stages--) {
      const int Kc = opts[oi].Kc, resident = opts[oi].resident;
      const int nkg = p.Cin / Kc, kgu = resident ? nkg : 1, upt = resident ? 1 : nkg;
      const int pitch = Kc * 2;
      const size_t b_unit = (size_t)ntaps * Kc * kgu * p.Npad * 2;
      if (!resident && b_unit > 120 * 1024) continue;
      const size_t avail = budget - (resident ? ((w_total + 1023) & ~(size_t)1023) : 0);
      for (int nmb = nmb_cap; nmb >= 1; nmb--) {
        const int max_rows = nmb * 128;
        int tD = 1, tY = 0, L = 0, Lr = 0, region = 0, mode = -1;
        if (p.kD > 1) {
          L = p.X < 128 ? p.X : 128; Lr = L;
          tD = max_rows / L; if (tD > p.oD) tD = p.oD; if (tD < 1) tD = 1;
          region = (tD + p.kD - 1) * L; mode = 1;
          if (tD + p.kD - 1 > 256 || L > 256) continue;
        } else if (p.Y == 1 && p.D == 1 && p.kY == 1 && p.kX == 1) {
          // flat 1x1x1: the positions are ONE tensor-map dimension, loaded unit (<= 256, the TMA box limit) rows at a time;
          // a ragged tail (X not a multiple of the unit) is filled by the TMA unit and clipped by the epilogue
          const int unit = (max_rows >= 256 && p.X >= 256) ? 256 : 128;
          int nblk = max_rows / unit; if (nblk < 1) continue;
          if (nblk * unit > p.X) nblk = (p.X + unit - 1) / unit;
          L = nblk * unit; Lr = L; tD = 1; region = L; mode = 2;
          p.fshift = unit == 256 ? 8 : 7;
        } else {
          if (p.Xp > 256) break;
          tY = max_rows / p.Xp;
          if (tY < 1) continue;
          if (tY >= p.oY) {
            tY = p.oY;
            Lr = (tY + p.kY - 1) * p.Xp;
            tD = (max_rows - tY * p.Xp) / Lr + 1; if (tD > p.oD) tD = p.oD; if (tD < 1) tD = 1;
          } else {
            Lr = (tY + p.kY - 1) * p.Xp; tD = 1;
          }
          L = tY * p.Xp; region = tD * Lr; mode = 0;
          if (tY + p.kY - 1 > 256 || tD > 256) continue;
        }
        const int M_total = (tD - 1) * Lr + L;
        const int nmb_t = (M_total + 127) / 128;
        const int maxoff = (p.kD - 1) * Lr + maxinner;
        int rows_alloc = nmb_t * 128 + maxoff;
        if (rows_alloc < region) rows_alloc = region;
        const size_t sub = ((size_t)rows_alloc * pitch + 1023) & ~(size_t)1023;
        const size_t a_unit = sub * kgu;
        const size_t stage = a_unit + (resident ? 0 : ((b_unit + 1023) & ~(size_t)1023));
        int nst = (int)(avail / stage);
        if (nst > WS_MAX_STAGES) nst = WS_MAX_STAGES;
        if (nst < want_stages) continue;
        { static int cap = -1; if (cap < 0) { const char* e = getenv("FFPN_WS_STAGES"); cap = e ? atoi(e) : 4; } if (nst > cap) nst = cap; }   // tuning
        // two issuer warps need a stage ring each (tile parity): 2 + 2 stages, or 1 + 1 when a tile is one unit
        p.niss = (nst >= 4 || (nst >= 2 && upt == 1)) ? 2 : 1;
        p.nst_ring0 = p.niss == 2 ? (nst + 1) / 2 : nst;
        p.nst_ring1 = p.niss == 2 ? nst / 2 : 0;
        p.tD = tD; p.tY = tY; p.L = L; p.Lr = Lr; p.tma_mode = mode;
        p.Kc = Kc; p.nkg = nkg; p.pitch = pitch; p.kgu = kgu; p.upt = upt;
        p.region_rows = region; p.sub_bytes = (int)sub; p.a_unit_bytes = (int)a_unit; p.stage_bytes = (int)stage; p.nstages = nst;
        p.w_resident = resident; p.b_unit_bytes = (unsigned)b_unit; p.b_total_bytes = (unsigned)w_total;
        p.tx_bytes = (unsigned)((size_t)kgu * region * pitch + (resident ? 0 : b_unit));
        int cols = nmb_t * p.colstride * p.nbuf, tc = 32;
        while (tc < cols) tc <<= 1;
        if (tc > 512) continue;
        p.tmem_cols = tc;
        for (int t = 0; t < ntaps; t++) {
          const int dx = t % p.kX, dy = (t / p.kX) % p.kY, dd = t / (p.kX * p.kY);
          p.tapdesc[t] = (unsigned)((dd * Lr + dy * p.Xp + dx - p.pX + p.hl) * pitch) >> 4;
        }
        p.tap_dx = (unsigned)pitch >> 4; p.tap_dy = (unsigned)(p.Xp * pitch) >> 4; p.tap_dd = (unsigned)(Lr * pitch) >> 4;
        p.nD = (p.oD + tD - 1) / tD;
        p.nI = (p.Qout + L - 1) / L;
        const int ntiles = p.NB * p.nD * p.nI;
        int gx = num_sms / w.nchunks; if (gx < 1) gx = 1;
        if (gx > ntiles) gx = ntiles;
        if (gx > FFPN_STAT_ROWS) gx = FFPN_STAT_ROWS;
        w.grid = dim3(gx, w.nchunks);
        w.smem = WS_HDR + (resident ? ((w_total + 1023) & ~(size_t)1023) : 0) + (size_t)nst * stage;
        w.box[0] = Kc;
        if (mode == 0) { w.box[1] = p.Xp; w.box[2] = tY + p.kY - 1; w.box[3] = tD; }
        else if (mode == 1) { w.box[1] = L; w.box[2] = tD + p.kD - 1; w.box[3] = 1; }
        else { w.box[1] = 1 << p.fshift; w.box[2] = 1; w.box[3] = 1; }
        w.ok = true;
        return w;
      }
    }
  }
  return w;
}

// Small problems (deep encoder levels, the en-face decoder): the default plan takes the largest tile and the widest N chunk,
// which leaves most of the 148 SMs idle (up_concat4: 16 tiles x 1 chunk, every CTA streaming 1.8 MB of weights).  With
// FFPN_WS_SPREAD=1, when the default plan gives fewer CTAs than SMs, smaller tiles and narrower N chunks (>= 32 columns: a K=16
// MMA costs ~40 cycles for any N <= 32) are tried and the plan with the most CTAs wins.  Measured on B200 (same box): alone,
// up_concat4 forward 59.2 -> 35.5 us and level-5 31.5 -> 28.4 us; but the C2 training step gets SLOWER, 9.33 -> 9.44 ms, because
// there these kernels already overlap the large level-1/2 kernels of the other branch streams and the extra CTAs take SMs
// away from them.  Hence off by default; it is the right setting for a single-stream / small-batch inference use.
WsPlan make_ws_plan(const ffpn_conv_desc* d, bool transposed, int num_sms) {
  WsPlan best = make_ws_plan_with(d, transposed, num_sms, 0, 0);
  static int spread = -1;
  if (spread < 0) { const char* e = getenv("FFPN_WS_SPREAD"); spread = (e && atoi(e) == 1) ? 1 : 0; }
  if (!best.ok || !spread) return best;
  auto ctas = [&](const WsPlan& w) { const int c = (int)(w.grid.x * w.grid.y); return c < num_sms ? c : num_sms; };
  if (ctas(best) >= num_sms) return best;
  const int npads[3] = {0, 64, 32};                      // 0: the default chunk width
  for (int ni = 0; ni < 3; ni++) {
    if (npads[ni] != 0 && (npads[ni] >= best.p.Npad || best.p.Cout % npads[ni] != 0)) continue;
    for (int nmb = 8; nmb >= 1; nmb >>= 1) {
      if (ni == 0 && nmb == 8) continue;                   // = the default plan
      WsPlan cand = make_ws_plan_with(d, transposed, num_sms, npads[ni], nmb);
      if (!cand.ok) continue;
      if (ctas(cand) > ctas(best)) best = cand;
      if (ctas(best) >= num_sms) return best;
    }
  }
  return best;
}

bool encode_ws_map(CUtensorMap* m, const WsPlan& w, const void* x, bool nan_fill, int in_mult = 1) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return false;
  const WsParams& p = w.p;
  const TcParams& c = w.base.p;
  if (in_mult != 1 && p.tma_mode != 2) return false;                    // the strided row pitch exists for the flat 1x1x1 view only
  const cuuint64_t cb = (cuuint64_t)p.Cin * 2 * (cuuint64_t)in_mult;    // bytes between consecutive input positions
  cuuint64_t dims[4], strides[3];
  cuuint32_t box[4], es[4] = {1, 1, 1, 1};
  dims[0] = (cuuint64_t)p.Cin;
  if (p.tma_mode == 0) {
    dims[1] = p.X; dims[2] = p.Y; dims[3] = p.D;
    strides[0] = cb; strides[1] = (cuuint64_t)(p.Y == 1 ? p.X : c.inY) * cb; strides[2] = (cuuint64_t)c.inD * cb;
  } else if (p.tma_mode == 1) {
    dims[1] = p.X; dims[2] = p.D; dims[3] = p.NB;
    strides[0] = cb; strides[1] = (cuuint64_t)c.inD * cb;
    strides[2] = (cuuint64_t)(p.NB > 1 ? c.inNB : (long long)c.inD * p.D) * cb;
  } else {
    dims[1] = (cuuint64_t)p.X; dims[2] = 1; dims[3] = 1;
    strides[0] = cb; strides[1] = (cuuint64_t)p.X * cb; strides[2] = (cuuint64_t)p.X * cb;
  }
  for (int i = 0; i < 4; i++) box[i] = (cuuint32_t)w.box[i];
  const CUtensorMapSwizzle sw = p.pitch == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : p.pitch == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             nan_fill ? CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA : CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool ws_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("FFPN_WS"); v = (e && atoi(e) == 0) ? 0 : 1; }
  return v != 0;
}

}  // namespace

// Returns 0 = launched, 1 = error (message set), -1 = geometry not handled by this kernel (caller falls back).
// bsc / bsh != nullptr (dgrad only): MODE 2 -- `addend` is then the raw output y of the BatchNorm + ReLU that produced the conv's
// input, and stat_partial receives [rows][2][Cout] = sums of G and G * y (ffpn_bn_bwd_reduce's layout).
static int conv_ws_launch(ffpn_ctx* ctx, const ffpn_conv_desc* d, bool transposed, const void* x, const float* in_scale,
                          const float* in_shift, int in_relu, const float* w, const void* addend, void* y, float* stat_partial,
                          int* stat_rows, void* ws, size_t ws_bytes, cudaStream_t st, const ffpn_bn_fin* fin,
                          const float* bsc = nullptr, const float* bsh = nullptr) {
  if (!ws_enabled()) return -1;
  const bool bnr = bsc != nullptr;
  if (bnr && (!transposed || addend == nullptr || bsh == nullptr || stat_partial == nullptr || fin != nullptr || d->Cin % 16 != 0)) return -1;
  if (bnr) {
    // Measured on B200 (C2 shapes, same box), fused vs dgrad + bn_bwd_reduce: level 1 123 vs 70 + 48 us, level 2 83 vs 47 + 27,
    // level 3 54 vs 31 + 16, level 4 41 vs 31 + 10; training step 9.12 ms with the fusion everywhere, 8.99 ms for maps of up to
    // 40 000 positions only, 8.95 ms without it.  The epilogue is the critical path of this kernel and the fused variant adds a
    // dependent global load (y) per 32-row block to it -- one or two kilobytes in flight per warp cannot cover the DRAM latency
    // (an L2 prefetch one tile ahead recovers 18 of the 71 us at level 1), while the separate reduce kernel streams at 85 % of
    // HBM peak and, in the step, hides behind the weight gradients of the side streams.  So the fused path is OFF unless
    // FFPN_WS_BNR_MAXPOS (largest map, in positions, that takes it) says otherwise; the entry point then launches the two kernels.
    const char* e = getenv("FFPN_WS_BNR_MAXPOS");                        // read per call: tests toggle it
    const long long maxpos = e ? atoll(e) : 0;
    if ((long long)d->B * d->S * d->W * d->H > maxpos) return -1;
  }
  ffpn_conv_desc dp;
  int in_mult = 1;
  const bool strided111 = !transposed && ffpn_make_strided111_desc(d, &dp, &in_mult);   // strided 1x1x1 shortcut -> flat conv over a strided view
  const bool pair = !strided111 && ffpn_make_pair_desc(d, &dp);          // depth-strided projection conv -> stride-1 on the pair view
  int pair2_on = -1;                                                      // read per call: tests toggle it
  // Measured on B200 (same box): level-1 16->16 forward 82.3 us plain vs 84.6 us on the pair view, level-2 32->32 54.1 vs 90.2 us, step 9.88
  // vs 10.13 ms -- the MMA / shared-memory work drops as predicted but the epilogue (N = 2 Cout columns per row, statistics by
  // shuffles) becomes the bound.  Correct (parity suite green with it on) but off unless FFPN_WS_PAIR2=1.
  if (pair2_on < 0) { const char* e = getenv("FFPN_WS_PAIR2"); pair2_on = (e && atoi(e) == 1) ? 1 : 0; }
  bool pair2 = !pair && !strided111 && pair2_on && fin == nullptr && !bnr && ffpn_make_pair2_desc(d, &dp);   // narrow 3-tap stride-1 conv -> pair view of input and output
  WsPlan pl = make_ws_plan((pair || pair2 || strided111) ? &dp : d, transposed, ctx->num_sms);
  if (pair2 && (!pl.ok || pl.nchunks != 1 || pl.p.kgu != 1 || pl.p.upt != 1 || pl.p.kX != 3)) {   // one resident K-group expected
    pair2 = false;
    pl = make_ws_plan(d, transposed, ctx->num_sms);
  }
  if (!pl.ok) return -1;
  if (bnr && pair && pl.nchunks != 1) return -1;                       // the two halves of a pair would land in different CTAs
  if (bnr && !pair && pl.p.Cout != d->Cin) return -1;                  // dgrad of an X-strided conv: N = stride x Cin interleaved positions
  WsParams& p = pl.p;
  const size_t need = (size_t)pl.nchunks * p.b_total_bytes;
  if (ws == nullptr || ws_bytes < need) return -1;
  CUtensorMap tmap;
  if (!encode_ws_map(&tmap, pl, x, in_scale != nullptr, in_mult)) return -1;
  const void* wimg;
  {
    TcParams q = pl.base.p;                                             // geometry + packmode of the shared packer
    q.Npad = p.Npad;
    if (pair) q.packmode = transposed ? 4 : 3;
    if (pair2) q.packmode = transposed ? 6 : 5;
    bool packed_now = false;
    wimg = ffpn_tc_pack_weights(ctx, w, ws, strided111 ? &dp : d, q, pl.nchunks, p.Kc, st, &packed_now);   // d: the ORIGINAL descriptor (weight layout)
    if (packed_now) FFPN_CHECK_LAUNCH(ctx, "pack_weights");
  }
  p.sc = in_scale; p.sh = in_shift; p.wp = (const bf16*)wimg; p.addend = (const bf16*)addend; p.y = (bf16*)y; p.stat = stat_partial;
  p.bsc = bsc; p.bsh = bsh; p.zero = 0u;
  p.bn_mod = (bnr && pair) ? d->Cin : 0;
  p.fold = pair2 ? (transposed ? d->Cin : d->Cout) : (bnr && pair) ? d->Cin : 0;
  p.dbg = ffpn_debug_env("FFPN_TC_DEBUG");
  p.fin_on = 0;
  if (fin != nullptr) {
    if (stat_partial == nullptr) return -1;
    if (ctx->d_counter == nullptr) {
      if (cudaMalloc(&ctx->d_counter, FFPN_FIN_SLOTS * sizeof(unsigned)) != cudaSuccess ||
          cudaMemset(ctx->d_counter, 0, FFPN_FIN_SLOTS * sizeof(unsigned)) != cudaSuccess)
        FFPN_FAIL(ctx, "conv_ws: cannot allocate the arrival counters");
    }
    // one counter word per launch in flight: convs running concurrently on branch / wgrad side streams must not share one
    p.fin_on = 1; p.fin = *fin; p.fin_counter = ctx->d_counter + ctx->fin_slot;
    ctx->fin_slot = (ctx->fin_slot + 1) % FFPN_FIN_SLOTS;
  }
  p.aff_mod = ((pair || pair2) && !transposed) ? d->Cin : 0;
  { const char* e = getenv("FFPN_PDL_EARLY"); p.pdl_early = (e && atoi(e) == 0) ? 0 : 1; }
  p.pair2 = pair2 ? (transposed ? d->Cin : d->Cout) : 0;
  p.relu = in_relu; p.has_aff = in_scale != nullptr; p.has_stats = stat_partial != nullptr; p.has_add = addend != nullptr && !bnr;
  if (!(ctx->attr_mask & FFPN_ATTR_WS)) {                              // per device: the attribute belongs to the function ON the current device
    const void* fns[] = {(const void*)conv_ws_kernel<0, 0, false>, (const void*)conv_ws_kernel<0, 1, false>, (const void*)conv_ws_kernel<1, 0, false>,
                         (const void*)conv_ws_kernel<2, 0, false>, (const void*)conv_ws_kernel<2, 0, true>,  (const void*)conv_ws_kernel<0, 2, false>,
                         (const void*)conv_ws_kernel<1, 2, false>, (const void*)conv_ws_kernel<2, 2, false>, (const void*)conv_ws_kernel<2, 2, true>};
    for (const void* fn : fns) {
      const cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) FFPN_FAIL(ctx, "conv_ws: cannot raise dynamic smem: %s", cudaGetErrorString(e));
    }
    ctx->attr_mask |= FFPN_ATTR_WS;
  }
  {
    static int verbose = -1;
    if (verbose < 0) { const char* e = getenv("FFPN_WS_VERBOSE"); verbose = e ? atoi(e) : 0; }
    if (verbose)
      fprintf(stderr, "conv_ws %s: Cin %d Cout %d Npad %d taps %dx%dx%d mode %d tD %d tY %d L %d Lr %d rows %d Kc %d kgu %d upt %d resident %d "
              "stages %d (issuers %d) stage_bytes %d smem %zu tmem %d x%d grid (%u,%u) tiles %d\n", transposed ? "dgrad" : "fwd", p.Cin, p.Cout, p.Npad, p.kD,
              p.kY, p.kX, p.tma_mode, p.tD, p.tY, p.L, p.Lr, p.region_rows, p.Kc, p.kgu, p.upt, p.w_resident, p.nstages, p.niss, p.stage_bytes,
              pl.smem, p.tmem_cols, p.nbuf, pl.grid.x, pl.grid.y, p.NB * p.nD * p.nI);
  }
  static int trace_mode = -1;
  if (trace_mode < 0) trace_mode = ffpn_debug_env("FFPN_WS_TRACE");
  p.trace = nullptr;
  if (trace_mode) {
    cudaMalloc(&p.trace, 64 * 16 * sizeof(long long));
    cudaMemset(p.trace, 0, 64 * 16 * sizeof(long long));
  }
  // statistics in registers: N = 16 / 32 per warp; N = 64 with the two warps of a quadrant splitting the columns (32 each)
  const int nreg = p.has_add ? 0 : (p.has_stats && p.Npad == 16) ? 1 : (p.has_stats && (p.Npad == 32 || p.Npad == 64)) ? 2 : 0;
  if (p.has_add) ffpn_launch(conv_ws_kernel<0, 1, false>, pl.grid, WS_THREADS, pl.smem, st, p, tmap);
  else if (bnr && nreg == 1) ffpn_launch(conv_ws_kernel<1, 2, false>, pl.grid, WS_THREADS, pl.smem, st, p, tmap);
  else if (bnr && nreg == 2 && p.Npad == 64) ffpn_launch(conv_ws_kernel<2, 2, true>, pl.grid, WS_THREADS, pl.smem, st, p, tmap);
  else if (bnr && nreg == 2) ffpn_launch(conv_ws_kernel<2, 2, false>, pl.grid, WS_THREADS, pl.smem, st, p, tmap);
  else if (bnr) ffpn_launch(conv_ws_kernel<0, 2, false>, pl.grid, WS_THREADS, pl.smem, st, p, tmap);
  else if (nreg == 1) ffpn_launch(conv_ws_kernel<1, 0, false>, pl.grid, WS_THREADS, pl.smem, st, p, tmap);
  else if (nreg == 2 && p.Npad == 64) ffpn_launch(conv_ws_kernel<2, 0, true>, pl.grid, WS_THREADS, pl.smem, st, p, tmap);
  else if (nreg == 2) ffpn_launch(conv_ws_kernel<2, 0, false>, pl.grid, WS_THREADS, pl.smem, st, p, tmap);
  else ffpn_launch(conv_ws_kernel<0, 0, false>, pl.grid, WS_THREADS, pl.smem, st, p, tmap);
  FFPN_CHECK_LAUNCH(ctx, transposed ? "conv_dgrad_ws" : "conv_fwd_ws");
  if (trace_mode) {
    static long long h[64 * 16];
    cudaDeviceSynchronize();
    cudaMemcpy(h, p.trace, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(p.trace);
    const long long t0 = h[2];
    fprintf(stderr, "tile: prod[emptyOK issued] mma[top temptyOK fullOK issued committed] epi[top tfullOK done] xf[fullOK done]  (cycles since MMA top of tile 0)\n");
    for (int t = 0; t < 64 && h[t * 16 + 2]; t++) {
      fprintf(stderr, "%2d:", t);
      for (int k = 0; k < 12; k++) fprintf(stderr, " %7lld", h[t * 16 + k] ? h[t * 16 + k] - t0 : -1);
      fprintf(stderr, "\n");
    }
    trace_mode = 0;
  }
  if (stat_rows) *stat_rows = (int)pl.grid.x;
  return 0;
}

// Plan introspection (host only, no GPU needed): which kernel family takes this conv and how it is tiled.
bool ffpn_wgrad_ws_supported(const ffpn_conv_desc* d);
extern "C" int ffpn_conv_plan_info(const ffpn_conv_desc* d, int transposed, char* buf, size_t n) {
  if (!d || !buf || n == 0) return 1;
  if (transposed == 2) {                                    // weight gradient
    snprintf(buf, n, ffpn_wgrad_ws_supported(d) ? "conv_wgrad_ws" : "not on the warp-specialised kernel (weight gradient)");
    return 0;
  }
  ffpn_conv_desc dp;
  int mult = 1;
  const bool s111 = !transposed && ffpn_make_strided111_desc(d, &dp, &mult);
  const bool pair = !s111 && ffpn_make_pair_desc(d, &dp);
  const Plan base = ffpn_tc_make_plan((s111 || pair) ? &dp : d, transposed != 0, 148);
  const WsPlan pl = make_ws_plan((s111 || pair) ? &dp : d, transposed != 0, 148);
  if (!pl.ok) {
    snprintf(buf, n, "not on the warp-specialised kernel (base geometry %s, view %s)", base.ok ? "ok" : "unsupported",
             s111 ? "strided-1x1x1" : pair ? "pair" : "none");
    return 0;
  }
  const WsParams& p = pl.p;
  snprintf(buf, n, "conv_ws %s view %s: Cin %d Cout %d Npad %d x%d taps %dx%dx%d mode %d tD %d tY %d L %d Lr %d Kc %d kgu %d upt %d resident %d "
           "stages %d issuers %d smem %zu tmem %d x%d grid (%u,%u) tiles %d", transposed ? "dgrad" : "fwd", s111 ? "strided-1x1x1" : pair ? "pair" : "none",
           p.Cin, p.Cout, p.Npad, pl.nchunks, p.kD, p.kY, p.kX, p.tma_mode, p.tD, p.tY, p.L, p.Lr, p.Kc, p.kgu, p.upt, p.w_resident, p.nstages, p.niss,
           pl.smem, p.tmem_cols, p.nbuf, pl.grid.x, pl.grid.y, p.NB * p.nD * p.nI);
  return 0;
}

// Would conv_ws_launch take this geometry (host-side planning only, no launch)?
bool ffpn_conv_ws_supported(const ffpn_conv_desc* d, bool transposed, bool has_aff, bool relu) {
  if (d->dtype != FFPN_BF16 || !ws_enabled()) return false;
  if (has_aff && !relu) return false;
  ffpn_conv_desc dp;
  int mult = 1;
  const bool view = (!transposed && ffpn_make_strided111_desc(d, &dp, &mult)) || ffpn_make_pair_desc(d, &dp);
  return make_ws_plan(view ? &dp : d, transposed, 148).ok;
}

int ffpn_conv_fwd_ws(ffpn_ctx* ctx, const ffpn_conv_desc* d, bool transposed, const void* x, const float* in_scale,
                     const float* in_shift, int in_relu, const float* w, const void* addend, void* y, float* stat_partial,
                     int* stat_rows, void* ws, size_t ws_bytes, cudaStream_t st) {
  return conv_ws_launch(ctx, d, transposed, x, in_scale, in_shift, in_relu, w, addend, y, stat_partial, stat_rows, ws, ws_bytes, st, nullptr);
}

// dgrad fused with pass 1 of the BatchNorm + ReLU backward of the conv's input.  -1: geometry not handled here.
int ffpn_conv_dgrad_ws_bnr(ffpn_ctx* ctx, const ffpn_conv_desc* d, const void* dy, const float* w, const void* y_prev, const float* bn_scale,
                           const float* bn_shift, void* dx, float* partial, int* rows, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int r = conv_ws_launch(ctx, d, true, dy, nullptr, nullptr, 0, w, y_prev, dx, partial, rows, ws, ws_bytes, st, nullptr, bn_scale, bn_shift);
  if (r >= 0) ctx->routes[FFPN_ROUTE_WS]++;
  return r;
}

// Forward conv with the BatchNorm finalize of its output fused in (training mode).  -1: geometry not handled here.
int ffpn_conv_fwd_ws_bn(ffpn_ctx* ctx, const ffpn_conv_desc* d, const void* x, const float* in_scale, const float* in_shift, int in_relu,
                        const float* w, void* y, float* stat_partial, int* stat_rows, void* ws, size_t ws_bytes, cudaStream_t st,
                        double count, float momentum, float eps, const float* gamma, const float* beta, float* rmean, float* rvar,
                        float* scale, float* shift, float* smean, float* sinvstd) {
  // Measured on B200 (C2 step, same box A/B): 13.3 ms with the last-CTA finalize vs 12.9 ms with the separate 3-block finalize
  // kernel -- the serial tail in one CTA costs more than the launch it saves.  Off unless FFPN_FUSED_FIN=1.
  { const char* e = getenv("FFPN_FUSED_FIN"); if (!(e && atoi(e))) return -1; }      // read per call: tests toggle it
  ffpn_bn_fin fin;
  fin.count = count; fin.momentum = momentum; fin.eps = eps; fin.gamma = gamma; fin.beta = beta; fin.running_mean = rmean;
  fin.running_var = rvar; fin.scale = scale; fin.shift = shift; fin.save_mean = smean; fin.save_invstd = sinvstd;
  const int r = conv_ws_launch(ctx, d, false, x, in_scale, in_shift, in_relu, w, nullptr, y, stat_partial, stat_rows, ws, ws_bytes, st, &fin);
  if (r >= 0) ctx->routes[FFPN_ROUTE_WS]++;
  return r;
}
