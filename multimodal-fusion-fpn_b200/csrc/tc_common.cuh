// Shared pieces of the tcgen05 convolution kernels (conv_tc.cu, conv_ws.cu): PTX wrappers, the canonical conv
// geometry (TcParams / Plan) and the weight packer.
#pragma once
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline EncodeTiledFn get_encode_tiled() {
  // resolved through the runtime so that the library has no link-time dependency on libcuda (it must load on CPU boxes)
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}


constexpr int TC_THREADS = 256;
constexpr int SCR_STRIDE = 17;                       // per-warp 32x16 transpose scratch, padded

struct TcParams {
  // canonical geometry (stride-1 convolution)
  int NB, D, Y, X, oD, oY, oX;
  int kD, kY, kX, pD, pY, pX;
  long long inNB, inD, inY, outNB, outD, outY;       // position strides (X stride is 1)
  int Cin, Cout, Npad;
  int Xp, tD, L, Lr, nD, nI, Qout;
  int sX, nsets, hl;                                 // X stride, residue sets staged separately, left halo slots
  int packmode;                                      // 0 fwd, 1 dgrad (stride 1), 2 dgrad of an X-strided conv
  int use_tma, tma_mode, tY;                         // A tiles staged by TMA: 0 lines (C,X,Y,D) 1 slices (C,Xflat,D,NB) 2 flat (C,256,P/256)
  unsigned tma_bytes;                                // bytes of one unit's TMA loads (expect_tx)
  int dbg;                                           // FFPN_TC_DEBUG bitmask (timing experiments only): 1 no MMA, 2 no epilogue, 4 no staging
  int KG, nkg, colstride, tmem_cols;
  int rows_alloc, region_rows;
  int relu, has_aff, has_stats, has_add;
  unsigned a_bytes, b_bytes;                         // per K-group smem bytes
  const bf16* x;
  const float* sc;
  const float* sh;
  const bf16* wp;                                    // packed weights
  const bf16* addend;
  bf16* y;
  float* stat;                                       // [gridDim.x][2][Cout]
};

static __device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

static __device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
static __device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
static __device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
static __device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
static __device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // UMMA shared-memory descriptor (cute/arch/mma_sm100_desc.hpp): start[0,14) lbo[16,30) sbo[32,46) version=1 at
  // [46,48), base_offset 0, layout_type[61,64) = SWIZZLE_NONE (0).  All in 16-byte units.
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
static __device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
static __device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
static __device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

constexpr int N_ISSUE = 4;          // warps whose lane 0 issues tcgen05.mma (different accumulator blocks each)
constexpr int HDR_TAPS = 32;        // byte offset of the tap-offset table in the smem header
constexpr int HDR_STATS = 160;      // byte offset of the per-CTA statistics accumulators

static __device__ __forceinline__ uint4 bn_relu_bf16x8(uint4 v, const float (&s)[8], const float (&h)[8], int relu) {
  uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int q = 0; q < 4; q++) {
    float f0 = fmaf(__uint_as_float(u[q] << 16), s[2 * q], h[2 * q]);
    float f1 = fmaf(__uint_as_float(u[q] & 0xffff0000u), s[2 * q + 1], h[2 * q + 1]);
    if (relu) { f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f); }
    __nv_bfloat162 hh = __floats2bfloat162_rn(f0, f1);
    u[q] = *reinterpret_cast<uint32_t*>(&hh);
  }
  return make_uint4(u[0], u[1], u[2], u[3]);
}

static __device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;                       // src-size 0 => the 16 bytes are zero-filled (padding)
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}

static __device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tmap, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

static __device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* tmap, int c0, int c1, int c2, int c3, int c4, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

struct Plan {
  TcParams p;
  size_t smem;
  int grid;
  int nchunks;
  bool simple;     // single-buffer kernel, 4 CTAs per SM
  bool ok;
};

// The projection's depth-strided conv (1,1,3) stride (1,1,2) pad (0,0,1) is a stride-1 conv on the PAIR VIEW of its input:
// x[.., H, C] reinterpreted (for free, it is contiguous) as x'[.., H/2, 2C] (even position | odd position), two taps
// (pair o-1, pair o) and remapped weights W'[co][(h,ci)][t'] = W[co][ci][dx], (t',h): (0,1)->dx 0, (1,0)->1, (1,1)->2,
// (0,0)->zero.  Forward, dgrad and wgrad of the strided conv then run on the stride-1 warp-specialised kernels.
static inline bool ffpn_make_pair_desc(const ffpn_conv_desc* d, ffpn_conv_desc* dp) {
  if (d->dtype != FFPN_BF16 || d->kS != 1 || d->kW != 1 || d->kH != 3 || d->sS != 1 || d->sW != 1 || d->sH != 2 || d->pS != 0 ||
      d->pW != 0 || d->pH != 1 || (d->H & 1) || d->oH != d->H / 2)
    return false;
  *dp = *d;
  dp->H = d->H / 2; dp->Cin = 2 * d->Cin; dp->kH = 2; dp->sH = 1; dp->pH = 1;
  return true;
}

// The projection's shortcut, a 1x1x1 conv with stride (1,1,s) (fusion3D2D.py:317-326), reads every s-th depth position:
// x[.., H, C] seen as x'[.., H/s, s*C] with only the first C channels of each row used -- a plain 1x1x1 conv over H/s
// positions whose input rows are s*C elements apart.  The TMA tensor map expresses the row pitch, so the warp-specialised
// flat (1x1x1) path runs it unchanged; *mult is the pitch multiplier for the input map.
static inline bool ffpn_make_strided111_desc(const ffpn_conv_desc* d, ffpn_conv_desc* ds, int* mult) {
  if (d->dtype != FFPN_BF16 || d->kS != 1 || d->kW != 1 || d->kH != 1 || d->sS != 1 || d->sW != 1 || d->sH <= 1 || d->pS != 0 ||
      d->pW != 0 || d->pH != 0 || d->H % d->sH != 0 || d->oH != d->H / d->sH)
    return false;
  *ds = *d;
  ds->H = d->H / d->sH; ds->sH = 1;
  *mult = d->sH;
  return true;
}

// Narrow stride-1 convs with three taps along the contiguous axis, (kS=1, kW, 3) pad (.,.,1), are shared-memory bound on the
// tensor core's reads of the tap views.  On the pair view of input AND output -- x'[.., H/2, 2Cin],
// y'[.., H/2, 2Cout] -- the conv is again a 3-tap stride-1 conv, W'[(ho,co)][(hi,ci)][t'] = W[co][ci][dx] with
// dx = 2(t'-1) + hi - ho + 1 (zero when outside 0..2), and the outer pair taps have an all-zero K half (t'=0 only reads the odd
// element, t'=2 only the even one), which the MMA loop skips: 4 instead of 6 K=16 tap reads per two positions.
static inline bool ffpn_make_pair2_desc(const ffpn_conv_desc* d, ffpn_conv_desc* dp) {
  if (d->dtype != FFPN_BF16 || d->kS != 1 || d->kH != 3 || d->sS != 1 || d->sW != 1 || d->sH != 1 || d->pS != 0 || d->pH != 1 ||
      (d->H & 1) || d->oH != d->H || d->H < 4 || !(d->Cin == 16 || d->Cin == 32) || !(d->Cout == 16 || d->Cout == 32))
    return false;
  *dp = *d;
  dp->H = d->H / 2; dp->oH = d->oH / 2; dp->Cin = 2 * d->Cin; dp->Cout = 2 * d->Cout;
  return true;
}

// BatchNorm finalize fused into the producing conv kernel (done by the last CTA to finish): ffpn_conv_fwd_bn
struct ffpn_bn_fin {
  double count;
  float momentum, eps;
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
  float* scale;
  float* shift;
  float* save_mean;
  float* save_invstd;
};

// canonical geometry + tiling of the staged (no-swizzle) kernels; conv_ws.cu re-tiles on top of the geometry
Plan ffpn_tc_make_plan(const ffpn_conv_desc* d, bool transposed, int num_sms);
// fp32 master weights -> bf16 smem image [nchunk][kg][tap][kc][n (Npad)][8] (mode: 0 fwd, 1 dgrad, 2 strided dgrad, 3/4 pair view).
// Returns the device pointer of the image: `ws` after a per-call packing launch, or the arena slot when the packed-weight
// arena is replaying (no launch).  *launched tells the caller whether a kernel was enqueued.
const void* ffpn_tc_pack_weights(ffpn_ctx* ctx, const float* w, void* ws, const ffpn_conv_desc* d, const TcParams& p, int nchunks,
                                 int KG, cudaStream_t st, bool* launched);
