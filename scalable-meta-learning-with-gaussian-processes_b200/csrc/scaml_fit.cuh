// Fused per-task GP fit kernel: kernel-matrix assembly -> blocked Cholesky -> triangular
// inverse -> (K^-1 contraction with dK/dtheta) for M tasks x R hyper-parameter rows.
//
// One persistent 128-thread CTA owns one evaluation at a time; three CTAs share an SM
// (<= 75 KB shared memory) so that the latency-bound steps of one evaluation (the pivot
// chain of the diagonal tiles, exp-heavy epilogues) overlap with the tile products of the
// other two.  The n_pad x n_pad lower triangle lives in a per-CTA global workspace of
// dense 32x32 fp64 tiles (8 KB each, mostly L2 resident).  Every n^3-class step is the
// same micro-kernel: a warp owns a 32x32 output tile of the 64x64 super-tile and issues
// FP64 tensor-core MMAs (mma.sync m8n8k4 -> SASS DMMA),
//        C(8x8) += A[kk][r] (8x4) * B[kk][c] (4x8)
// on operands staged by cp.async (double-buffered 16-deep half tiles) into shared memory
// with a padded row stride of 36 doubles, which makes the DMMA fragment loads
// bank-conflict free.  Plain DFMA register tiles are bound by shared-memory return
// bandwidth (128 B/clk/SM) at any affordable tile size -- measured 38 % FP64-pipe
// utilisation (profiles/r1_v1_*) -- because every lane re-reads broadcast operands;
// DMMA fragments are distributed across the warp and need 4x fewer bytes per MAC.
// Tile layouts are chosen per phase so that the contraction index is always the slow
// index of the staged tile (no transposes):
//   L   (Cholesky factor, off-diagonal super-tiles)  column-major tiles ("C")
//   L^-1                                              row-major tiles    ("R")
//   D^-1 (inverse of a 64x64 diagonal super-tile)     R in the workspace, C in shared memory
// K and K^-1 never reach memory: K is recomputed from the length-scaled inputs in the
// epilogues, K^-1 super-tiles are contracted with dK/dtheta straight out of the
// accumulator registers.
//
// Math: SURVEY.md appendix A.3-A.5; reference call sites scamlgp/utils.py:171-177,190-192
// (objective), scamlgp/model.py:25-70 (constraints/priors), model.py:176-188 (task loop).
#pragma once
#include "scaml_device.cuh"

namespace scaml {

#ifdef SCAML_PROF
__shared__ long long profsm[16];  // per-CTA phase cycle counters (diagnostics build only)
#define PROF_MARK(ph)                   \
  do {                                  \
    if (threadIdx.x == 0) {             \
      const long long now_ = clock64(); \
      profsm[ph] += now_ - prof_last;   \
      prof_last = now_;                 \
    }                                   \
  } while (0)
#else
#define PROF_MARK(ph) \
  do {                \
  } while (0)
#endif

enum { kModeLmlGrad = 0, kModeFactorize = 1 };
#ifdef SCAML_ABLATE
// timing-only ablation build (WRONG results by design): which phase is on the critical path?
__device__ int g_ablate = 0;
#define ABL(bit) ((g_ablate & (bit)) != 0)
#else
#define ABL(bit) false
#endif
constexpr int kFitThreads = 128;
constexpr int kFitWarps = kFitThreads / 32;
constexpr int kLd = 36;              // shared-memory row stride of a staged tile (doubles)
constexpr int kTileS = kBS * kLd;    // padded tile in shared memory (1152 doubles)
constexpr int kHalfS = 16 * kLd;     // padded 16-deep half tile (576 doubles)
constexpr int kHalfG = kTile / 2;    // dense half tile in global memory (512 doubles)
constexpr int kStage = 8 * kHalfS;   // 2 stages x 4 half tiles == 4 full padded tiles (4608)

struct FitParams {
  const double* X;          // [M][n_max][d]
  const double* y;          // [M][n_max]
  const int32_t* n_valid;   // [M] or null
  const double* theta_raw;  // [M][R][P]
  const double* jitter;     // [M][R] or null
  const int32_t* skip;      // [M][R] or null
  double* lml;              // [M][R]
  double* grad;             // [M][R][P]
  int32_t* info;            // [M][R]
  double* linv_out;         // factorize: [M][ntiles][1024] C-layout
  double* alpha_out;        // factorize: [M][n_pad]
  double* theta_out;        // factorize: [M][P] constrained
  double* workspace;
  long long ws_stride;  // doubles per CTA slot
  long long* prof;      // SCAML_PROF builds only: [grid][16] cycle counters per phase
  int M, R, n_max, n_pad, d, mode;
  // dynamic scheduling: sched[0] = next-item counter (zeroed before the launch), sched[1] = number of items when
  // `order` is given; order = active evaluations sorted by descending cost (scaml_fit_schedule_kernel) or null
  // (items are the evaluations 0 .. M R - 1 themselves)
  int32_t* sched;
  const int32_t* order;
  int ladder;  // 1: psd_safe_cholesky jitter ladder (+1e-8, +1e-7, +1e-6) inside the kernel when a pivot fails
  int sms;  // SM count (co-resident CTAs are blockIdx.x, blockIdx.x + sms, ...)
  int kcache;  // 4-warp RBF kernel: cache kappa between the assembly and the gradient epilogue
  scaml_hyper_spec spec;
};

// workspace slot: lower tiles, tri(NB) x 1024 doubles
inline long long fit_tile_doubles_host(int n_pad) {
  const int NB = n_pad / kBS;
  return (long long)((NB * (NB + 1)) / 2) * kTile;
}
// + kappa cache of the 4-warp kernel: one 64x64 block per lower super-tile (written by the assembly epilogue,
// read back by the gradient epilogue instead of recomputing distances and exponentials; the Matern kernels
// cache kd = -2 dkappa/dr^2 as well, in the second half of the block)
inline long long fit_ws_doubles_host(int n_pad, int d) {
  (void)d;
  const int NS = n_pad / kSB;
  return fit_tile_doubles_host(n_pad) + (long long)((NS * (NS + 1)) / 2) * kSB * kSB * 2;  // kappa and kd
}
// shared memory (doubles): stage 4608 | dinvc 3456 | z,alpha 2*n_pad | red 128 |
//                          gsm 4*kMaxP | par 4*kMaxP+8 | flags 2
// (y is read from global memory where the forward substitution needs it, 64 values per super-row: keeping a copy here
//  cost 4 KB at n = 512 and with it the third co-resident CTA for every n_pad in 384 .. 512)
// -DSCAML_FIT_PROBE4: TIMING-ONLY probe build (WRONG results by design, experiments only -- scripts/fit_probe4.sh): what
// would a fourth co-resident CTA buy?  The D^-1 tiles alias the staging area (44 KB of shared memory per CTA instead of
// 72 KB), the kernel is compiled for 128 registers (__launch_bounds__(128, 4)) and pivot failures are ignored, so the
// instruction stream and the memory traffic of an evaluation stay what they are while four CTAs fit an SM.
#ifdef SCAML_FIT_PROBE4
constexpr int kFitDinvTiles = 0, kFitMinCtas = 4;
#else
constexpr int kFitDinvTiles = 3, kFitMinCtas = 3;
#endif
inline size_t fit_smem_bytes(int n_pad, int d) {
  (void)d;
  return sizeof(double) * (size_t)(kStage + kFitDinvTiles * kTileS + 2 * (size_t)n_pad + 128 + kFitWarps * kMaxP +
                                   5 * kMaxP + 8 + 2);
}

// per-thread coordinates: warp (rb, cb) owns one 32x32 tile of the 64x64 super-tile as a 4x4
// grid of 8x8 DMMA tiles; lane (g, t4) owns element rows 8i+g, cols 8j+2*t4+{0,1}.
struct FThr {
  int tid, warp, lane;
  int role;    // which 32x32 tile of the 64x64 super-tile this warp owns for the current evaluation
  int rb, cb;  // = role >> 1, role & 1 (0/1) -- warp-uniform
  int g, t4;   // lane >> 2, lane & 3
};
SCAML_DEVICE FThr make_fthr() {
  FThr t;
  t.tid = threadIdx.x;
  t.warp = t.tid >> 5;
  t.lane = t.tid & 31;
  t.role = t.warp;
  t.rb = t.role >> 1;
  t.cb = t.role & 1;
  t.g = t.lane >> 2;
  t.t4 = t.lane & 3;
  return t;
}

typedef double Acc[4][4][2];

SCAML_DEVICE void facc_zero(Acc& acc) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
}

// acc(32x32) += A[kk][r] * B[kk][c] over NK4 steps of 4 kk.  Ap/Bp: padded k-major tiles (row
// stride kLd) at their first kk row.  LOWER: skip the strictly-upper 8x8 tiles (diagonal tiles).
// LOWER is a template parameter behind a warp-uniform branch, not a predicate on the DMMAs: a predicated-off
// DMMA still holds the issuing warp for its 16 issue cycles (measured on the fused cross-covariance of the
// prediction kernel), so predication saves pipe time but no warp time.
template <int NK4, bool LOWER>
SCAML_DEVICE void fmma_impl(Acc& acc, const double* __restrict__ Ap, const double* __restrict__ Bp, const FThr& t) {
  const double* ar = Ap + t.t4 * kLd + t.g;
  const double* br = Bp + t.t4 * kLd + t.g;
#pragma unroll 2
  for (int s = 0; s < NK4; ++s) {
    const double a[4] = {ar[0], ar[8], ar[16], ar[24]};
    const double b[4] = {br[0], br[8], br[16], br[24]};
    ar += 4 * kLd;
    br += 4 * kLd;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j <= i || !LOWER) dmma884(acc[i][j], a[i], b[j]);
  }
}
template <int NK4>
SCAML_DEVICE void fmma(Acc& acc, const double* __restrict__ Ap, const double* __restrict__ Bp, const FThr& t,
                       bool lower) {
#ifdef SCAML_FIT_PREDICATED_LOWER  // A/B: the round-1 form (one loop, predicated DMMAs)
  const double* ar = Ap + t.t4 * kLd + t.g;
  const double* br = Bp + t.t4 * kLd + t.g;
#pragma unroll 2
  for (int s = 0; s < NK4; ++s) {
    const double a[4] = {ar[0], ar[8], ar[16], ar[24]};
    const double b[4] = {br[0], br[8], br[16], br[24]};
    ar += 4 * kLd;
    br += 4 * kLd;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j <= i || !lower) dmma884(acc[i][j], a[i], b[j]);
  }
#else
  if (lower) fmma_impl<NK4, true>(acc, Ap, Bp, t);
  else fmma_impl<NK4, false>(acc, Ap, Bp, t);
#endif
}

// The same product when an operand tile is triangular: the 8-row blocks that are identically zero in a k4-step are
// not multiplied.  kbase = index of the first kk row of this call inside its 32 x 32 tile.  Modes:
//   1  operand[kk][x] != 0 only for x <= kk  (rows of L^-1, R-layout)      -> blocks  <= (kk0 + 3) / 8
//   2  operand[kk][x] != 0 only for x >= kk  (D^-1 of a diagonal block, C-layout) -> blocks >= kk0 / 8
// Everything that selects a DMMA is a template parameter and the k loop is fully unrolled, so the skipped products
// do not exist in the instruction stream (round 2 measured the same skipping with run-time predicates as a LOSS:
// a predicated-off DMMA costs the warp the same 16 issue cycles as an executed one).
template <int NK4, int AMODE, int BMODE, bool LOWER, int KBASE>
SCAML_DEVICE void fmma_tri_impl(Acc& acc, const double* __restrict__ Ap, const double* __restrict__ Bp,
                                const FThr& t) {
  const double* ar = Ap + t.t4 * kLd + t.g;
  const double* br = Bp + t.t4 * kLd + t.g;
#pragma unroll
  for (int s = 0; s < NK4; ++s) {
    const int kk0 = KBASE + 4 * s;
    const int up = (kk0 + 3) >> 3, lo = kk0 >> 3;
    const int ihi = (AMODE == 1) ? up : 3, ilo = (AMODE == 2) ? lo : 0;
    const int jhi = (BMODE == 1) ? up : 3, jlo = (BMODE == 2) ? lo : 0;
    double a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = (i >= ilo && i <= ihi) ? ar[8 * i + 4 * s * kLd] : 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = (j >= jlo && j <= jhi) ? br[8 * j + 4 * s * kLd] : 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (i >= ilo && i <= ihi && j >= jlo && j <= jhi && (j <= i || !LOWER)) dmma884(acc[i][j], a[i], b[j]);
  }
}
// streamed products: only the (amode, bmode, lower, kbase) combinations that occur are instantiated.  A diagonal K^-1
// super-tile reads the triangular tiles L^-1(2I, 2I) / (2I+1, 2I+1) on both sides in roles (0,0) / (1,1) (which are
// `lower` as well) and on one side in role (1,0); off the diagonal one operand at most is triangular.  A whole
// 16-deep sub-chunk (NK4 = 4) starts at kbase 0 or 16, the halves of a split tile (NK4 = 2, never `lower`, one
// triangular operand) at 0, 8, 16, 24.
template <int KBASE>
SCAML_DEVICE void fmma_tri4_k(Acc& acc, const double* __restrict__ Ap, const double* __restrict__ Bp, const FThr& t,
                              int amode, int bmode) {
  if (amode == 1 && bmode == 1) fmma_tri_impl<4, 1, 1, true, KBASE>(acc, Ap, Bp, t);
  else if (amode == 1) fmma_tri_impl<4, 1, 0, false, KBASE>(acc, Ap, Bp, t);
  else fmma_tri_impl<4, 0, 1, false, KBASE>(acc, Ap, Bp, t);
}
SCAML_DEVICE void fmma_tri4(Acc& acc, const double* __restrict__ Ap, const double* __restrict__ Bp, const FThr& t,
                            int amode, int bmode, int kbase) {
  if (kbase == 0) fmma_tri4_k<0>(acc, Ap, Bp, t, amode, bmode);  // warp-uniform
  else fmma_tri4_k<16>(acc, Ap, Bp, t, amode, bmode);
}
template <int KBASE>
SCAML_DEVICE void fmma_tri2_k(Acc& acc, const double* __restrict__ Ap, const double* __restrict__ Bp, const FThr& t,
                              int amode) {
  if (amode == 1) fmma_tri_impl<2, 1, 0, false, KBASE>(acc, Ap, Bp, t);
  else fmma_tri_impl<2, 0, 1, false, KBASE>(acc, Ap, Bp, t);
}
SCAML_DEVICE void fmma_tri2(Acc& acc, const double* __restrict__ Ap, const double* __restrict__ Bp, const FThr& t,
                            int amode, int kbase) {
  switch (kbase) {  // warp-uniform
    case 0: fmma_tri2_k<0>(acc, Ap, Bp, t, amode); break;
    case 8: fmma_tri2_k<8>(acc, Ap, Bp, t, amode); break;
    case 16: fmma_tri2_k<16>(acc, Ap, Bp, t, amode); break;
    default: fmma_tri2_k<24>(acc, Ap, Bp, t, amode); break;
  }
}

// ---- chunk sources: which global tiles feed chunk `ck` of a super-tile product ------- //
// Scalar fields only: arrays indexed by the (runtime) warp coordinates would be placed in local memory
// and put LDL/STL round trips on the critical path of every pipeline step.
struct ChunkPtrs {
  const double* a0;
  const double* a1;
  const double* b0;
  const double* b1;
  int zoff;  // first global row/col index covered by this chunk (piggy-backed GEMV)
  SCAML_DEVICE bool a_ok(int rb) const { return rb == 0 ? a0 != nullptr : a1 != nullptr; }
  SCAML_DEVICE bool b_ok(int cb) const { return cb == 0 ? b0 != nullptr : b1 != nullptr; }
};
SCAML_DEVICE long long fit_tile_doubles(int n_pad) {
  const int NB = n_pad / kBS;
  return (long long)((NB * (NB + 1)) / 2) * kTile;
}
SCAML_DEVICE const double* wtile(const double* W, int bi, int bj) { return W + (size_t)(tri(bi) + bj) * kTile; }
SCAML_DEVICE double* wtile_w(double* W, int bi, int bj) { return W + (size_t)(tri(bi) + bj) * kTile; }

// Cholesky update of super-tile (I,J): sum_{kb < 2J} L(I,kb) L(J,kb)^T  (C-layout tiles)
struct CholSrc {
  const double* W;
  int I, J;
  SCAML_DEVICE int count() const { return 2 * J; }
  SCAML_DEVICE bool same() const { return I == J; }
  SCAML_DEVICE int a_tri(int, int) const { return 0; }
  SCAML_DEVICE int b_tri(int, int) const { return 0; }
  SCAML_DEVICE ChunkPtrs get(int ck) const {
    ChunkPtrs c;
    c.a0 = wtile(W, 2 * I, ck);
    c.a1 = wtile(W, 2 * I + 1, ck);
    c.b0 = wtile(W, 2 * J, ck);
    c.b1 = wtile(W, 2 * J + 1, ck);
    c.zoff = ck * kBS;
    return c;
  }
};
// Triangular inverse, super-tile (I,J), I>J: S = sum_{K=J}^{I-1} L(I,K) Linv(K,J)
// A = L tiles (C-layout, contraction over their columns); B = Linv tiles (R-layout, rows)
struct TrtriSrc {
  const double* W;
  int I, J;
  SCAML_DEVICE int count() const { return 2 * (I - J); }
  SCAML_DEVICE bool same() const { return false; }
  SCAML_DEVICE int a_tri(int, int) const { return 0; }
  // chunk 0 / 1 read the lower-triangular diagonal tiles Linv(2J, 2J) / Linv(2J+1, 2J+1) as the B operand of cb = 0 / 1
  SCAML_DEVICE int b_tri(int ck, int cb) const { return (ck < 2 && ck == cb) ? 1 : 0; }
  SCAML_DEVICE ChunkPtrs get(int ck) const {
    ChunkPtrs c;
    const int kcol = 2 * J + ck;
    c.a0 = wtile(W, 2 * I, kcol);
    c.a1 = wtile(W, 2 * I + 1, kcol);
    c.b0 = wtile(W, kcol, 2 * J);
    c.b1 = (kcol >= 2 * J + 1) ? wtile(W, kcol, 2 * J + 1) : nullptr;
    c.zoff = kcol * kBS;
    return c;
  }
};
// K^-1 super-tile (I,J), I>=J: sum_{K>=I} Linv(K,I)^T Linv(K,J)  (R-layout, contraction over rows)
struct LauumSrc {
  const double* W;
  int I, J, NS;
  SCAML_DEVICE int count() const { return 2 * (NS - I); }
  SCAML_DEVICE bool same() const { return I == J; }
  // chunk 0 / 1 read the lower-triangular diagonal tiles Linv(2I, 2I) / Linv(2I+1, 2I+1) as the A operand of rb = 0 / 1
  // (and as the B operand of cb = 0 / 1 on diagonal super-tiles)
  SCAML_DEVICE int a_tri(int ck, int rb) const { return (ck < 2 && ck == rb) ? 1 : 0; }
  SCAML_DEVICE int b_tri(int ck, int cb) const { return (I == J && ck < 2 && ck == cb) ? 1 : 0; }
  SCAML_DEVICE ChunkPtrs get(int ck) const {
    ChunkPtrs c;
    const int krow = 2 * I + ck;
    c.a0 = wtile(W, krow, 2 * I);
    c.a1 = (krow >= 2 * I + 1) ? wtile(W, krow, 2 * I + 1) : nullptr;
    c.b0 = wtile(W, krow, 2 * J);
    c.b1 = (krow >= 2 * J + 1) ? wtile(W, krow, 2 * J + 1) : nullptr;
    c.zoff = krow * kBS;
    return c;
  }
};

// dense half tile (16 x 32 doubles, contiguous 4 KB) global -> padded shared rows: 2 x 16 B / thread
SCAML_DEVICE void half_async(double* sdst, const double* gsrc, int tid) {
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int c2 = tid + u * kFitThreads;
    const int row = c2 >> 4, j = c2 & 15;
    cp_async16(sdst + row * kLd + 2 * j, gsrc + row * kBS + 2 * j);
  }
}

// sub-chunk s = 2*ck + h: rows [16h, 16h+16) of the four tiles of chunk ck
template <class Src>
SCAML_DEVICE void stage_issue(const Src& src, int s, double* st, int tid) {
  const ChunkPtrs c = src.get(s >> 1);
  const int off = (s & 1) * kHalfG;
  if (ABL(16)) return;
  if (c.a0) half_async(st, c.a0 + off, tid);
  if (c.a1) half_async(st + kHalfS, c.a1 + off, tid);
  if (!src.same()) {
    if (c.b0) half_async(st + 2 * kHalfS, c.b0 + off, tid);
    if (c.b1) half_async(st + 3 * kHalfS, c.b1 + off, tid);
  }
  cp_async_commit();
}

// acc += sum over chunks; optional piggy-backed GEMV  pig[c] += sum_kk A[kk][c] * zv[zoff+kk]
// (c = tid & 63 over the 64 A columns of the super-tile, kk-half = tid >> 6).
// DIAG: super-tile on the diagonal -> roles (0,0), (1,1) compute the lower 8x8 tiles only (10 of 16); the one full
// tile (1,0) is split along the contraction between its own warp and the warp of the idle role (0,1): 40 / 32 / 32 / 40
// DMMAs per step instead of 40 / 64 / 0 / 40, partial sums merged by `split_merge` in a fixed order.
// History of the two work-trimming variants (profiles/r2_fit_split_tri_ab.txt, r2_fit_predication.txt):
//   * both LOST while the skipped products were predicated-off DMMAs (split -0.8 %, triangular skipping -3.6 %): a
//     predicated-off DMMA holds the issuing warp for the same 16 issue cycles as an executed one, so roles (0,0) / (1,1)
//     still took 64 DMMA slots per step and nothing got shorter.
//   * with the skipping moved into template parameters behind warp-uniform branches (`fmma_impl`) the split pays:
//     +2.1 % on top of +1.6 % from the templated lower-triangle product itself (-DSCAML_FIT_NOSPLIT is the A/B switch).
//   * -DSCAML_FIT_TRI (zero 8-row blocks of triangular operand tiles left out of the instruction stream, `fmma_tri_impl`:
//     8 % fewer DMMAs) still loses 1-3 %: the variants add ~2000 instructions to a 12 k-instruction kernel whose
//     three co-resident CTAs sit in different phases (instruction-fetch stalls) -- kept off.
// On return every thread has passed a __syncthreads after its last read of `stage`.
template <class Src>
SCAML_DEVICE void gemm_global(Acc& acc, const Src& src, double* stage, const FThr& t, bool diag, bool piggy,
                              double& pig, const double* zv) {
  const int n = 2 * src.count();
  if (n <= 0) return;
#ifdef SCAML_FIT_NOSPLIT
  const bool helper = false;
  if (diag && t.rb < t.cb) {  // A/B variant: role (0,1) idles, role (1,0) multiplies the whole tile
    stage_issue(src, 0, stage, t.tid);
    for (int s = 0; s < n; ++s) {
      cp_async_wait<0>();
      __syncthreads();
      if (s + 1 < n) stage_issue(src, s + 1, stage + ((s + 1) & 1) * 4 * kHalfS, t.tid);
      if (piggy) {
        const ChunkPtrs c = src.get(s >> 1);
        const double* As = stage + (s & 1) * 4 * kHalfS;
        const int col = t.tid & 63, q = t.tid >> 6;
        if (c.a_ok(col >> 5)) {
          const double* ap = As + (col >> 5) * kHalfS + (col & 31) + q * 8 * kLd;
          const double* zp = zv + c.zoff + (s & 1) * 16 + q * 8;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) pig = fma(ap[kk * kLd], zp[kk], pig);
        }
      }
    }
    __syncthreads();
    return;
  }
  const int erb = t.rb, ecb = t.cb;
  const bool split = false;
#else
  const bool helper = diag && t.rb < t.cb;  // role (0,1): works on tile (1,0)
  const int erb = helper ? 1 : t.rb, ecb = helper ? 0 : t.cb;
  const bool split = diag && erb == 1 && ecb == 0;
#endif
  const bool lower = diag && (erb == ecb);
  const int koff = helper ? 8 : 0;  // first kk row (inside a 16-deep sub-chunk) of this warp's share of a split tile
  stage_issue(src, 0, stage, t.tid);
  for (int s = 0; s < n; ++s) {
    double* st = stage + (s & 1) * 4 * kHalfS;
#ifdef SCAML_PROF
    const long long pc0 = clock64();
#endif
    // ONE barrier per step: after it sub-chunk s is visible to everyone and everyone has finished reading
    // sub-chunk s-1, whose buffer the prefetch of s+1 may therefore overwrite.  (A third buffer in phase D -- dinvc is
    // idle there -- with the prefetch two steps ahead measured -0.3 %: profiles/r2_fit_predication.txt.)
    cp_async_wait<0>();
    if (!ABL(131072)) __syncthreads();  // (ablation bit 131072: what do the per-step barriers of the streamed products cost?)
    if (s + 1 < n) stage_issue(src, s + 1, stage + ((s + 1) & 1) * 4 * kHalfS, t.tid);
#ifdef SCAML_PROF
    const long long pc1 = clock64();
#endif
    const int ck = s >> 1;
    const ChunkPtrs c = src.get(ck);
    const double* As = st;
    const double* Bs = src.same() ? st : st + 2 * kHalfS;
    const bool bvalid = src.same() ? c.a_ok(ecb) : c.b_ok(ecb);
    if (c.a_ok(erb) && bvalid && !ABL(8)) {
#ifndef SCAML_FIT_TRI
      const int amode = 0, bmode = 0;
#else
      const int amode = src.a_tri(ck, erb), bmode = src.b_tri(ck, ecb);
#endif
      const double* ap = As + erb * kHalfS + koff * kLd;
      const double* bp = Bs + ecb * kHalfS + koff * kLd;
      const int kbase = (s & 1) * 16 + koff;
      if (split) {
        if (amode | bmode) fmma_tri2(acc, ap, bp, t, amode, kbase);
        else fmma<2>(acc, ap, bp, t, false);
      } else {
        if (amode | bmode) fmma_tri4(acc, ap, bp, t, amode, bmode, kbase);
        else fmma<4>(acc, ap, bp, t, lower);
      }
    }
    if (piggy && !ABL(8192)) {
      const int col = t.tid & 63, q = t.tid >> 6;
      if (c.a_ok(col >> 5)) {
        const double* ap = As + (col >> 5) * kHalfS + (col & 31) + q * 8 * kLd;
        const double* zp = zv + c.zoff + (s & 1) * 16 + q * 8;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) pig = fma(ap[kk * kLd], zp[kk], pig);
      }
    }
#ifdef SCAML_PROF_GEMM
    if (threadIdx.x == 0) {
      profsm[14] += pc1 - pc0;          // wait + barrier + issue next
      profsm[15] += clock64() - pc1;    // compute
    }
#endif
  }
  __syncthreads();
}

// partial sums of the split tile (1,0) of a diagonal super-tile: the helper warp (role (0,1)) parks its accumulators in
// `scratch` (thread-private slots, >= 1024 doubles) BEFORE a __syncthreads, the owner adds them AFTER it
SCAML_DEVICE void split_put(double* scratch, const Acc& acc, const FThr& t) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      scratch[(8 * i + 2 * j) * 32 + t.lane] = acc[i][j][0];
      scratch[(8 * i + 2 * j + 1) * 32 + t.lane] = acc[i][j][1];
    }
}
SCAML_DEVICE void split_merge(Acc& acc, const double* scratch, const FThr& t) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[i][j][0] += scratch[(8 * i + 2 * j) * 32 + t.lane];
      acc[i][j][1] += scratch[(8 * i + 2 * j + 1) * 32 + t.lane];
    }
}

// products of shared-memory resident 64x64 operands (2 chunks of full padded tiles); tile addresses are
// computed arithmetically from the warp coordinates (no pointer tables -> no local memory).
//   full 2x2 operand `F` : tile (r, c) at F + (2 r + c) * kTileS   (or (2 c + r) when stored chunk-major)
//   lower-triangular `D`: tiles (0,0), (1,0), (1,1) at D, D + kTileS, D + 2 kTileS; (0,1) is null
// trsm  (phase B): acc = C_in * D^-T ; A chunk ck = C_in tile (rb, ck) at stage + (2 rb + ck) kTileS,
//                  B[kk][c] = D^-1(c, kk): chunk ck, col-tile cb -> D tile (cb, ck), null for cb < ck
SCAML_DEVICE void gemm_smem_trsm(Acc& acc, const double* cin, const double* dinvc, const FThr& t) {
  if (ABL(128)) return;
#pragma unroll
  for (int ck = 0; ck < 2; ++ck) {
    if (t.cb < ck) continue;
    // D^-1 tiles (0,0) and (1,1) are lower triangular: B[kk][c] = D^-1(c, kk) vanishes for c < kk
#ifndef SCAML_FIT_TRI
    fmma<8>(acc, cin + (2 * t.rb + ck) * kTileS, dinvc + (t.cb + ck) * kTileS, t, false);
#else
    if (t.cb + ck != 1) fmma_tri_impl<8, 0, 2, false, 0>(acc, cin + (2 * t.rb + ck) * kTileS, dinvc + (t.cb + ck) * kTileS, t);
    else fmma<8>(acc, cin + (2 * t.rb + ck) * kTileS, dinvc + (t.cb + ck) * kTileS, t, false);
#endif
  }
}
// trtri (phase C): acc = D^-1 * S ; A[kk][r] = D^-1(r, kk): chunk ck, row-tile rb -> D tile (rb, ck), null for
//                  rb < ck ; B chunk ck = S tile (ck, cb) at stage + (2 ck + cb) kTileS
SCAML_DEVICE void gemm_smem_trtri(Acc& acc, const double* dinvc, const double* sst, const FThr& t) {
  if (ABL(128)) return;
#pragma unroll
  for (int ck = 0; ck < 2; ++ck) {
    if (t.rb < ck) continue;
    // A[kk][r] = D^-1(r, kk) vanishes for r < kk on the triangular tiles (0,0), (1,1)
#ifndef SCAML_FIT_TRI
    fmma<8>(acc, dinvc + (t.rb + ck) * kTileS, sst + (2 * ck + t.cb) * kTileS, t, false);
#else
    if (t.rb + ck != 1) fmma_tri_impl<8, 2, 0, false, 0>(acc, dinvc + (t.rb + ck) * kTileS, sst + (2 * ck + t.cb) * kTileS, t);
    else fmma<8>(acc, dinvc + (t.rb + ck) * kTileS, sst + (2 * ck + t.cb) * kTileS, t, false);
#endif
  }
}

// warp's 32x32 accumulator -> one tile with row/col stride `ld`, column-major ("C") or row-major ("R")
SCAML_DEVICE void store_tile_C(double* blk, int ld, const Acc& acc, const FThr& t, double scale) {
  if (ABL(4096) && ld == kBS) return;
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      double* p = blk + (8 * j + 2 * t.t4 + e) * ld + t.g;
#pragma unroll
      for (int i = 0; i < 4; ++i) p[8 * i] = scale * acc[i][j][e];
    }
}
SCAML_DEVICE void store_tile_R(double* blk, int ld, const Acc& acc, const FThr& t, double scale) {
  if (ABL(4096) && ld == kBS) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double* p = blk + (8 * i + t.g) * ld + 2 * t.t4;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<double2*>(p + 8 * j) = make_double2(scale * acc[i][j][0], scale * acc[i][j][1]);
  }
}

// ---- x-block: length-scaled inputs of the 64 row points (super-tile I) and the 64 column points
// (super-tile J) of one epilogue, xblk[k * 128 + p] (p < 64: rows, p >= 64: columns).  Thread `tid`
// owns point p = tid in every dimension; the first kXpre dimensions are fetched into registers
// BEFORE the tile product so that their latency is hidden behind it.
constexpr int kXpre = 8;
SCAML_DEVICE void xpre_load(double (&xp)[kXpre], const double* Xm, int I, int J, int nv, int d, int tid) {
  const int a = (tid < kSB) ? I * kSB + tid : J * kSB + (tid - kSB);
#pragma unroll
  for (int k = 0; k < kXpre; ++k) xp[k] = (k < d && a < nv && !ABL(65536)) ? __ldg(Xm + (size_t)a * d + k) : 0.0;
}
// `th` here holds the RECIPROCAL lengthscales
SCAML_DEVICE void xblk_store(double* xblk, const double (&xp)[kXpre], const double* Xm, const double* th, int I,
                             int J, int nv, int d, int tid) {
#pragma unroll
  for (int k = 0; k < kXpre; ++k)
    if (k < d) xblk[k * 128 + tid] = xp[k] * th[k];
  if (d > kXpre) {
    const int a = (tid < kSB) ? I * kSB + tid : J * kSB + (tid - kSB);
    for (int k = kXpre; k < d; ++k) xblk[k * 128 + tid] = (a < nv) ? __ldg(Xm + (size_t)a * d + k) * th[k] : 0.0;
  }
}

// squared scaled distances between the thread's row points ra, ra + 8 and its 8 column points
SCAML_DEVICE void pair_r2(double (&r2)[16], const double* xblk, int d, int ra, int cb0) {
#pragma unroll
  for (int u = 0; u < 16; ++u) r2[u] = 0.0;
  if (ABL(512)) return;
#pragma unroll 2
  for (int k = 0; k < d; ++k) {
    const double* xr = xblk + k * 128;
    const double xa0 = xr[ra], xa1 = xr[ra + 8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double2 xb = *reinterpret_cast<const double2*>(xr + cb0 + 8 * j);
      const double d00 = xa0 - xb.x, d01 = xa0 - xb.y, d10 = xa1 - xb.x, d11 = xa1 - xb.y;
      r2[2 * j] = fma(d00, d00, r2[2 * j]);
      r2[2 * j + 1] = fma(d01, d01, r2[2 * j + 1]);
      r2[8 + 2 * j] = fma(d10, d10, r2[8 + 2 * j]);
      r2[8 + 2 * j + 1] = fma(d11, d11, r2[8 + 2 * j + 1]);
    }
  }
}

// ---- epilogue 1: acc <- K_y(I,J) - acc, K recomputed from the scaled inputs ----------- //
// kc: this thread's slots of the super-tile's kappa cache ([pass h][8 pairs of values][128 threads] as double2),
// or null (Matern kernels recompute: their gradient needs a second function of r)
template <int KIND>
SCAML_DEVICE void assemble_tile(Acc& acc, int I, int J, const FThr& t, const double* xblk, int d, int nv, double os,
                                double diag_add, double2* kc) {
  const int ra = t.rb * kBS + t.g, cb0 = kSB + t.cb * kBS + 2 * t.t4;
  const int a0 = I * kSB + ra, b0 = J * kSB + t.cb * kBS + 2 * t.t4;
#pragma unroll
  for (int h = 0; h < 2; ++h) {  // two passes of two row-tiles keep r2 at 32 registers
    double r2[16];  // [i][j][e] -> 8 i + 2 j + e
    pair_r2(r2, xblk, d, ra + 16 * h, cb0);
    if (KIND != SCAML_KERNEL_RBF && kc != nullptr) {
      double kdv[16];
      kappa_n<KIND, 16, true>(r2, r2, kdv);
#pragma unroll
      for (int u = 0; u < 8; ++u) st_stream(kc + (16 + h * 8 + u) * kFitThreads, make_double2(kdv[2 * u], kdv[2 * u + 1]));
    } else if (!ABL(2)) {
      kappa_n<KIND, 16, false>(r2, r2, r2);  // 16 independent exponentials, interleaved
    }
    if (kc != nullptr) {
#pragma unroll
      for (int u = 0; u < 8; ++u) st_stream(kc + (h * 8 + u) * kFitThreads, make_double2(r2[2 * u], r2[2 * u + 1]));
    }
#ifdef SCAML_FIT_INTERIOR_ASM
    // super-tiles off the diagonal and fully inside the valid range need no per-element masks (warp-uniform branch; the
    // same multiply and subtract per element, so results are bit-identical to the masked loop).  OFF by default: in this
    // epilogue the shortcut measured 0 % at n = 256 and -1.9 % at n = 512 (profiles/r2_fit_interior_ab.txt), while the
    // same shortcut in the gradient epilogue below pays at every shape.
    if (I != J && (I + 1) * kSB <= nv) {
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const double k = os * r2[8 * i + 2 * j + e];
            acc[2 * h + i][j][e] = k - acc[2 * h + i][j][e];
          }
      continue;
    }
#endif
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int a = a0 + 8 * (2 * h + i), b = b0 + 8 * j + e;
          double k = os * r2[8 * i + 2 * j + e];
          if (a == b) k += diag_add;
          if (a >= nv || b >= nv) k = (a == b) ? 1.0 : 0.0;
          acc[2 * h + i][j][e] = k - acc[2 * h + i][j][e];
        }
  }
}

// ---- epilogue 2: contract the K^-1 super-tile in `acc` with dK/dtheta ----------------- //
// accumulates into gsm[warp][0..d-1] (lengthscales), [d] (outputscale), [d+1] (trace W).
// acc is overwritten by t_ab = wgt * W_ab * kd_ab.
template <int KIND>
SCAML_DEVICE void grad_tile(Acc& acc, int I, int J, const FThr& t, const double* xblk, const double* av, int d,
                            int nv, double* gsm, const double2* kc) {
  const int ra = t.rb * kBS + t.g, cb0 = kSB + t.cb * kBS + 2 * t.t4;
  const int a0 = I * kSB + ra, b0 = J * kSB + t.cb * kBS + 2 * t.t4;
  double accS = 0.0, accT = 0.0;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    double r2[16], kdv[16];  // [i][j][e] -> 8 i + 2 j + e
    if (kc != nullptr) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const double2 v = ld_stream(kc + (h * 8 + u) * kFitThreads);
        r2[2 * u] = v.x, r2[2 * u + 1] = v.y;
        if (KIND != SCAML_KERNEL_RBF) {
          const double2 q = ld_stream(kc + (16 + h * 8 + u) * kFitThreads);
          kdv[2 * u] = q.x, kdv[2 * u + 1] = q.y;
        }
      }
    } else if (KIND == SCAML_KERNEL_RBF) {
      pair_r2(r2, xblk, d, ra + 16 * h, cb0);
      if (!ABL(4)) kappa_n<KIND, 16, false>(r2, r2, r2);  // kd == kappa for the RBF kernel (kdv unused)
    } else {
      pair_r2(r2, xblk, d, ra + 16 * h, cb0);
      kappa_n<KIND, 16, true>(r2, r2, kdv);
    }
#ifndef SCAML_FIT_NOINTERIOR
    // interior super-tile (off the diagonal, fully inside the valid range): every element counts twice, no masks, no
    // trace term -- the same operations per element as the masked loop below, so results are bit-identical; +1.1 % at
    // n = 256, +0.8 % at n = 512 (profiles/r2_fit_interior_ab.txt)
    if (I != J && (I + 1) * kSB <= nv) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const double ava = av[a0 + 8 * (2 * h + i)];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const double kap = r2[8 * i + 2 * j + e];
            const double kd = (KIND == SCAML_KERNEL_RBF) ? kap : kdv[8 * i + 2 * j + e];
            const double Wab = ava * av[b0 + 8 * j + e] - acc[2 * h + i][j][e];
            const double wk = 2.0 * Wab;
            accS = fma(wk, kap, accS);
            acc[2 * h + i][j][e] = wk * kd;
          }
      }
      continue;
    }
#endif
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int a = a0 + 8 * (2 * h + i);
      const double ava = av[a];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int b = b0 + 8 * j + e;
          const double kap = r2[8 * i + 2 * j + e];
          const double kd = (KIND == SCAML_KERNEL_RBF) ? kap : kdv[8 * i + 2 * j + e];
          const bool use = (a >= b) && (a < nv) && (b < nv);
          const double wgt = use ? ((a == b) ? 1.0 : 2.0) : 0.0;
          const double Wab = ava * av[b] - acc[2 * h + i][j][e];
          const double wk = wgt * Wab;
          accS = fma(wk, kap, accS);
          acc[2 * h + i][j][e] = wk * kd;
          if (use && a == b) accT += Wab;
        }
    }
  }
  double* gw = gsm + t.role * kMaxP;  // per ROLE: the reduction order does not depend on the warp rotation
  for (int k = 0; k < (ABL(1024) ? 0 : d); ++k) {
    const double* xr = xblk + k * 128;
    double xa[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) xa[i] = xr[ra + 8 * i];
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double2 xb = *reinterpret_cast<const double2*>(xr + cb0 + 8 * j);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double d0 = xa[i] - xb.x, d1 = xa[i] - xb.y;
        s = fma(acc[i][j][0], d0 * d0, s);
        s = fma(acc[i][j][1], d1 * d1, s);
      }
    }
    s = warp_sum(s);
    if (t.lane == 0) gw[k] += s;
  }
  accS = warp_sum(accS);
  accT = warp_sum(accT);
  if (t.lane == 0) {
    gw[d] += accS;
    gw[d + 1] += accT;
  }
}

// ---- warp-level 32x32 Cholesky + triangular inverse (row r of the tile per lane) ------ //
// Dsm: SPD tile, C-layout padded (lower part valid).  Lc: padded tile scratch (receives L, C-layout).
// Pz: >= 1024-double scratch (XOR-swizzled transpose buffer).  Outputs: XC (C-layout padded inverse,
// shared), XRs (R-layout padded, shared, optional), XRg (R-layout dense, global, optional).
// Returns 0 or the 1-based failing pivot; adds sum_k log(d_kk) (= log det of the tile) to *logdet.
// Both sweeps are ROLLED loops over the pivot with a rotating register window (slot 0 is always the
// current column) so the body stays ~100 instructions (a fully unrolled version was a 50 KB
// instruction stream: 10 % of the v3 stall samples were no_instruction).  The window shrinks every 8
// pivots (32, 24, 16, 8 slots).  Slots that have rotated past the tile edge read a few doubles beyond
// Lc's rows (still inside this CTA's shared memory) and only ever feed other dead slots.
// 1 / sqrt(pivot).  libdevice's rsqrt() is a ~20-instruction dependent chain with slow-path branches, and it sits on the
// critical path of every one of the n pivots of an evaluation.  -DSCAML_FAST_RSQRT: hardware seed (MUFU.RSQ64H, relative
// error ~2^-22) + ONE cubic step  e = 1 - d y^2,  y <- y + y e (1/2 + 3/8 e)  (error (5/16) eps^3 + rounding: <= 2 ulp).
SCAML_DEVICE double rsqrt_pivot(double d) {
#if defined(SCAML_EMU) || !defined(SCAML_FAST_RSQRT)
  return rsqrt(d);
#else
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double t = d * y;
  const double e = fma(-t, y, 1.0);
  return fma(y * e, fma(0.375, e, 0.5), y);
#endif
}
// pivot check shared by the chain: non-positive / non-finite pivot -> remember the first failure, continue with 1
SCAML_DEVICE double checked_pivot(double dkk, int k, int& fail) {
  if (!(dkk > 0.0) || !(dkk < 1e300)) {
    if (fail == 0) fail = k + 1;
    dkk = 1.0;
  }
  return dkk;
}
// dnext / rsnext: the current pivot A(k,k) of the Schur complement and its rsqrt, produced one step ahead so
// that the (software, ~20 dependent instructions) rsqrt of pivot k+1 overlaps the trailing update of step k.
template <int W>
SCAML_DEVICE void chol_steps(double (&a)[kBS], int k0, double* Lc, int lane, int& fail, double& mydiag, double& myrs,
                             double& dnext, double& rsnext) {
#pragma unroll 1
  for (int k = k0; k < k0 + 8; ++k) {
    const double dkk = dnext, rs = rsnext;
    if (lane == k) {
      mydiag = dkk;
      myrs = rs;
    }
    const double lrk = a[0] * rs;  // a[i] = A(lane, k + i)
    Lc[k * kLd + lane] = lrk;
    // next pivot without the shared-memory round trip: lane k+1 needs only its own L(k+1,k)
    double dn = __shfl_sync(0xffffffffu, fma(-lrk, lrk, a[1]), (k + 1) & 31);
    if (k + 1 < kBS) dn = checked_pivot(dn, k + 1, fail);
    dnext = dn;
    rsnext = rsqrt_pivot(dn);
    __syncwarp();
    const double* lk = Lc + k * kLd + k;  // lk[j] = L(k + j, k)
#pragma unroll
    for (int j = 1; j < W; ++j) a[j - 1] = fma(-lrk, lk[j], a[j]);
  }
}
template <int W>
SCAML_DEVICE void inv_steps(double (&b)[kBS], int j0, const double* Lc, double* XC, double* Pz, int lane, double myrs) {
#pragma unroll 1
  for (int j = j0; j > j0 - 8; --j) {
    const double rj = __shfl_sync(0xffffffffu, myrs, j);
    const double b0 = b[0] * rj;  // X(lane, j);  b[i] = w(j - i)
    XC[j * kLd + lane] = b0;
    Pz[lane * kBS + (j ^ lane)] = b0;
    const double* lj = Lc + j * kLd + j;  // lj[-i * kLd] = L(j, j - i)
#pragma unroll
    for (int i = 1; i < W; ++i) b[i - 1] = fma(-lj[-i * kLd], b0, b[i]);
  }
}
// ---- the same factorisation as a two-warp pipeline (-DSCAML_FIT_FOLLOW; measured, NOT the default) ---------- //
// Result on the B200 (profiles/r2_fit_follower_ab.txt): correct (all parity tests green) and 4 % faster with ONE CTA per
// SM, but 1-3 % slower with the three co-resident CTAs the kernel runs with -- the chain step itself gets slower
// (fence + flag store, a second warp on the tile's shared-memory rows) and the follower competes for the FP64 pipe
// that the other CTAs' products need.
// The inverse does not have to wait for the whole factor: with lane = COLUMN c of X = L^-1 and the pending sums
// acc[r'] = delta(r', c) - sum_{k < r} L(r', k) X(k, c) kept in a rotating register window, row r of X is final as soon
// as column r of L is (X(r, c) = acc[r] / L(r, r), then acc[r'] -= L(r', r) X(r, c) for r' > r) -- so a FOLLOWER warp
// computes X one pivot step behind the CHAIN warp and the 32 inverse steps disappear from the critical path of every
// diagonal block (they were 244 of 531 cycles per pivot, profiles/r2_fit_phase_cycles.txt).  Hand-shake: the chain warp
// stores column k of L and 1 / L(k, k) to shared memory and publishes `k + 1` in a shared counter (fence + volatile
// store by lane 0); the follower spins on the counter.  X is written directly in the three layouts the callers need
// (no transpose buffer).
SCAML_DEVICE void chain_publish(volatile int* prog, int v, int lane) {
  if (lane == 0) {
    __threadfence_block();
    *prog = v;
  }
}
SCAML_DEVICE void chain_wait(volatile int* prog, int v, int lane) {
  if (lane == 0) {
    while (*prog < v) {
#ifdef SCAML_EMU
      sched_yield();
#endif
    }
    __threadfence_block();
  }
  __syncwarp();
}
template <int W>
SCAML_DEVICE void chol_steps_pub(double (&a)[kBS], int k0, double* Lc, double* dg, volatile int* prog, int pbase,
                                 int lane, int& fail, double& mydiag, double& dnext, double& rsnext) {
#pragma unroll 1
  for (int k = k0; k < k0 + 8; ++k) {
    const double dkk = dnext, rs = rsnext;
    if (lane == k) {
      mydiag = dkk;
      dg[k] = rs;
    }
    const double lrk = a[0] * rs;  // a[i] = A(lane, k + i)
    Lc[k * kLd + lane] = lrk;
    double dn = __shfl_sync(0xffffffffu, fma(-lrk, lrk, a[1]), (k + 1) & 31);
    if (k + 1 < kBS) dn = checked_pivot(dn, k + 1, fail);
    dnext = dn;
    rsnext = rsqrt_pivot(dn);
    __syncwarp();
    chain_publish(prog, pbase + k + 1, lane);  // column k of L and dg[k] are complete
    const double* lk = Lc + k * kLd + k;  // lk[j] = L(k + j, k)
#pragma unroll
    for (int j = 1; j < W; ++j) a[j - 1] = fma(-lrk, lk[j], a[j]);
  }
}
template <int W>
SCAML_DEVICE void inv_follow_steps(double (&a)[kBS], int r0, const double* Lc, const double* dg, volatile int* prog,
                                   int pbase, double* XC, double* XRs, double* XRg, int lane) {
#pragma unroll 1
  for (int r = r0; r < r0 + 8; ++r) {
    chain_wait(prog, pbase + r + 1, lane);
    const double x = a[0] * dg[r];  // X(r, lane); a[i] = pending sum of row r + i (0 for lane > r)
    XC[lane * kLd + r] = x;
    if (XRs) XRs[r * kLd + lane] = x;
    if (XRg) XRg[r * kBS + lane] = x;
    const double* lk = Lc + r * kLd + r;  // lk[j] = L(r + j, r)
#pragma unroll
    for (int j = 1; j < W; ++j) a[j - 1] = fma(-x, lk[j], a[j]);
  }
}
// chain warp: Cholesky of the tile in Dsm (C-layout padded), L -> Lc, 1 / L(k,k) -> dg[0..31]; returns the failing pivot
SCAML_DEVICE int chol_32_pub(const double* Dsm, double* Lc, double* dg, volatile int* prog, int pbase, double* logdet,
                             int lane) {
#if defined(SCAML_PROF) && !defined(SCAML_PROF_GEMM)
  const long long pc_a = clock64();
#endif
  double a[kBS];
#pragma unroll
  for (int c = 0; c < kBS; ++c) a[c] = Dsm[c * kLd + lane];
  int fail = 0;
  double mydiag = 1.0;
  double dnext = checked_pivot(__shfl_sync(0xffffffffu, a[0], 0), 0, fail);
  double rsnext = rsqrt_pivot(dnext);
  chol_steps_pub<32>(a, 0, Lc, dg, prog, pbase, lane, fail, mydiag, dnext, rsnext);
  chol_steps_pub<24>(a, 8, Lc, dg, prog, pbase, lane, fail, mydiag, dnext, rsnext);
  chol_steps_pub<16>(a, 16, Lc, dg, prog, pbase, lane, fail, mydiag, dnext, rsnext);
  chol_steps_pub<8>(a, 24, Lc, dg, prog, pbase, lane, fail, mydiag, dnext, rsnext);
  double ld = log(mydiag);
  ld = warp_sum(ld);
  if (lane == 0) *logdet += ld;
#if defined(SCAML_PROF) && !defined(SCAML_PROF_GEMM)
  if (lane == 0) profsm[14] += clock64() - pc_a;  // chain warp: load + 32 Cholesky pivot steps + log det
#endif
  return fail;
}
// follower warp: X = L^-1 one step behind the chain warp
SCAML_DEVICE void inv_32_follow(const double* Lc, const double* dg, volatile int* prog, int pbase, double* XC,
                                double* XRs, double* XRg, int lane) {
#if defined(SCAML_PROF) && !defined(SCAML_PROF_GEMM)
  const long long pc_a = clock64();
#endif
  double a[kBS];
#pragma unroll
  for (int i = 0; i < kBS; ++i) a[i] = (i == lane) ? 1.0 : 0.0;
  inv_follow_steps<32>(a, 0, Lc, dg, prog, pbase, XC, XRs, XRg, lane);
  inv_follow_steps<24>(a, 8, Lc, dg, prog, pbase, XC, XRs, XRg, lane);
  inv_follow_steps<16>(a, 16, Lc, dg, prog, pbase, XC, XRs, XRg, lane);
  inv_follow_steps<8>(a, 24, Lc, dg, prog, pbase, XC, XRs, XRg, lane);
  __syncwarp();
#if defined(SCAML_PROF) && !defined(SCAML_PROF_GEMM)
  if (lane == 0) profsm[15] += clock64() - pc_a;  // follower warp: 32 inverse steps incl. waiting for the chain warp
#endif
}

// ---- X = L^-1 of a 32 x 32 lower-triangular factor, blocked 8 x 8 (default; -DSCAML_FIT_SWEEPINV: the sweep) -------- //
// The substitution sweep above is 32 dependent steps of up to 31 broadcast loads + FMAs on ONE warp while the other
// three warps of the CTA wait.  Blocked:  X_aa = L_aa^-1  for the four diagonal blocks at once (lane = 8 a + j solves
// column j of block a: 7 dependent rows), then the sub-diagonals one after the other,
//     X_ab = -X_aa (sum_{c = b}^{a-1} L_ac X_cb),        a - b = 1, 2, 3  (3, 2, 1 independent blocks each),
// with every 8 x 8 x 8 product as two DMMAs: 32 DMMAs and ~40 dependent steps instead of 32 x 31 FMAs in 32 steps.
// The sum S leaves the accumulator layout through a 64-double scratch per block to become the B operand of X_aa S.
// Lc: L (C-layout padded).  myrs: lane k holds 1 / L(k, k).  Outputs as for the sweep: XC (C-layout padded), XRs
// (R-layout padded, optional), XRg (R-layout dense global, optional), zeros above the diagonal.  Pz: >= 192 doubles.
SCAML_DEVICE void inv_32_blocked(const double* Lc, double* Pz, double* XC, double* XRs, double* XRg, int lane,
                                 double myrs) {
  const int a = lane >> 3, j = lane & 7;
  const int g = lane >> 2, t4 = lane & 3;
  // diagonal blocks: column j of X_aa by forward substitution (rows above j stay 0)
  {
    double x[8];
    const double* La = Lc + (8 * a) * kLd + 8 * a;  // La[k * kLd + r] = L_aa(r, k)
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const double rsr = __shfl_sync(0xffffffffu, myrs, 8 * a + r);
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int k = 0; k < r; ++k) {
        if (k & 1) s1 = fma(La[k * kLd + r], x[k], s1);
        else s0 = fma(La[k * kLd + r], x[k], s0);
      }
      x[r] = (r < j) ? 0.0 : ((r == j) ? rsr : -(s0 + s1) * rsr);
    }
    double* xc = XC + (8 * a + j) * kLd + 8 * a;
#pragma unroll
    for (int r = 0; r < 8; r += 2) *reinterpret_cast<double2*>(xc + r) = make_double2(x[r], x[r + 1]);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (XRs) XRs[(8 * a + r) * kLd + 8 * a + j] = x[r];
      if (XRg) XRg[(8 * a + r) * kBS + 8 * a + j] = x[r];
    }
    // zeros above the block diagonal: block (a, bb), bb > a, column j of it
    for (int bb = a + 1; bb < 4; ++bb) {
      double* zc = XC + (8 * bb + j) * kLd + 8 * a;
#pragma unroll
      for (int r = 0; r < 8; r += 2) *reinterpret_cast<double2*>(zc + r) = make_double2(0.0, 0.0);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (XRs) XRs[(8 * a + r) * kLd + 8 * bb + j] = 0.0;
        if (XRg) XRg[(8 * a + r) * kBS + 8 * bb + j] = 0.0;
      }
    }
  }
  __syncwarp();
  // sub-diagonals: the blocks of one sub-diagonal are independent
#pragma unroll
  for (int dlt = 1; dlt < 4; ++dlt) {
    double R[3][2];
#pragma unroll
    for (int b = 0; b + dlt < 4; ++b) {
      const int aa = b + dlt;
      double S[2] = {0.0, 0.0};
#pragma unroll
      for (int c = b; c < aa; ++c)
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2)  // S += L_ac X_cb : A[row][k] = L(8 aa + row, 8 c + k), B[k][col] = X(8 c + k, 8 b + col)
          dmma884(S, Lc[(8 * c + 4 * s2 + t4) * kLd + 8 * aa + g], XC[(8 * b + g) * kLd + 8 * c + 4 * s2 + t4]);
      *reinterpret_cast<double2*>(Pz + 64 * b + 8 * g + 2 * t4) = make_double2(S[0], S[1]);  // S(g, 2 t4 + e)
    }
    __syncwarp();
#pragma unroll
    for (int b = 0; b + dlt < 4; ++b) {
      const int aa = b + dlt;
      R[b][0] = R[b][1] = 0.0;
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2)  // R = X_aa S : A[row][k] = X(8 aa + row, 8 aa + k), B[k][col] = S(k, col)
        dmma884(R[b], XC[(8 * aa + 4 * s2 + t4) * kLd + 8 * aa + g], Pz[64 * b + 8 * (4 * s2 + t4) + g]);
    }
    __syncwarp();  // S consumed before the next sub-diagonal overwrites the scratch
#pragma unroll
    for (int b = 0; b + dlt < 4; ++b) {
      const int aa = b + dlt;
      const double v0 = -R[b][0], v1 = -R[b][1];  // X_ab(g, 2 t4 + e)
      XC[(8 * b + 2 * t4) * kLd + 8 * aa + g] = v0;
      XC[(8 * b + 2 * t4 + 1) * kLd + 8 * aa + g] = v1;
      if (XRs) *reinterpret_cast<double2*>(XRs + (8 * aa + g) * kLd + 8 * b + 2 * t4) = make_double2(v0, v1);
      if (XRg) *reinterpret_cast<double2*>(XRg + (8 * aa + g) * kBS + 8 * b + 2 * t4) = make_double2(v0, v1);
    }
    __syncwarp();  // X of this sub-diagonal visible to the next one
  }
}

SCAML_DEVICE int chol_inv_32(const double* Dsm, double* Lc, double* Pz, double* XC, double* XRs, double* XRg,
                             double* logdet, int lane) {
#if defined(SCAML_PROF) && !defined(SCAML_PROF_GEMM)
  const long long pc_a = clock64();
#endif
  double a[kBS];
#pragma unroll
  for (int c = 0; c < kBS; ++c) a[c] = Dsm[c * kLd + lane];
  int fail = 0;
  double mydiag = 1.0, myrs = 1.0;
  double dnext = checked_pivot(__shfl_sync(0xffffffffu, a[0], 0), 0, fail);
  double rsnext = rsqrt_pivot(dnext);
  if (!ABL(32)) {
    chol_steps<32>(a, 0, Lc, lane, fail, mydiag, myrs, dnext, rsnext);
    chol_steps<24>(a, 8, Lc, lane, fail, mydiag, myrs, dnext, rsnext);
    chol_steps<16>(a, 16, Lc, lane, fail, mydiag, myrs, dnext, rsnext);
    chol_steps<8>(a, 24, Lc, lane, fail, mydiag, myrs, dnext, rsnext);
  }
#if defined(SCAML_PROF) && !defined(SCAML_PROF_GEMM)
  const long long pc_b = clock64();
  if (lane == 0) profsm[14] += pc_b - pc_a;  // chain warp: load + 32 Cholesky pivot steps
#endif
  double ld = log(mydiag);
  ld = warp_sum(ld);
  if (lane == 0) *logdet += ld;
#ifndef SCAML_FIT_SWEEPINV
  __syncwarp();
  inv_32_blocked(Lc, Pz, XC, XRs, XRg, lane, myrs);
#else
  // inverse, row `lane` of X = L^-1 by a backward column sweep on L^T
  double b[kBS];
#pragma unroll
  for (int i = 0; i < kBS; ++i) b[i] = (i == kBS - 1 - lane) ? 1.0 : 0.0;
  if (!ABL(1)) {
    inv_steps<32>(b, 31, Lc, XC, Pz, lane, myrs);
    inv_steps<24>(b, 23, Lc, XC, Pz, lane, myrs);
    inv_steps<16>(b, 15, Lc, XC, Pz, lane, myrs);
    inv_steps<8>(b, 7, Lc, XC, Pz, lane, myrs);
  }
  __syncwarp();
  if (XRs != nullptr || XRg != nullptr) {
#pragma unroll 4
    for (int r = 0; r < kBS; ++r) {
      const double v = Pz[r * kBS + (lane ^ r)];
      if (XRs) XRs[r * kLd + lane] = v;
      if (XRg) XRg[r * kBS + lane] = v;
    }
  }
  __syncwarp();
#endif
#if defined(SCAML_PROF) && !defined(SCAML_PROF_GEMM)
  if (lane == 0) profsm[15] += clock64() - pc_b;  // chain warp: log det + 32 inverse steps + transposed outputs
#endif
  return fail;
}

// 32x32x32 product by the whole CTA on padded tiles: out(r,c) = sum_kk A[kk][r] * B[kk][c];
// warp w owns the 16x16 quadrant (w>>1, w&1) as 2x2 DMMA tiles.
typedef double SAcc[2][2][2];
SCAML_DEVICE void small_gemm(SAcc& o, const double* A, const double* B, const FThr& t) {
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) o[i][j][0] = o[i][j][1] = 0.0;
  const double* ar = A + t.t4 * kLd + 16 * (t.warp >> 1) + t.g;
  const double* br = B + t.t4 * kLd + 16 * (t.warp & 1) + t.g;
  if (ABL(256)) return;
#pragma unroll 4
  for (int s = 0; s < 8; ++s) {
    const double a0 = ar[0], a1 = ar[8], b0 = br[0], b1 = br[8];
    ar += 4 * kLd;
    br += 4 * kLd;
    dmma884(o[0][0], a0, b0);
    dmma884(o[0][1], a0, b1);
    dmma884(o[1][0], a1, b0);
    dmma884(o[1][1], a1, b1);
  }
}
// element (i,j,e) of the small accumulator <-> row 16*(w>>1) + 8i + g, col 16*(w&1) + 8j + 2*t4 + e
SCAML_DEVICE void small_store_C(double* blk, int ld, const SAcc& o, const FThr& t, double scale) {
  const int r0 = 16 * (t.warp >> 1) + t.g, c0 = 16 * (t.warp & 1) + 2 * t.t4;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) blk[(c0 + 8 * j + e) * ld + r0 + 8 * i] = scale * o[i][j][e];
}
SCAML_DEVICE void small_store_R(double* blk, int ld, const SAcc& o, const FThr& t, double scale) {
  const int r0 = 16 * (t.warp >> 1) + t.g, c0 = 16 * (t.warp & 1) + 2 * t.t4;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
      *reinterpret_cast<double2*>(blk + (r0 + 8 * i) * ld + c0 + 8 * j) =
          make_double2(scale * o[i][j][0], scale * o[i][j][1]);
}

// ---- factorise + invert the 64x64 diagonal super-tile held in `stage` ----------------- //
// stage tiles (C-layout, padded): T0 = D00, T1 = free, T2 = D10, T3 = D11.   dinvc tiles V0..V2.
// Results: V0, V1, V2 = D^-1 tiles (0,0), (1,0), (1,1) in C-layout (shared);
//          R-layout dense tiles of D^-1 written to the workspace diagonal (wd00, wd10, wd11).
// *logdet (shared scalar) accumulates log det; *flag receives the failing pivot (1-based).
#ifdef SCAML_PROF
#define DIAG_PROF_ARGS , long long& prof_last
#define DIAG_PROF_PASS , prof_last
#else
#define DIAG_PROF_ARGS
#define DIAG_PROF_PASS
#endif
SCAML_DEVICE void diag_factor(double* stage, double* dinvc, double* wd00, double* wd10, double* wd11, double* logdet,
                              int* flag, int pivot_base, const FThr& t, int chain_warp, double* dg DIAG_PROF_ARGS) {
  // flag[3]: progress counter of the chain -> follower hand-shake (zeroed by the caller before its last barrier)
  volatile int* prog = flag + 3;
  const int inv_warp = (chain_warp + 1) & (kFitWarps - 1);
  double* T0 = stage;
  double* T1 = stage + kTileS;
  double* T2 = stage + 2 * kTileS;
  double* T3 = stage + 3 * kTileS;
  double* V0 = dinvc;
  double* V1 = dinvc + kTileS;
  double* V2 = dinvc + 2 * kTileS;
#ifdef SCAML_FIT_FOLLOW
  if (t.warp == chain_warp) {
    // L scratch = V1; X00: C-layout -> V0, R-layout -> T1 and workspace (written by the follower warp)
    const int f = chol_32_pub(T0, V1, dg, prog, 0, logdet, t.lane);
    if (f && t.lane == 0 && *flag == 0) *flag = pivot_base + f;
  } else if (t.warp == inv_warp) {
    inv_32_follow(V1, dg, prog, 0, V0, T1, wd00, t.lane);
  }
#else
  (void)prog, (void)inv_warp, (void)dg;
  if (t.warp == chain_warp) {
    // L scratch = V1, transpose scratch = V2, X00: C-layout -> V0, R-layout -> T1 and workspace
    const int f = chol_inv_32(T0, V1, V2, V0, T1, wd00, logdet, t.lane);
    if (f && t.lane == 0 && *flag == 0) *flag = pivot_base + f;
  }
#endif
  PROF_MARK(11);
  __syncthreads();
  SAcc o;
  // L10 = D10 * X00^T      (A = D10 C-layout, B[kk][c] = X00(c,kk) = C-layout X00)
  small_gemm(o, T2, V0, t);
  __syncthreads();
  small_store_C(T2, kLd, o, t, 1.0);  // T2 now holds L10 (C-layout)
  __syncthreads();
  // D11 -= L10 L10^T ;  Tm = L10 * X00  (B[kk][c] = X00(kk,c) = R-layout X00 in T1) -> V1 (R-layout)
  small_gemm(o, T2, T2, t);
  {
    const int r0 = 16 * (t.warp >> 1) + t.g, c0 = 16 * (t.warp & 1) + 2 * t.t4;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) T3[(c0 + 8 * j + e) * kLd + r0 + 8 * i] -= o[i][j][e];
  }
  small_gemm(o, T2, T1, t);
  small_store_R(V1, kLd, o, t, 1.0);
  __syncthreads();
  PROF_MARK(12);
#ifdef SCAML_FIT_FOLLOW
  if (t.warp == chain_warp) {
    // L scratch = T1 (X00 R-layout is dead)
    const int f = chol_32_pub(T3, T1, dg + kBS, prog, kBS, logdet, t.lane);
    if (f && t.lane == 0 && *flag == 0) *flag = pivot_base + kBS + f;
  } else if (t.warp == inv_warp) {
    inv_32_follow(T1, dg + kBS, prog, kBS, V2, nullptr, wd11, t.lane);
  }
#else
  if (t.warp == chain_warp) {
    // L scratch = T1 (X00 R-layout is dead), transpose scratch = T0 (D00 is dead)
    const int f = chol_inv_32(T3, T1, T0, V2, nullptr, wd11, logdet, t.lane);
    if (f && t.lane == 0 && *flag == 0) *flag = pivot_base + kBS + f;
  }
#endif
  PROF_MARK(13);
  __syncthreads();
  // X10 = -X11 * Tm        (A[kk][r] = X11(r,kk) = C-layout X11 in V2, B = Tm R-layout in V1)
  small_gemm(o, V2, V1, t);
  __syncthreads();  // every thread has consumed Tm before V1 is overwritten
  small_store_C(V1, kLd, o, t, -1.0);
  small_store_R(wd10, kBS, o, t, -1.0);
  __syncthreads();
}

// D^-1 of diagonal super-tile I (R-layout dense tiles in the workspace) -> dinvc (C-layout padded)
// Coalesced cp.async into `stage` (which must be idle), then a shared->shared transpose; ends with
// a __syncthreads.
SCAML_DEVICE void load_dinvc(double* dinvc, double* stage, const double* W, int I, int tid) {
  const double* src[3] = {wtile(W, 2 * I, 2 * I), wtile(W, 2 * I + 1, 2 * I), wtile(W, 2 * I + 1, 2 * I + 1)};
#pragma unroll
  for (int b = 0; b < 3; ++b) {
    half_async(stage + b * kTileS, src[b], tid);
    half_async(stage + b * kTileS + kHalfS, src[b] + kHalfG, tid);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  if (!ABL(16384)) {
#pragma unroll
    for (int b = 0; b < 3; ++b) {
#pragma unroll 4
      for (int idx = tid; idx < kTile; idx += kFitThreads) {
        const int r = idx >> 5, c = idx & 31;
        dinvc[b * kTileS + c * kLd + r] = stage[b * kTileS + r * kLd + c];
      }
    }
  }
  __syncthreads();
}

// out[r] = sum_kk D^-1(r,kk) v[kk] over a 64x64 lower-triangular D^-1 held as dinvc (C-layout tiles)
SCAML_DEVICE void dinv_matvec(double* out, const double* dinvc, const double* v, double* red, const FThr& t) {
  const int r = t.tid & 63, q = t.tid >> 6;  // q: 32-wide kk half
  double s = 0.0;
  for (int kk = q * 32; kk < (ABL(32768) ? 0 : q * 32 + 32); ++kk) {
    if (kk > r) break;
    const int rb_ = r >> 5, kb_ = kk >> 5;
    const double* blk = dinvc + ((rb_ == 0) ? 0 : (kb_ == 0 ? kTileS : 2 * kTileS));
    s = fma(blk[(kk & 31) * kLd + (r & 31)], v[kk], s);
  }
  red[q * 64 + r] = s;
  __syncthreads();
  if (t.tid < 64) out[t.tid] = red[t.tid] + red[64 + t.tid];
  __syncthreads();
}

// ---- schedule: active evaluations, most expensive first --------------------------------------------------- //
// One CTA.  Buckets the rows that are not skipped by their super-tile count NS = ceil(n_valid / 64) (cost ~ NS^3) with
// shared-memory counters and writes them to `order` bucket by bucket, largest first; sched[0] = 0 (the work counter of
// the fit kernel), sched[1] = number of active rows.  The order inside a bucket is whatever the atomics give: results
// do not depend on the schedule.  Rows with an invalid n_valid stay in the list (the fit kernel reports info = -1).
constexpr int kSchedThreads = 1024;
constexpr int kSchedBuckets = 64;
template <int UNUSED = 0>  // a template only so that the header can be included in several translation units
__global__ void __launch_bounds__(kSchedThreads) scaml_fit_schedule_kernel(const int32_t* skip, const int32_t* n_valid,
                                                                           int M, int R, int n_max, int32_t* sched,
                                                                           int32_t* order) {
  __shared__ int cnt[kSchedBuckets], base[kSchedBuckets];
  const int tid = threadIdx.x, E = M * R;
  if (tid < kSchedBuckets) cnt[tid] = 0;
  __syncthreads();
  for (int e = tid; e < E; e += kSchedThreads) {
    if (skip != nullptr && skip[e] != 0) continue;
    const int nv = n_valid ? n_valid[e / R] : n_max;
    int b = (nv + kSB - 1) / kSB;
    b = b < 1 ? 1 : (b > kSchedBuckets - 1 ? kSchedBuckets - 1 : b);
    atomicAdd(&cnt[b], 1);
  }
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int b = kSchedBuckets - 1; b >= 0; --b) {
      base[b] = run;
      run += cnt[b];
      cnt[b] = 0;
    }
    sched[0] = 0;
    sched[1] = run;
  }
  __syncthreads();
  for (int e = tid; e < E; e += kSchedThreads) {
    if (skip != nullptr && skip[e] != 0) continue;
    const int nv = n_valid ? n_valid[e / R] : n_max;
    int b = (nv + kSB - 1) / kSB;
    b = b < 1 ? 1 : (b > kSchedBuckets - 1 ? kSchedBuckets - 1 : b);
    order[base[b] + atomicAdd(&cnt[b], 1)] = e;
  }
}

template <int KIND>
__global__ void __launch_bounds__(kFitThreads, kFitMinCtas) scaml_fit_kernel(const FitParams p) {
  SCAML_DYN_SMEM(double, sm);
  FThr t = make_fthr();
  const int d = p.d, P = p.d + 2, n_pad_max = p.n_pad;
  double* stage = sm;             // 2 stages x 4 padded half tiles | 4 full padded tiles (C_in / S / diag)
  double* dinvc = stage + (kFitDinvTiles ? kStage : 0);  // 3 padded tiles (probe build: aliased with `stage`)
  double* zv = stage + kStage + kFitDinvTiles * kTileS;
  double* av = zv + n_pad_max;
  double* red = av + n_pad_max;  // 128
  double* gsm = red + 128;       // kFitWarps * kMaxP
  double* par = gsm + kFitWarps * kMaxP;
  double* th = par;
  double* lp = par + kMaxP;
  double* dlp = par + 2 * kMaxP;
  double* chain = par + 3 * kMaxP;
  double* invl = par + 4 * kMaxP;  // reciprocal lengthscales (x * (1/l): an FP64 division is ~10 pipe slots)
  double* scal = par + 5 * kMaxP;  // [0] logdet
  int* flag = reinterpret_cast<int*>(scal + 8);

#ifdef SCAML_FIT_PROBE_HOTWS
  // TIMING-ONLY probe (WRONG results): the co-resident CTAs of an SM share ONE workspace slot per SM-slot pair modulo 32,
  // so the whole tile / kappa workspace of the grid is 32 slots (~20 MB) and stays L2-resident -- what does the DRAM
  // spill of the 270 MB workspace cost?
  double* W = p.workspace + (size_t)(blockIdx.x & 31) * p.ws_stride;
#else
  double* W = p.workspace + (size_t)blockIdx.x * p.ws_stride;
#endif
  // kappa cache behind the tiles; the mapping thread -> slot depends on the thread id only, so the warp-role
  // rotation (which is per evaluation) does not matter
  double2* kcache = (p.kcache && p.mode == kModeLmlGrad)
                        ? reinterpret_cast<double2*>(W + fit_tile_doubles(p.n_pad))
                        : nullptr;
#ifdef SCAML_PROF
  long long prof_last = clock64();
  if (threadIdx.x < 16) profsm[threadIdx.x] = 0;
  __syncthreads();
#endif
  const int E = p.M * p.R;
  const scaml_hyper_spec& sp = p.spec;
  // The four tile roles of a super-tile carry unequal work (on diagonal super-tiles role (1,0) multiplies a
  // full tile, (0,0)/(1,1) the lower 8x8 blocks only and (0,1) nothing, epilogues included), and warp w of
  // every co-resident CTA sits on sub-partition w % 4.  Rotating the warp -> role map by the CTA's slot on
  // its SM and by the evaluation index spreads the heavy role over the four tensor pipes; the pivot chains run
  // on the warp that holds the lightest role.  Results do not depend on the rotation (role-indexed reductions).
  const int slot = p.sms > 0 ? (int)(blockIdx.x / p.sms) : 0;

  // Work distribution: every CTA pulls the next item from a global counter (one atomic per evaluation), so CTAs
  // that drew short evaluations (small n_i, rows that fail early) or skipped rows simply pull more; with `order`
  // the items are the active rows only, most expensive (largest n_i) first.  Results do not depend on which CTA
  // runs an evaluation (tested bit-identical under permutation), so the schedule is free.
  const int n_items = (p.order != nullptr) ? p.sched[1] : E;
  for (int it = 0;; ++it) {
    __syncthreads();  // previous evaluation fully retired before shared state (and the work slot) is rewritten
    if (t.tid == 0) flag[2] = atomicAdd(p.sched, 1);
    __syncthreads();
    const int item = flag[2];
    if (item >= n_items) break;
    const int e = (p.order != nullptr) ? p.order[item] : item;
    if (p.order == nullptr && p.skip != nullptr && p.skip[e] != 0) continue;
    const int m = e / p.R;
    const int nv = p.n_valid ? p.n_valid[m] : p.n_max;
    if (nv < 1 || nv > p.n_max) {
      if (t.tid == 0) p.info[e] = -1;
      continue;
    }
    const int NS = (nv + kSB - 1) / kSB, n_pad = NS * kSB;
    const int rot = slot + it;
    t.role = (t.warp + rot) & (kFitWarps - 1);
    t.rb = t.role >> 1;
    t.cb = t.role & 1;
    const int chain_warp = (1 - rot) & (kFitWarps - 1);  // the warp whose role is (0,1)

    // ---- parameters: Interval transform, priors, chain rule -------------------------- //
    if (t.tid < P) {
      const double raw = p.theta_raw[(size_t)e * P + t.tid];
      double lo, hi, p1, p2;
      int pk;
      if (t.tid < d) {
        lo = sp.ls_lo, hi = sp.ls_hi, pk = sp.ls_prior, p1 = sp.ls_p1, p2 = sp.ls_p2;
      } else if (t.tid == d) {
        lo = sp.os_lo, hi = sp.os_hi, pk = sp.os_prior, p1 = sp.os_p1, p2 = sp.os_p2;
      } else {
        lo = sp.noise_lo, hi = sp.noise_hi, pk = sp.noise_prior, p1 = sp.noise_p1, p2 = sp.noise_p2;
      }
      const double sg = sigmoid(raw);
      const double v = lo + (hi - lo) * sg;
      th[t.tid] = v;
      lp[t.tid] = log_prior(pk, p1, p2, v);
      dlp[t.tid] = dlog_prior(pk, p1, p2, v);
      chain[t.tid] = (hi - lo) * sg * (1.0 - sg);
      invl[t.tid] = 1.0 / v;
      if (p.mode == kModeFactorize) p.theta_out[(size_t)e * P + t.tid] = v;
    }
    if (t.tid == 0) {
      scal[0] = 0.0;
      *flag = 0;
    }
    for (int i = t.tid; i < kFitWarps * kMaxP; i += kFitThreads) gsm[i] = 0.0;
    __syncthreads();
    const double os = th[d];
    const double* Xm = p.X + (size_t)m * p.n_max * d;
    const double* ym = p.y + (size_t)m * p.n_max;

    Acc acc;
    double pig = 0.0;
    bool failed = false;
    PROF_MARK(0);

    // ================= phase B: blocked left-looking Cholesky ========================== //
    const bool upper_warp = (t.rb == 0 && t.cb == 1);  // idle on diagonal super-tiles
    // psd_safe_cholesky (linear_operator): on a failed pivot the factorisation is repeated with 1e-8,
    // 1e-7, 1e-6 added to the diagonal of the ORIGINAL matrix -- here inside the kernel (p.ladder), so that the host
    // never reads `info` back between the rounds of the L-BFGS driver.
    double jit = p.jitter ? p.jitter[e] : 0.0;
    for (int attempt = 0;; ++attempt) {
    const double diag_add = th[d + 1] + jit;
    failed = false;
    for (int J = 0; J < NS && !failed; ++J) {
      for (int I = J; I < NS; ++I) {
        const bool diag = (I == J);
        facc_zero(acc);
        double xp[kXpre];
        xpre_load(xp, Xm, I, J, nv, d, t.tid);
        CholSrc src{W, I, J};
        gemm_global(acc, src, stage, t, diag, false, pig, nullptr);
        PROF_MARK(1);
        // split tile (1,0) of a diagonal super-tile: dinvc is free here (all trsm of column J-1 have retired)
#ifndef SCAML_FIT_NOSPLIT
        if (diag && J > 0 && upper_warp) split_put(dinvc, acc, t);
#endif
        xblk_store(stage, xp, Xm, invl, I, J, nv, d, t.tid);  // stage is idle: x-block lives there
        __syncthreads();
#ifndef SCAML_FIT_NOSPLIT
        if (diag && J > 0 && t.rb == 1 && t.cb == 0) split_merge(acc, dinvc, t);
#endif
        if (!(diag && upper_warp))
          assemble_tile<KIND>(acc, I, J, t, stage, d, nv, os, diag_add,
                              kcache ? kcache + ((size_t)(tri(I) + J) * 32) * kFitThreads + t.tid : nullptr);
        __syncthreads();  // x-block consumed before C_in overwrites it
        if (!(diag && upper_warp)) store_tile_C(stage + (t.rb * 2 + t.cb) * kTileS, kLd, acc, t, 1.0);
        if (diag && t.tid == 0) flag[3] = 0;  // chain -> follower progress counter of diag_factor
        __syncthreads();
        PROF_MARK(2);
        if (diag) {
          diag_factor(stage, dinvc, wtile_w(W, 2 * J, 2 * J), wtile_w(W, 2 * J + 1, 2 * J),
                      wtile_w(W, 2 * J + 1, 2 * J + 1), &scal[0], flag, J * kSB, t, chain_warp, red DIAG_PROF_PASS);
          PROF_MARK(3);
#if !defined(SCAML_FIT_PROBE4) && !defined(SCAML_FIT_PROBE_HOTWS)
          if (*flag != 0 && !ABL(0x7fffffff)) {
            failed = true;
            break;
          }
#endif
        } else {
          // L(I,J) = C * D^-T : A chunk kb2 = C tiles (rb,kb2); B[kk][c] = D^-1(c,kk)
          facc_zero(acc);
          gemm_smem_trsm(acc, stage, dinvc, t);
          store_tile_C(wtile_w(W, 2 * I + t.rb, 2 * J + t.cb), kBS, acc, t, 1.0);
          __syncthreads();
          PROF_MARK(4);
        }
      }
    }
    if (!failed || !p.ladder || attempt == 3) break;
    jit = (attempt == 0) ? 1e-8 : ((attempt == 1) ? 1e-7 : 1e-6);
    __syncthreads();
    if (t.tid == 0) {
      scal[0] = 0.0;
      *flag = 0;
    }
    __syncthreads();
    }
    if (failed) {
      if (t.tid == 0) {
        p.info[e] = *flag;
        if (p.mode == kModeLmlGrad) p.lml[e] = nan("");
      }
      if (p.mode == kModeLmlGrad && t.tid < P) p.grad[(size_t)e * P + t.tid] = nan("");
      continue;
    }

    // ================= phase C: triangular inverse (row-wise), z = L^-1 y ============== //
    for (int I = 0; I < NS; ++I) {
      __syncthreads();
      load_dinvc(dinvc, stage, W, I, t.tid);
      pig = 0.0;
      for (int J = 0; J < I; ++J) {
        facc_zero(acc);
        TrtriSrc src{W, I, J};
        gemm_global(acc, src, stage, t, false, J == 0, pig, zv);
        PROF_MARK(5);
        store_tile_R(stage + (t.rb * 2 + t.cb) * kTileS, kLd, acc, t, 1.0);  // S, R-layout tiles (kb2, cb)
        __syncthreads();
        facc_zero(acc);
        gemm_smem_trtri(acc, dinvc, stage, t);
        store_tile_R(wtile_w(W, 2 * I + t.rb, 2 * J + t.cb), kBS, acc, t, -1.0);
        if (p.mode == kModeFactorize)
          store_tile_C(p.linv_out + ((size_t)m * tri(n_pad_max / kBS) + tri(2 * I + t.rb) + 2 * J + t.cb) * kTile, kBS,
                       acc, t, -1.0);
        __syncthreads();
        PROF_MARK(6);
      }
      // z_I = D_I^-1 (y_I - sum_{K<I} L(I,K) z_K)
      red[t.tid] = pig;
      __syncthreads();
      if (t.tid < 64) {
        const int iy = I * kSB + t.tid;
        av[t.tid] = ((iy < nv) ? __ldg(ym + iy) : 0.0) - (red[t.tid] + red[64 + t.tid]);  // av: scratch here
      }
      __syncthreads();
      dinv_matvec(zv + I * kSB, dinvc, av, red, t);
      PROF_MARK(7);
      if (p.mode == kModeFactorize) {
        // diagonal tiles of L^-1 in C-layout = dinvc (padded -> dense)
        double* lo = p.linv_out + (size_t)m * tri(n_pad_max / kBS) * kTile;
        for (int i = t.tid; i < kTile; i += kFitThreads) {
          const int si = (i >> 5) * kLd + (i & 31);
          lo[(size_t)(tri(2 * I) + 2 * I) * kTile + i] = dinvc[si];
          lo[(size_t)(tri(2 * I + 1) + 2 * I) * kTile + i] = dinvc[kTileS + si];
          lo[(size_t)(tri(2 * I + 1) + 2 * I + 1) * kTile + i] = dinvc[2 * kTileS + si];
        }
      }
    }
    __syncthreads();

    if (p.mode == kModeFactorize) {
      // alpha = L^-T z : warp w owns 32-columns bj = w, w+4, ... ; lane = column inside the tile
      const int NB = 2 * NS;
      for (int bj = t.warp; bj < NB; bj += kFitWarps) {
        double s = 0.0;
        for (int bi = bj; bi < NB; ++bi) {
          const double* blk = wtile(W, bi, bj);  // R-layout: (r,c) at r*32+c
#pragma unroll 8
          for (int r = 0; r < kBS; ++r) s = fma(__ldcg(blk + r * kBS + t.lane), zv[bi * kBS + r], s);
        }
        p.alpha_out[(size_t)m * n_pad_max + bj * kBS + t.lane] = s;
      }
      for (int i = n_pad + t.tid; i < n_pad_max; i += kFitThreads) p.alpha_out[(size_t)m * n_pad_max + i] = 0.0;
      if (t.tid == 0) p.info[e] = 0;
      continue;
    }

    // ================= phase D: K^-1 super-tiles, fused gradient contraction ========== //
    for (int I = 0; I < NS; ++I) {
      for (int jj = 0; jj <= I; ++jj) {  // diagonal super-tile first: it completes alpha_I
        const bool diag = (jj == 0);
        const int J = diag ? I : jj - 1;
        facc_zero(acc);
        pig = 0.0;
        double xp[kXpre];
        xpre_load(xp, Xm, I, J, nv, d, t.tid);
        LauumSrc src{W, I, J, NS};
        gemm_global(acc, src, stage, t, diag, diag, pig, zv);
        PROF_MARK(8);
#ifndef SCAML_FIT_NOSPLIT
        if (diag && upper_warp) split_put(dinvc, acc, t);  // dinvc is not used in phase D
#endif
        xblk_store(stage, xp, Xm, invl, I, J, nv, d, t.tid);
        if (diag) red[t.tid] = pig;
        __syncthreads();
#ifndef SCAML_FIT_NOSPLIT
        if (diag && t.rb == 1 && t.cb == 0) split_merge(acc, dinvc, t);
#endif
        if (diag) {
          if (t.tid < 64) av[I * kSB + t.tid] = red[t.tid] + red[64 + t.tid];
          __syncthreads();
        }
        if (!(diag && upper_warp))
          grad_tile<KIND>(acc, I, J, t, stage, av, d, nv, gsm,
                          kcache ? kcache + ((size_t)(tri(I) + J) * 32) * kFitThreads + t.tid : nullptr);
        __syncthreads();  // x-block consumed before the next product stages tiles over it
        PROF_MARK(9);
      }
    }
    __syncthreads();
    // quad = z^T z (fixed order), final scalars
    {
      double q = 0.0;
      for (int i = t.tid; i < n_pad; i += kFitThreads) q = fma(zv[i], zv[i], q);
      q = warp_sum(q);
      if (t.lane == 0) red[t.warp] = q;
      __syncthreads();
      if (t.tid < P) {
        double g = 0.0;
        for (int w = 0; w < kFitWarps; ++w) g += gsm[w * kMaxP + t.tid];
        const double gt = (t.tid < d) ? 0.5 * os * g / th[t.tid] : 0.5 * g;
        p.grad[(size_t)e * P + t.tid] = (gt + dlp[t.tid]) * chain[t.tid] / (double)nv;
      }
      if (t.tid == 0) {
        double quad = 0.0;
        for (int w = 0; w < kFitWarps; ++w) quad += red[w];
        double prior = 0.0;
        for (int k = 0; k < P; ++k) prior += lp[k];
        p.lml[e] = (-0.5 * (quad + scal[0] + (double)nv * kLog2Pi) + prior) / (double)nv;
        p.info[e] = 0;
      }
    }
    PROF_MARK(10);
  }
#ifdef SCAML_PROF
  __syncthreads();
  if (p.prof != nullptr && threadIdx.x < 16) p.prof[(size_t)blockIdx.x * 16 + threadIdx.x] = profsm[threadIdx.x];
#endif
}

}  // namespace scaml
