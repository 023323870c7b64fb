// Fused per-task GP fit kernel, 8-warp variant (the default): same algorithm, phases, tile layouts and
// workspace as scaml_fit.cuh (read its header first), but a 256-thread CTA in which every warp owns a
// 32 x 16 slice of the 64 x 64 super-tile (8 DMMA accumulators instead of 16).  Two such CTAs share an SM:
// 16 resident warps instead of 12, two warps of the same evaluation per tensor pipe in the tile products,
// and twice the threads in the exp-heavy epilogues -- the 4-warp kernel left the shared FP64/DMMA pipe 50 %
// idle because each of its phases is latency bound at one warp per sub-partition (profiles/r1_*).
// Numerically the two variants perform the same operations in the same order per output element
// (identical accumulation order over kk; gradient partials are reduced per tile role in a fixed order).
#pragma once
#include "scaml_fit.cuh"

namespace scaml {
namespace f8 {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

// shared memory (doubles): stage 4608 | dinvc 3456 | y,z,alpha 3*n_pad | red 256 | gsm 8*kMaxP | par 5*kMaxP+8 | flags 2
inline size_t smem_bytes(int n_pad, int d) {
  (void)d;
  return sizeof(double) * (size_t)(kStage + 3 * kTileS + 3 * (size_t)n_pad + 256 + kWarps * kMaxP + 5 * kMaxP + 8 + 2);
}

// warp role r in 0..7: rb = r >> 2 (tile row), cq = r & 3 (16-column group of the 64 columns)
struct Thr {
  int tid, warp, lane;
  int role;
  int rb, cb;  // tile row / col inside the super-tile (0/1) -- warp-uniform
  int cin;     // first column inside the tile (0 or 16)
  int g, t4;   // lane >> 2, lane & 3
};
SCAML_DEVICE void set_role(Thr& t, int role) {
  t.role = role;
  t.rb = role >> 2;
  t.cb = (role >> 1) & 1;
  t.cin = (role & 1) * 16;
}
SCAML_DEVICE Thr make_thr() {
  Thr t;
  t.tid = threadIdx.x;
  t.warp = t.tid >> 5;
  t.lane = t.tid & 31;
  t.g = t.lane >> 2;
  t.t4 = t.lane & 3;
  set_role(t, t.warp);
  return t;
}

// acc[i][j][e]: row 8 i + g, column cin + 8 j + 2 t4 + e of the warp's tile (i < 4, j < 2)
typedef double Acc8[4][2][2];

SCAML_DEVICE void acc_zero(Acc8& acc) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
}

// acc(32 x 16) += A[kk][r] * B[kk][cin + c] over NK4 steps of 4 kk.  Ap/Bp: padded k-major tiles (row stride
// kLd) at their first kk row.  LOWER: skip the strictly-upper 8x8 blocks (diagonal tiles).
// LOWER and the warp's first column block are template parameters behind warp-uniform branches: a predicated-off
// DMMA still holds the issuing warp for its 16 issue cycles (see fmma in scaml_fit.cuh).
template <int NK4, bool LOWER, int JJ0>
SCAML_DEVICE void mma8_impl(Acc8& acc, const double* __restrict__ Ap, const double* __restrict__ Bp, const Thr& t) {
  const double* ar = Ap + t.t4 * kLd + t.g;
  const double* br = Bp + t.t4 * kLd + t.cin + t.g;
#pragma unroll 2
  for (int s = 0; s < NK4; ++s) {
    const double a[4] = {ar[0], ar[8], ar[16], ar[24]};
    const double b[2] = {br[0], br[8]};
    ar += 4 * kLd;
    br += 4 * kLd;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
        if (!LOWER || JJ0 + j <= i) dmma884(acc[i][j], a[i], b[j]);
  }
}
template <int NK4>
SCAML_DEVICE void mma8(Acc8& acc, const double* __restrict__ Ap, const double* __restrict__ Bp, const Thr& t,
                       bool lower) {
  if (!lower) mma8_impl<NK4, false, 0>(acc, Ap, Bp, t);
  else if ((t.cin >> 3) == 0) mma8_impl<NK4, true, 0>(acc, Ap, Bp, t);  // first 8-column block of this warp: 0 or 2
  else mma8_impl<NK4, true, 2>(acc, Ap, Bp, t);
}

// dense half tile (16 x 32 doubles, contiguous 4 KB) global -> padded shared rows: 16 B / thread
SCAML_DEVICE void half_async8(double* sdst, const double* gsrc, int tid) {
  const int row = tid >> 4, j = tid & 15;
  cp_async16(sdst + row * kLd + 2 * j, gsrc + row * kBS + 2 * j);
}

// sub-chunk s = 2*ck + h: rows [16h, 16h+16) of the four tiles of chunk ck
template <class Src>
SCAML_DEVICE void stage_issue8(const Src& src, int s, double* st, int tid) {
  const ChunkPtrs c = src.get(s >> 1);
  const int off = (s & 1) * kHalfG;
  if (c.a0) half_async8(st, c.a0 + off, tid);
  if (c.a1) half_async8(st + kHalfS, c.a1 + off, tid);
  if (!src.same()) {
    if (c.b0) half_async8(st + 2 * kHalfS, c.b0 + off, tid);
    if (c.b1) half_async8(st + 3 * kHalfS, c.b1 + off, tid);
  }
  cp_async_commit();
}

// acc += sum over chunks; optional piggy-backed GEMV  pig[c] += sum_kk A[kk][c] * zv[zoff+kk]
// (c = tid & 63 over the 64 A columns of the super-tile, kk quarter = tid >> 6).
// DIAG: super-tile on the diagonal -> the warps of tile (0,1) idle, tiles (0,0),(1,1) compute lower blocks only.
// On return every thread has passed a __syncthreads after its last read of `stage`.
template <class Src>
SCAML_DEVICE void gemm_global8(Acc8& acc, const Src& src, double* stage, const Thr& t, bool diag, bool piggy,
                               double& pig, const double* zv) {
  const int n = 2 * src.count();
  if (n <= 0) return;
  const bool active = !(diag && t.rb < t.cb);
  const bool lower = diag && (t.rb == t.cb);
  stage_issue8(src, 0, stage, t.tid);
  for (int s = 0; s < n; ++s) {
    double* st = stage + (s & 1) * 4 * kHalfS;
    // ONE barrier per step: after it sub-chunk s is visible to everyone and everyone has finished reading
    // sub-chunk s-1, whose buffer the prefetch of s+1 may therefore overwrite.
    cp_async_wait<0>();
    __syncthreads();
    if (s + 1 < n) stage_issue8(src, s + 1, stage + ((s + 1) & 1) * 4 * kHalfS, t.tid);
    const ChunkPtrs c = src.get(s >> 1);
    const double* As = st;
    const double* Bs = src.same() ? st : st + 2 * kHalfS;
    const bool bvalid = src.same() ? c.a_ok(t.cb) : c.b_ok(t.cb);
    if (active && c.a_ok(t.rb) && bvalid) mma8<4>(acc, As + t.rb * kHalfS, Bs + t.cb * kHalfS, t, lower);
    if (piggy) {
      const int col = t.tid & 63, q = t.tid >> 6;
      if (c.a_ok(col >> 5)) {
        const double* ap = As + (col >> 5) * kHalfS + (col & 31) + q * 4 * kLd;
        const double* zp = zv + c.zoff + (s & 1) * 16 + q * 4;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) pig = fma(ap[kk * kLd], zp[kk], pig);
      }
    }
  }
  __syncthreads();
}

// products of shared-memory resident 64x64 operands (see gemm_smem_trsm / gemm_smem_trtri in scaml_fit.cuh)
SCAML_DEVICE void smem_trsm8(Acc8& acc, const double* cin, const double* dinvc, const Thr& t) {
#pragma unroll
  for (int ck = 0; ck < 2; ++ck) {
    if (t.cb < ck) continue;
    mma8<8>(acc, cin + (2 * t.rb + ck) * kTileS, dinvc + (t.cb + ck) * kTileS, t, false);
  }
}
SCAML_DEVICE void smem_trtri8(Acc8& acc, const double* dinvc, const double* sst, const Thr& t) {
#pragma unroll
  for (int ck = 0; ck < 2; ++ck) {
    if (t.rb < ck) continue;
    mma8<8>(acc, dinvc + (t.rb + ck) * kTileS, sst + (2 * ck + t.cb) * kTileS, t, false);
  }
}

// warp's 32x16 accumulator -> its tile (base `blk`, row/col stride `ld`), column-major ("C") or row-major ("R")
SCAML_DEVICE void store8_C(double* blk, int ld, const Acc8& acc, const Thr& t, double scale) {
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      double* p = blk + (t.cin + 8 * j + 2 * t.t4 + e) * ld + t.g;
#pragma unroll
      for (int i = 0; i < 4; ++i) p[8 * i] = scale * acc[i][j][e];
    }
}
SCAML_DEVICE void store8_R(double* blk, int ld, const Acc8& acc, const Thr& t, double scale) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double* p = blk + (8 * i + t.g) * ld + t.cin + 2 * t.t4;
#pragma unroll
    for (int j = 0; j < 2; ++j)
      *reinterpret_cast<double2*>(p + 8 * j) = make_double2(scale * acc[i][j][0], scale * acc[i][j][1]);
  }
}

// ---- x-block: xblk[k * 128 + p] (p < 64: row points of super-tile I, p >= 64: column points of J).
// Two threads per point: thread tid owns point tid & 127 in the dimensions k = kh, kh + 2, ... (kh = tid >> 7);
// the first kXh of them are fetched into registers BEFORE the tile product.
constexpr int kXh = 4;
SCAML_DEVICE void xpre_load8(double (&xp)[kXh], const double* Xm, int I, int J, int nv, int d, int tid) {
  const int pt = tid & 127, kh = tid >> 7;
  const int a = (pt < kSB) ? I * kSB + pt : J * kSB + (pt - kSB);
#pragma unroll
  for (int u = 0; u < kXh; ++u) {
    const int k = 2 * u + kh;
    xp[u] = (k < d && a < nv) ? __ldg(Xm + (size_t)a * d + k) : 0.0;
  }
}
// `th` here holds the RECIPROCAL lengthscales
SCAML_DEVICE void xblk_store8(double* xblk, const double (&xp)[kXh], const double* Xm, const double* th, int I, int J,
                              int nv, int d, int tid) {
  const int pt = tid & 127, kh = tid >> 7;
#pragma unroll
  for (int u = 0; u < kXh; ++u) {
    const int k = 2 * u + kh;
    if (k < d) xblk[k * 128 + pt] = xp[u] * th[k];
  }
  if (d > 2 * kXh) {
    const int a = (pt < kSB) ? I * kSB + pt : J * kSB + (pt - kSB);
    for (int k = 2 * kXh + kh; k < d; k += 2) xblk[k * 128 + pt] = (a < nv) ? __ldg(Xm + (size_t)a * d + k) * th[k] : 0.0;
  }
}

// squared scaled distances between the thread's 4 row points ra + 8 i and its 4 column points: r2[4 i + 2 j + e]
SCAML_DEVICE void pair_r2_8(double (&r2)[16], const double* xblk, int d, int ra, int cb0) {
#pragma unroll
  for (int u = 0; u < 16; ++u) r2[u] = 0.0;
#pragma unroll 2
  for (int k = 0; k < d; ++k) {
    const double* xr = xblk + k * 128;
    const double2 xb0 = *reinterpret_cast<const double2*>(xr + cb0);
    const double2 xb1 = *reinterpret_cast<const double2*>(xr + cb0 + 8);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double xa = xr[ra + 8 * i];
      const double d0 = xa - xb0.x, d1 = xa - xb0.y, d2 = xa - xb1.x, d3 = xa - xb1.y;
      r2[4 * i + 0] = fma(d0, d0, r2[4 * i + 0]);
      r2[4 * i + 1] = fma(d1, d1, r2[4 * i + 1]);
      r2[4 * i + 2] = fma(d2, d2, r2[4 * i + 2]);
      r2[4 * i + 3] = fma(d3, d3, r2[4 * i + 3]);
    }
  }
}

// ---- epilogue 1: acc <- K_y(I,J) - acc, K recomputed from the scaled inputs ----------- //
// kc: this thread's slots of the super-tile's kappa cache ([8 pairs of values][256 threads] as double2) or null
template <int KIND>
SCAML_DEVICE void assemble8(Acc8& acc, int I, int J, const Thr& t, const double* xblk, int d, int nv, double os,
                            double diag_add, double2* kc) {
  const int ra = t.rb * kBS + t.g, cb0 = kSB + t.cb * kBS + t.cin + 2 * t.t4;
  const int a0 = I * kSB + ra, b0 = J * kSB + t.cb * kBS + t.cin + 2 * t.t4;
  double r2[16];
  pair_r2_8(r2, xblk, d, ra, cb0);
  if (KIND != SCAML_KERNEL_RBF && kc != nullptr) {
    double kdv[16];
    kappa_n<KIND, 16, true>(r2, r2, kdv);
#pragma unroll
    for (int u = 0; u < 8; ++u) st_stream(kc + (8 + u) * kThreads, make_double2(kdv[2 * u], kdv[2 * u + 1]));
  } else {
    kappa_n<KIND, 16, false>(r2, r2, r2);  // 16 independent exponentials, interleaved
  }
  if (kc != nullptr) {
#pragma unroll
    for (int u = 0; u < 8; ++u) st_stream(kc + u * kThreads, make_double2(r2[2 * u], r2[2 * u + 1]));
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int a = a0 + 8 * i, b = b0 + 8 * j + e;
        double k = os * r2[4 * i + 2 * j + e];
        if (a == b) k += diag_add;
        if (a >= nv || b >= nv) k = (a == b) ? 1.0 : 0.0;
        acc[i][j][e] = k - acc[i][j][e];
      }
}

// ---- epilogue 2: contract the K^-1 super-tile in `acc` with dK/dtheta ----------------- //
// accumulates into gsm[role][0..d-1] (lengthscales), [d] (outputscale), [d+1] (trace W).
// acc is overwritten by t_ab = wgt * W_ab * kd_ab.
template <int KIND>
SCAML_DEVICE void grad8(Acc8& acc, int I, int J, const Thr& t, const double* xblk, const double* av, int d, int nv,
                        double* gsm, const double2* kc) {
  const int ra = t.rb * kBS + t.g, cb0 = kSB + t.cb * kBS + t.cin + 2 * t.t4;
  const int a0 = I * kSB + ra, b0 = J * kSB + t.cb * kBS + t.cin + 2 * t.t4;
  double accS = 0.0, accT = 0.0;
  {
    double r2[16], kdv[16];
    if (kc != nullptr) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const double2 v = ld_stream(kc + u * kThreads);
        r2[2 * u] = v.x, r2[2 * u + 1] = v.y;
        if (KIND != SCAML_KERNEL_RBF) {
          const double2 q = ld_stream(kc + (8 + u) * kThreads);
          kdv[2 * u] = q.x, kdv[2 * u + 1] = q.y;
        }
      }
    } else if (KIND == SCAML_KERNEL_RBF) {
      pair_r2_8(r2, xblk, d, ra, cb0);
      kappa_n<KIND, 16, false>(r2, r2, r2);  // kd == kappa for the RBF kernel (kdv unused)
    } else {
      pair_r2_8(r2, xblk, d, ra, cb0);
      kappa_n<KIND, 16, true>(r2, r2, kdv);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int a = a0 + 8 * i;
      const double ava = av[a];
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int b = b0 + 8 * j + e;
          const double kap = r2[4 * i + 2 * j + e];
          const double kd = (KIND == SCAML_KERNEL_RBF) ? kap : kdv[4 * i + 2 * j + e];
          const bool use = (a >= b) && (a < nv) && (b < nv);
          const double wgt = use ? ((a == b) ? 1.0 : 2.0) : 0.0;
          const double Wab = ava * av[b] - acc[i][j][e];
          const double wk = wgt * Wab;
          accS = fma(wk, kap, accS);
          acc[i][j][e] = wk * kd;
          if (use && a == b) accT += Wab;
        }
    }
  }
  double* gw = gsm + t.role * kMaxP;  // per ROLE: the reduction order does not depend on the warp rotation
  for (int k = 0; k < d; ++k) {
    const double* xr = xblk + k * 128;
    const double2 xb0 = *reinterpret_cast<const double2*>(xr + cb0);
    const double2 xb1 = *reinterpret_cast<const double2*>(xr + cb0 + 8);
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double xa = xr[ra + 8 * i];
      const double d0 = xa - xb0.x, d1 = xa - xb0.y, d2 = xa - xb1.x, d3 = xa - xb1.y;
      s = fma(acc[i][0][0], d0 * d0, s);
      s = fma(acc[i][0][1], d1 * d1, s);
      s = fma(acc[i][1][0], d2 * d2, s);
      s = fma(acc[i][1][1], d3 * d3, s);
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

// 32x32x32 product by the whole CTA on padded tiles: out(r,c) = sum_kk A[kk][r] * B[kk][c];
// warp w owns rows 16 (w >> 2) + {0..15}, columns 8 (w & 3) + {0..7} as 2 DMMA tiles.
typedef double SAcc8[2][2];
SCAML_DEVICE void small_gemm8(SAcc8& o, const double* A, const double* B, const Thr& t) {
  o[0][0] = o[0][1] = o[1][0] = o[1][1] = 0.0;
  const double* ar = A + t.t4 * kLd + 16 * (t.warp >> 2) + t.g;
  const double* br = B + t.t4 * kLd + 8 * (t.warp & 3) + t.g;
#pragma unroll 4
  for (int s = 0; s < 8; ++s) {
    const double a0 = ar[0], a1 = ar[8], b0 = br[0];
    ar += 4 * kLd;
    br += 4 * kLd;
    dmma884(o[0], a0, b0);
    dmma884(o[1], a1, b0);
  }
}
// element (i,e) of the small accumulator <-> row 16 (w >> 2) + 8 i + g, col 8 (w & 3) + 2 t4 + e
SCAML_DEVICE void small_store8_C(double* blk, int ld, const SAcc8& o, const Thr& t, double scale) {
  const int r0 = 16 * (t.warp >> 2) + t.g, c0 = 8 * (t.warp & 3) + 2 * t.t4;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int e = 0; e < 2; ++e) blk[(c0 + e) * ld + r0 + 8 * i] = scale * o[i][e];
}
SCAML_DEVICE void small_store8_R(double* blk, int ld, const SAcc8& o, const Thr& t, double scale) {
  const int r0 = 16 * (t.warp >> 2) + t.g, c0 = 8 * (t.warp & 3) + 2 * t.t4;
#pragma unroll
  for (int i = 0; i < 2; ++i)
    *reinterpret_cast<double2*>(blk + (r0 + 8 * i) * ld + c0) = make_double2(scale * o[i][0], scale * o[i][1]);
}

// ---- factorise + invert the 64x64 diagonal super-tile held in `stage` (see diag_factor in scaml_fit.cuh) //
SCAML_DEVICE void diag_factor8(double* stage, double* dinvc, double* wd00, double* wd10, double* wd11, double* logdet,
                               int* flag, int pivot_base, const Thr& t, int chain_warp) {
  double* T0 = stage;
  double* T1 = stage + kTileS;
  double* T2 = stage + 2 * kTileS;
  double* T3 = stage + 3 * kTileS;
  double* V0 = dinvc;
  double* V1 = dinvc + kTileS;
  double* V2 = dinvc + 2 * kTileS;
  if (t.warp == chain_warp) {
    // L scratch = V1, transpose scratch = V2, X00: C-layout -> V0, R-layout -> T1 and workspace
    const int f = chol_inv_32(T0, V1, V2, V0, T1, wd00, logdet, t.lane);
    if (f && t.lane == 0 && *flag == 0) *flag = pivot_base + f;
  }
  __syncthreads();
  SAcc8 o;
  // L10 = D10 * X00^T      (A = D10 C-layout, B[kk][c] = X00(c,kk) = C-layout X00)
  small_gemm8(o, T2, V0, t);
  __syncthreads();
  small_store8_C(T2, kLd, o, t, 1.0);  // T2 now holds L10 (C-layout)
  __syncthreads();
  // D11 -= L10 L10^T ;  Tm = L10 * X00  (B[kk][c] = X00(kk,c) = R-layout X00 in T1) -> V1 (R-layout)
  small_gemm8(o, T2, T2, t);
  {
    const int r0 = 16 * (t.warp >> 2) + t.g, c0 = 8 * (t.warp & 3) + 2 * t.t4;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int e = 0; e < 2; ++e) T3[(c0 + e) * kLd + r0 + 8 * i] -= o[i][e];
  }
  small_gemm8(o, T2, T1, t);
  small_store8_R(V1, kLd, o, t, 1.0);
  __syncthreads();
  if (t.warp == chain_warp) {
    // L scratch = T1 (X00 R-layout is dead), transpose scratch = T0 (D00 is dead)
    const int f = chol_inv_32(T3, T1, T0, V2, nullptr, wd11, logdet, t.lane);
    if (f && t.lane == 0 && *flag == 0) *flag = pivot_base + kBS + f;
  }
  __syncthreads();
  // X10 = -X11 * Tm        (A[kk][r] = X11(r,kk) = C-layout X11 in V2, B = Tm R-layout in V1)
  small_gemm8(o, V2, V1, t);
  __syncthreads();  // every thread has consumed Tm before V1 is overwritten
  small_store8_C(V1, kLd, o, t, -1.0);
  small_store8_R(wd10, kBS, o, t, -1.0);
  __syncthreads();
}

// D^-1 of diagonal super-tile I (R-layout dense tiles in the workspace) -> dinvc (C-layout padded)
SCAML_DEVICE void load_dinvc8(double* dinvc, double* stage, const double* W, int I, int tid) {
  const double* src[3] = {wtile(W, 2 * I, 2 * I), wtile(W, 2 * I + 1, 2 * I), wtile(W, 2 * I + 1, 2 * I + 1)};
#pragma unroll
  for (int b = 0; b < 3; ++b) {
    half_async8(stage + b * kTileS, src[b], tid);
    half_async8(stage + b * kTileS + kHalfS, src[b] + kHalfG, tid);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
#pragma unroll
  for (int b = 0; b < 3; ++b) {
#pragma unroll 4
    for (int idx = tid; idx < kTile; idx += kThreads) {
      const int r = idx >> 5, c = idx & 31;
      dinvc[b * kTileS + c * kLd + r] = stage[b * kTileS + r * kLd + c];
    }
  }
  __syncthreads();
}

// out[r] = sum_kk D^-1(r,kk) v[kk] over a 64x64 lower-triangular D^-1 held as dinvc (C-layout tiles);
// four 16-wide kk quarters, summed in a fixed order
SCAML_DEVICE void dinv_matvec8(double* out, const double* dinvc, const double* v, double* red, const Thr& t) {
  const int r = t.tid & 63, q = t.tid >> 6;
  double s = 0.0;
  for (int kk = q * 16; kk < q * 16 + 16; ++kk) {
    if (kk > r) break;
    const int rb_ = r >> 5, kb_ = kk >> 5;
    const double* blk = dinvc + ((rb_ == 0) ? 0 : (kb_ == 0 ? kTileS : 2 * kTileS));
    s = fma(blk[(kk & 31) * kLd + (r & 31)], v[kk], s);
  }
  red[q * 64 + r] = s;
  __syncthreads();
  if (t.tid < 64) out[t.tid] = (red[t.tid] + red[64 + t.tid]) + (red[128 + t.tid] + red[192 + t.tid]);
  __syncthreads();
}

template <int KIND>
__global__ void __launch_bounds__(kThreads, 2) scaml_fit8_kernel(const FitParams p) {
  SCAML_DYN_SMEM(double, sm);
  Thr t = make_thr();
  const int d = p.d, P = p.d + 2, n_pad_max = p.n_pad;
  double* stage = sm;              // 2 stages x 4 padded half tiles | 4 full padded tiles (C_in / S / diag)
  double* dinvc = stage + kStage;  // 3 padded tiles
  double* yv = dinvc + 3 * kTileS;
  double* zv = yv + n_pad_max;
  double* av = zv + n_pad_max;
  double* red = av + n_pad_max;  // 256
  double* gsm = red + 256;       // kWarps * kMaxP
  double* par = gsm + kWarps * kMaxP;
  double* th = par;
  double* lp = par + kMaxP;
  double* dlp = par + 2 * kMaxP;
  double* chain = par + 3 * kMaxP;
  double* invl = par + 4 * kMaxP;  // reciprocal lengthscales (x * (1/l): an FP64 division is ~10 pipe slots)
  double* scal = par + 5 * kMaxP;  // [0] logdet
  int* flag = reinterpret_cast<int*>(scal + 8);

  double* W = p.workspace + (size_t)blockIdx.x * p.ws_stride;
  double2* kcache = (p.kcache && p.mode == kModeLmlGrad)
                        ? reinterpret_cast<double2*>(W + fit_tile_doubles(p.n_pad))
                        : nullptr;  // kappa cache behind the tiles (see scaml_fit.cuh)
  const int E = p.M * p.R;
  const scaml_hyper_spec& sp = p.spec;
  // warp w of every co-resident CTA sits on sub-partition w % 4; the eight tile roles carry unequal work on
  // diagonal super-tiles, so the warp -> role map is rotated by the CTA's slot on its SM and by the evaluation
  // index; the pivot chains run on a warp that holds an idle-on-diagonal role.  Results do not depend on the
  // rotation (role-indexed reductions).
  const int slot = p.sms > 0 ? (int)(blockIdx.x / p.sms) : 0;

  // dynamic work distribution over a global counter (see scaml_fit.cuh)
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
    const int rot = 2 * slot + it;
    set_role(t, (t.warp + rot) & (kWarps - 1));
    const int chain_warp = (2 - rot) & (kWarps - 1);  // the warp whose role is 2: tile (0,1), idle on diagonals

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
    for (int i = t.tid; i < kWarps * kMaxP; i += kThreads) gsm[i] = 0.0;
    __syncthreads();
    const double os = th[d];
    const double* Xm = p.X + (size_t)m * p.n_max * d;
    {
      const double* ym = p.y + (size_t)m * p.n_max;
      for (int i = t.tid; i < n_pad; i += kThreads) yv[i] = (i < nv) ? ym[i] : 0.0;
    }
    __syncthreads();

    Acc8 acc;
    double pig = 0.0;
    bool failed = false;

    // ================= phase B: blocked left-looking Cholesky ========================== //
    // in-kernel psd_safe_cholesky jitter ladder (p.ladder), see scaml_fit.cuh
    double jit = p.jitter ? p.jitter[e] : 0.0;
    for (int attempt = 0;; ++attempt) {
    const double diag_add = th[d + 1] + jit;
    failed = false;
    for (int J = 0; J < NS && !failed; ++J) {
      for (int I = J; I < NS; ++I) {
        const bool diag = (I == J);
        const bool idle = diag && (t.rb == 0 && t.cb == 1);  // tile (0,1) of a diagonal super-tile
        acc_zero(acc);
        double xp[kXh];
        xpre_load8(xp, Xm, I, J, nv, d, t.tid);
        CholSrc src{W, I, J};
        gemm_global8(acc, src, stage, t, diag, false, pig, nullptr);
        xblk_store8(stage, xp, Xm, invl, I, J, nv, d, t.tid);  // stage is idle: x-block lives there
        __syncthreads();
        if (!idle)
          assemble8<KIND>(acc, I, J, t, stage, d, nv, os, diag_add,
                          kcache ? kcache + ((size_t)(tri(I) + J) * 16) * kThreads + t.tid : nullptr);
        __syncthreads();  // x-block consumed before C_in overwrites it
        if (!idle) store8_C(stage + (t.rb * 2 + t.cb) * kTileS, kLd, acc, t, 1.0);
        __syncthreads();
        if (diag) {
          diag_factor8(stage, dinvc, wtile_w(W, 2 * J, 2 * J), wtile_w(W, 2 * J + 1, 2 * J),
                       wtile_w(W, 2 * J + 1, 2 * J + 1), &scal[0], flag, J * kSB, t, chain_warp);
          if (*flag != 0) {
            failed = true;
            break;
          }
        } else {
          // L(I,J) = C * D^-T
          acc_zero(acc);
          smem_trsm8(acc, stage, dinvc, t);
          store8_C(wtile_w(W, 2 * I + t.rb, 2 * J + t.cb), kBS, acc, t, 1.0);
          __syncthreads();
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
      load_dinvc8(dinvc, stage, W, I, t.tid);
      pig = 0.0;
      for (int J = 0; J < I; ++J) {
        acc_zero(acc);
        TrtriSrc src{W, I, J};
        gemm_global8(acc, src, stage, t, false, J == 0, pig, zv);
        store8_R(stage + (t.rb * 2 + t.cb) * kTileS, kLd, acc, t, 1.0);  // S, R-layout tiles (kb2, cb)
        __syncthreads();
        acc_zero(acc);
        smem_trtri8(acc, dinvc, stage, t);
        store8_R(wtile_w(W, 2 * I + t.rb, 2 * J + t.cb), kBS, acc, t, -1.0);
        if (p.mode == kModeFactorize)
          store8_C(p.linv_out + ((size_t)m * tri(n_pad_max / kBS) + tri(2 * I + t.rb) + 2 * J + t.cb) * kTile, kBS, acc,
                   t, -1.0);
        __syncthreads();
      }
      // z_I = D_I^-1 (y_I - sum_{K<I} L(I,K) z_K)
      red[t.tid] = pig;
      __syncthreads();
      if (t.tid < 64)  // av: scratch here
        av[t.tid] = yv[I * kSB + t.tid] - ((red[t.tid] + red[64 + t.tid]) + (red[128 + t.tid] + red[192 + t.tid]));
      __syncthreads();
      dinv_matvec8(zv + I * kSB, dinvc, av, red, t);
      if (p.mode == kModeFactorize) {
        // diagonal tiles of L^-1 in C-layout = dinvc (padded -> dense)
        double* lo = p.linv_out + (size_t)m * tri(n_pad_max / kBS) * kTile;
        for (int i = t.tid; i < kTile; i += kThreads) {
          const int si = (i >> 5) * kLd + (i & 31);
          lo[(size_t)(tri(2 * I) + 2 * I) * kTile + i] = dinvc[si];
          lo[(size_t)(tri(2 * I + 1) + 2 * I) * kTile + i] = dinvc[kTileS + si];
          lo[(size_t)(tri(2 * I + 1) + 2 * I + 1) * kTile + i] = dinvc[2 * kTileS + si];
        }
      }
    }
    __syncthreads();

    if (p.mode == kModeFactorize) {
      // alpha = L^-T z : warp w owns 32-columns bj = w, w+8, ... ; lane = column inside the tile
      const int NB = 2 * NS;
      for (int bj = t.warp; bj < NB; bj += kWarps) {
        double s = 0.0;
        for (int bi = bj; bi < NB; ++bi) {
          const double* blk = wtile(W, bi, bj);  // R-layout: (r,c) at r*32+c
#pragma unroll 8
          for (int r = 0; r < kBS; ++r) s = fma(__ldcg(blk + r * kBS + t.lane), zv[bi * kBS + r], s);
        }
        p.alpha_out[(size_t)m * n_pad_max + bj * kBS + t.lane] = s;
      }
      for (int i = n_pad + t.tid; i < n_pad_max; i += kThreads) p.alpha_out[(size_t)m * n_pad_max + i] = 0.0;
      if (t.tid == 0) p.info[e] = 0;
      continue;
    }

    // ================= phase D: K^-1 super-tiles, fused gradient contraction ========== //
    for (int I = 0; I < NS; ++I) {
      for (int jj = 0; jj <= I; ++jj) {  // diagonal super-tile first: it completes alpha_I
        const bool diag = (jj == 0);
        const int J = diag ? I : jj - 1;
        const bool idle = diag && (t.rb == 0 && t.cb == 1);
        acc_zero(acc);
        pig = 0.0;
        double xp[kXh];
        xpre_load8(xp, Xm, I, J, nv, d, t.tid);
        LauumSrc src{W, I, J, NS};
        gemm_global8(acc, src, stage, t, diag, diag, pig, zv);
        xblk_store8(stage, xp, Xm, invl, I, J, nv, d, t.tid);
        if (diag) red[t.tid] = pig;
        __syncthreads();
        if (diag) {
          if (t.tid < 64)
            av[I * kSB + t.tid] = (red[t.tid] + red[64 + t.tid]) + (red[128 + t.tid] + red[192 + t.tid]);
          __syncthreads();
        }
        if (!idle)
          grad8<KIND>(acc, I, J, t, stage, av, d, nv, gsm,
                      kcache ? kcache + ((size_t)(tri(I) + J) * 16) * kThreads + t.tid : nullptr);
        __syncthreads();  // x-block consumed before the next product stages tiles over it
      }
    }
    __syncthreads();
    // quad = z^T z (fixed order), final scalars
    {
      double q = 0.0;
      for (int i = t.tid; i < n_pad; i += kThreads) q = fma(zv[i], zv[i], q);
      q = warp_sum(q);
      if (t.lane == 0) red[t.warp] = q;
      __syncthreads();
      if (t.tid < P) {
        double g = 0.0;
        for (int w = 0; w < kWarps; ++w) g += gsm[w * kMaxP + t.tid];
        const double gt = (t.tid < d) ? 0.5 * os * g / th[t.tid] : 0.5 * g;
        p.grad[(size_t)e * P + t.tid] = (gt + dlp[t.tid]) * chain[t.tid] / (double)nv;
      }
      if (t.tid == 0) {
        double quad = 0.0;
        for (int w = 0; w < kWarps; ++w) quad += red[w];
        double prior = 0.0;
        for (int k = 0; k < P; ++k) prior += lp[k];
        p.lml[e] = (-0.5 * (quad + scal[0] + (double)nv * kLog2Pi) + prior) / (double)nv;
        p.info[e] = 0;
      }
    }
  }
}

}  // namespace f8
}  // namespace scaml
