// Standalone batched kernel-matrix assembly (K1): K[m] = s * kappa(X_m / l) + noise * I,
// dense [M][n_max][n_max] fp64.  HBM-store bound (M * n^2 * 8 bytes); symmetry is used so
// every exp() is evaluated once: a CTA computes one 64x64 tile (ti >= tj) in 4x4 register
// patches and stores it twice (as (ti,tj) rows and as the transposed (tj,ti) rows), both
// with 32 B per thread / 128-256 B contiguous per row segment.
// Reference call site: covar_module(X) + likelihood noise inside mll(model(X), y),
// scamlgp/utils.py:175-177; kernels scamlgp/model.py:44-70.
#pragma once
#include "scaml_device.cuh"
#include "scaml_tile256.cuh"

namespace scaml {

constexpr int kTLd = 65;  // odd row stride: column reads of the staged tile are (almost) conflict free

struct KmatParams {
  const double* X;
  const int32_t* n_valid;
  const double* theta;  // [M][P] constrained
  double* K;
  int M, n_max, d, nt;  // nt = tiles per dimension
  long long items;      // M * nt*(nt+1)/2
};

template <int KIND>
__global__ void __launch_bounds__(256, 4) scaml_kmat_kernel(const KmatParams p) {
  SCAML_DYN_SMEM(double, sm);
  const Thr t = make_thr();
  const int d = p.d, P = d + 2;
  double* xa = sm;            // [d][64] rows of tile ti (scaled)
  double* xb = sm + d * kSB;  // [d][64] rows of tile tj
  double* T = xb + d * kSB;   // 64 x kTLd staging tile
  double* invl = T + kSB * kTLd;  // [kMaxP] reciprocal lengthscales of the current task
  const int pairs = (p.nt * (p.nt + 1)) / 2;
  const bool vec_ok = (p.n_max % 2) == 0;
  for (long long it = blockIdx.x; it < p.items; it += gridDim.x) {
    const int m = (int)(it / pairs);
    const int pr = (int)(it - (long long)m * pairs);
    int ti = 0;
    while (tri(ti + 1) <= pr) ++ti;
    const int tj = pr - tri(ti);
    const int nv = p.n_valid ? p.n_valid[m] : p.n_max;
    const double* th = p.theta + (size_t)m * P;
    const double* Xm = p.X + (size_t)m * p.n_max * d;
    // 128 points (64 rows of tile ti, 64 of tile tj) x d coordinates, scaled by the lengthscales:
    // thread -> (point, coordinate parity); no runtime integer division.  The raw coordinates are fetched into
    // registers BEFORE the barrier (their latency overlaps the wait for the previous tile's stores) and scaled by
    // reciprocal lengthscales (an FP64 division per element was 20 % of this kernel's stall samples,
    // profiles/r2_kmat_kernel_source_hotspots.txt).
    const int pt = t.tid & 127, kh = t.tid >> 7;
    const int r = pt & 63;
    const int ga = (pt < kSB) ? ti * kSB + r : tj * kSB + r;
    double xq[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = kh + 2 * u;
      xq[u] = (k < d && ga < p.n_max) ? __ldg(Xm + (size_t)ga * d + k) : 0.0;
    }
    __syncthreads();  // previous tile fully written out before xa / xb / T / invl are overwritten
    if (t.tid < d) invl[t.tid] = 1.0 / th[t.tid];
    __syncthreads();
    {
      double* dst = (pt < kSB) ? xa : xb;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = kh + 2 * u;
        if (k < d) dst[k * kSB + r] = xq[u] * invl[k];
      }
      for (int k = kh + 8; k < d; k += 2) dst[k * kSB + r] = (ga < p.n_max) ? __ldg(Xm + (size_t)ga * d + k) * invl[k] : 0.0;
    }
    __syncthreads();
    const double os = th[d], noise = th[d + 1];
    const int ra = t.rb * kBS + t.rin, cb_ = t.cb * kBS + t.cin;
    double r2f[16];  // [i][j] -> 4 i + j
#pragma unroll
    for (int u = 0; u < 16; ++u) r2f[u] = 0.0;
    for (int k = 0; k < d; ++k) {
      const double2 a01 = *reinterpret_cast<const double2*>(xa + k * kSB + ra);
      const double2 a23 = *reinterpret_cast<const double2*>(xa + k * kSB + ra + 2);
      const double2 b01 = *reinterpret_cast<const double2*>(xb + k * kSB + cb_);
      const double2 b23 = *reinterpret_cast<const double2*>(xb + k * kSB + cb_ + 2);
      const double av[4] = {a01.x, a01.y, a23.x, a23.y};
      const double bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const double df = av[i] - bv[j];
          r2f[4 * i + j] = fma(df, df, r2f[4 * i + j]);
        }
    }
    kappa_n<KIND, 16, false>(r2f, r2f, r2f);  // 16 independent exponentials, interleaved
    const int a0 = ti * kSB + ra, b0 = tj * kSB + cb_;
    // interior tiles (off the diagonal, fully inside the valid range) need no per-element masks: CTA-uniform
    const bool interior = (ti != tj) && (ti * kSB + kSB <= nv);
    if (interior) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) T[(ra + i) * kTLd + cb_ + j] = os * r2f[4 * i + j];
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int a = a0 + i, b = b0 + j;
          double k = os * r2f[4 * i + j];
          if (a == b) k += noise;
          if (a >= nv || b >= nv) k = (a == b) ? 1.0 : 0.0;
          T[(ra + i) * kTLd + cb_ + j] = k;
        }
    }
    // stage the tile in shared memory, then write it (and, off the diagonal, its transpose) in whole rows:
    // one warp instruction = 512 B (tile) / 256 B (transpose) contiguous, every 32-byte sector written once
    __syncthreads();
    double* Km = p.K + (size_t)m * p.n_max * p.n_max;
    const int row0 = ti * kSB, col0 = tj * kSB;
    for (int r = t.warp; r < kSB; r += 8) {
      const int ga = row0 + r;
      if (ga >= p.n_max) break;
      double* q = Km + (size_t)ga * p.n_max + col0;
      const double v0 = T[r * kTLd + 2 * t.lane], v1 = T[r * kTLd + 2 * t.lane + 1];
      if (vec_ok && col0 + 2 * t.lane + 1 < p.n_max) {
        *reinterpret_cast<double2*>(q + 2 * t.lane) = make_double2(v0, v1);
      } else {
        if (col0 + 2 * t.lane < p.n_max) q[2 * t.lane] = v0;
        if (col0 + 2 * t.lane + 1 < p.n_max) q[2 * t.lane + 1] = v1;
      }
    }
    if (ti != tj) {
      for (int c = t.warp; c < kSB; c += 8) {
        const int gb = col0 + c;
        if (gb >= p.n_max) break;
        double* q = Km + (size_t)gb * p.n_max + row0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int r = t.lane + 32 * h;
          if (row0 + r < p.n_max) q[r] = T[r * kTLd + c];
        }
      }
    }
  }
}

template <int KIND>
int launch_kmat_k(const KmatParams& p, int grid, size_t smem, void* stream) {
#ifdef SCAML_EMU
  (void)stream;
  cuemu::launch(dim3(grid), dim3(256), smem, scaml_kmat_kernel<KIND>, p);
  return 0;
#else
  if (smem > 48 * 1024) {
    cudaError_t err = cudaFuncSetAttribute(scaml_kmat_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return (int)err;
  }
  scaml_kmat_kernel<KIND><<<grid, 256, smem, (cudaStream_t)stream>>>(p);
  return (int)cudaGetLastError();
#endif
}

inline int launch_kmat(const double* X, const int32_t* n_valid, const double* theta, double* K, int M, int n_max,
                       int d, int kernel, void* stream) {
  KmatParams p;
  p.X = X, p.n_valid = n_valid, p.theta = theta, p.K = K, p.M = M, p.n_max = n_max, p.d = d;
  p.nt = (n_max + kSB - 1) / kSB;
  p.items = (long long)M * ((p.nt * (p.nt + 1)) / 2);
  const size_t smem = sizeof(double) * (2 * (size_t)d * kSB + (size_t)kSB * kTLd + kMaxP);
  long long g = p.items;
#ifdef SCAML_EMU
  if (g > 4) g = 4;
#else
  if (g > 148LL * 8 * 4) g = 148LL * 8 * 4;
#endif
  switch (kernel) {
    case SCAML_KERNEL_RBF: return launch_kmat_k<SCAML_KERNEL_RBF>(p, (int)g, smem, stream);
    case SCAML_KERNEL_MATERN12: return launch_kmat_k<SCAML_KERNEL_MATERN12>(p, (int)g, smem, stream);
    case SCAML_KERNEL_MATERN32: return launch_kmat_k<SCAML_KERNEL_MATERN32>(p, (int)g, smem, stream);
    default: return launch_kmat_k<SCAML_KERNEL_MATERN52>(p, (int)g, smem, stream);
  }
}

}  // namespace scaml
