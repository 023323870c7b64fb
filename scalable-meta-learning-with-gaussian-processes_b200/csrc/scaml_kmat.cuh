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

struct KmatParams {
  const double* X;
  const int32_t* n_valid;
  const double* theta;  // [M][P] constrained
  double* K;
  int M, n_max, d, nt;  // nt = tiles per dimension
  long long items;      // M * nt*(nt+1)/2
};

template <int KIND>
__global__ void __launch_bounds__(256) scaml_kmat_kernel(const KmatParams p) {
  SCAML_DYN_SMEM(double, sm);
  const Thr t = make_thr();
  const int d = p.d, P = d + 2;
  double* xa = sm;            // [d][64] rows of tile ti (scaled)
  double* xb = sm + d * kSB;  // [d][64] rows of tile tj
  const int pairs = (p.nt * (p.nt + 1)) / 2;
  const bool vec_ok = (p.n_max % 2) == 0;
  for (long long it = blockIdx.x; it < p.items; it += gridDim.x) {
    const int m = (int)(it / pairs);
    int pr = (int)(it - (long long)m * pairs);
    int ti = 0;
    while (tri(ti + 1) <= pr) ++ti;
    const int tj = pr - tri(ti);
    const int nv = p.n_valid ? p.n_valid[m] : p.n_max;
    const double* th = p.theta + (size_t)m * P;
    const double* Xm = p.X + (size_t)m * p.n_max * d;
    __syncthreads();
    for (int i = t.tid; i < kSB * d; i += 256) {
      const int r = i / d, k = i - r * d;
      const int ga = ti * kSB + r, gb = tj * kSB + r;
      const double il = 1.0 / th[k];
      xa[k * kSB + r] = (ga < p.n_max) ? Xm[(size_t)ga * d + k] * il : 0.0;
      xb[k * kSB + r] = (gb < p.n_max) ? Xm[(size_t)gb * d + k] * il : 0.0;
    }
    __syncthreads();
    const double os = th[d], noise = th[d + 1];
    const int ra = t.rb * kBS + t.rin, cb_ = t.cb * kBS + t.cin;
    double r2[4][4];
    acc_zero(r2);
    for (int k = 0; k < d; ++k) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const double df = xa[k * kSB + ra + i] - xb[k * kSB + cb_ + j];
          r2[i][j] = fma(df, df, r2[i][j]);
        }
    }
    const int a0 = ti * kSB + ra, b0 = tj * kSB + cb_;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int a = a0 + i, b = b0 + j;
        double k = os * kappa_of<KIND>(r2[i][j]);
        if (a == b) k += noise;
        if (a >= nv || b >= nv) k = (a == b) ? 1.0 : 0.0;
        r2[i][j] = k;
      }
    double* Km = p.K + (size_t)m * p.n_max * p.n_max;
    const bool full = (a0 + 3 < p.n_max) && (b0 + 3 < p.n_max) && vec_ok;
    if (full) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        double* q = Km + (size_t)(a0 + i) * p.n_max + b0;
        *reinterpret_cast<double2*>(q) = make_double2(r2[i][0], r2[i][1]);
        *reinterpret_cast<double2*>(q + 2) = make_double2(r2[i][2], r2[i][3]);
      }
      if (ti != tj) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          double* q = Km + (size_t)(b0 + j) * p.n_max + a0;
          *reinterpret_cast<double2*>(q) = make_double2(r2[0][j], r2[1][j]);
          *reinterpret_cast<double2*>(q + 2) = make_double2(r2[2][j], r2[3][j]);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int a = a0 + i, b = b0 + j;
          if (a < p.n_max && b < p.n_max) {
            Km[(size_t)a * p.n_max + b] = r2[i][j];
            if (ti != tj) Km[(size_t)b * p.n_max + a] = r2[i][j];
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
  const size_t smem = sizeof(double) * 2 * (size_t)d * kSB;
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
