// Weighted ScaML-GP prediction (K6-K9): for a tile of 64 candidates a CTA walks over the
// tasks of its split, builds k*(X_m, candidates) once in shared memory, streams the packed
// L_m^-1 tiles (cp.async, double buffered) through the same 4x4 register-tiled micro-kernel
// as the fit, and folds  w_m (ybar_m + ystd_m k*^T alpha_m)  and
// w_m^2 ystd_m^2 (s_m - ||L_m^-1 k*||^2)  into per-candidate accumulators that stay in
// registers for the whole task loop -> the reduction over tasks happens inside the kernel
// in a fixed order (deterministic, no atomics).
// Reference: _compute_target_prior, scamlgp/model.py:108-135; posterior A.7 of SURVEY.md.
#pragma once
#include "scaml_device.cuh"
#include "scaml_tile256.cuh"

namespace scaml {

constexpr int kTB = 64;  // candidates per tile

struct PredParams {
  const double* X;
  const int32_t* n_valid;
  const double* theta;  // [M][P] constrained
  const double* linv;   // packed C-layout tiles
  const double* alpha;  // [M][n_pad]
  const double* ybar;
  const double* ystd;
  const double* w;
  const double* Xc;  // [B][d]
  double* mean;
  double* var;
  double* part;  // [nsplit][2][B] when nsplit > 1
  int M, n_max, n_pad, d, B, nsplit, ntile;
};

inline size_t predict_smem_bytes(int n_pad, int d) {
  return sizeof(double) *
         ((size_t)n_pad * kTB + 4096 + (size_t)d * n_pad + (size_t)n_pad + 2 * (size_t)d * kTB + 256 + 128 + 8);
}
inline int predict_nsplit(int M, int B, int num_sms) {
  const int ntile = (B + kTB - 1) / kTB;
  int ns = (2 * num_sms + ntile - 1) / ntile;
  if (ns < 1) ns = 1;
  if (ns > M) ns = M;
  return ns;
}
inline size_t predict_workspace_bytes(int M, int n_pad, int d, int B, int num_sms) {
  (void)n_pad;
  (void)d;
  const int ns = predict_nsplit(M, B, num_sms);
  return ns > 1 ? sizeof(double) * 2 * (size_t)ns * (size_t)B : 0;
}

template <int KIND>
__global__ void __launch_bounds__(256, 1) scaml_predict_kernel(const PredParams p) {
  SCAML_DYN_SMEM(double, sm);
  const Thr t = make_thr();
  const int d = p.d, P = d + 2, n_pad = p.n_pad;
  double* kst = sm;                           // n_pad x 64 as 32x32 tiles (ab*2+cb), [a][c]
  double* ast = kst + (size_t)n_pad * kTB;    // 2 stages x 2 tiles
  double* xst = ast + 4096;                   // [d][n_pad]
  double* alp = xst + (size_t)d * n_pad;      // [n_pad]
  double* xcr = alp + n_pad;                  // [d][64] raw candidates
  double* xcs = xcr + d * kTB;                // [d][64] scaled for the current task
  double* red = xcs + d * kTB;                // 256
  double* vsq = red + 256;                    // 128
  const long long lstride = (long long)tri(n_pad / kBS) * kTile;
  const int items = p.ntile * p.nsplit;
  const int mper = (p.M + p.nsplit - 1) / p.nsplit;

  for (int it = blockIdx.x; it < items; it += gridDim.x) {
    const int ct = it % p.ntile, spx = it / p.ntile;
    const int b0 = ct * kTB;
    const int m_lo = spx * mper, m_hi = (m_lo + mper < p.M) ? m_lo + mper : p.M;
    __syncthreads();
    for (int i = t.tid; i < kTB * d; i += 256) {
      const int c = i / d, k = i - c * d;
      xcr[k * kTB + c] = (b0 + c < p.B) ? p.Xc[(size_t)(b0 + c) * d + k] : 0.0;
    }
    double macc = 0.0, vacc = 0.0;
    for (int m = m_lo; m < m_hi; ++m) {
      const double wm = p.w[m];
      if (wm == 0.0) continue;
      const int nv = p.n_valid ? p.n_valid[m] : p.n_max;
      const int NS = (nv + kSB - 1) / kSB, npt = NS * kSB;
      const double* th = p.theta + (size_t)m * P;
      const double os = th[d];
      __syncthreads();
      {
        const double* Xm = p.X + (size_t)m * p.n_max * d;
        for (int i = t.tid; i < npt * d; i += 256) {
          const int a = i / d, k = i - a * d;
          xst[k * n_pad + a] = (a < nv) ? Xm[(size_t)a * d + k] / th[k] : 0.0;
        }
        for (int i = t.tid; i < npt; i += 256) alp[i] = p.alpha[(size_t)m * n_pad + i];
        for (int i = t.tid; i < kTB * d; i += 256) {
          const int k = i / kTB;
          xcs[i] = xcr[i] / th[k];
        }
      }
      __syncthreads();
      // ---- k*(X_m, candidates) and the mean partials -------------------------------- //
      {
        const int c = t.tid & 63, q = t.tid >> 6;
        double mu = 0.0;
        for (int a = q; a < npt; a += 4) {
          double r2 = 0.0;
          for (int k = 0; k < d; ++k) {
            const double df = xst[k * n_pad + a] - xcs[k * kTB + c];
            r2 = fma(df, df, r2);
          }
          const double kv = (a < nv) ? os * kappa_of<KIND>(r2) : 0.0;
          kst[((a >> 5) * 2 + (c >> 5)) * kTile + (a & 31) * kBS + (c & 31)] = kv;
          mu = fma(kv, alp[a], mu);
        }
        red[q * 64 + c] = mu;
      }
      __syncthreads();
      // ---- V = L^-1 k*  super-tile by super-tile, column sums of squares ------------- //
      const double* Lm = p.linv + (size_t)m * lstride;
      double vs[4] = {0.0, 0.0, 0.0, 0.0};
      for (int I = 0; I < NS; ++I) {
        double acc[4][4];
        acc_zero(acc);
        const int n = 2 * I + 2;
        // chunk ck: A tiles (2I+rb, ck) (null above the diagonal), B = kst tiles (ck, cb)
        {
          const double* a0 = Lm + (size_t)(tri(2 * I) + 0) * kTile;
          tile_async256(ast, a0, t.tid);
          tile_async256(ast + kTile, Lm + (size_t)(tri(2 * I + 1) + 0) * kTile, t.tid);
          cp_async_commit();
        }
        for (int ck = 0; ck < n; ++ck) {
          double* st = ast + (ck & 1) * 2 * kTile;
          if (ck + 1 < n) {
            double* sn = ast + ((ck + 1) & 1) * 2 * kTile;
            if (ck + 1 <= 2 * I) tile_async256(sn, Lm + (size_t)(tri(2 * I) + ck + 1) * kTile, t.tid);
            tile_async256(sn + kTile, Lm + (size_t)(tri(2 * I + 1) + ck + 1) * kTile, t.tid);
            cp_async_commit();
            cp_async_wait<1>();
          } else {
            cp_async_wait<0>();
          }
          __syncthreads();
          if (ck <= 2 * I + t.rb)
            mma_chunk(acc, st + t.rb * kTile + t.rin, kst + (ck * 2 + t.cb) * kTile + t.cin);
          __syncthreads();
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int i = 0; i < 4; ++i) vs[j] = fma(acc[i][j], acc[i][j], vs[j]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double s = vs[j];
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 8);
        s += __shfl_xor_sync(0xffffffffu, s, 16);
        if ((t.lane >> 2) == 0) vsq[t.rb * 64 + t.cb * kBS + t.cin + j] = s;
      }
      __syncthreads();
      if (t.tid < kTB) {
        const double mu = ((red[t.tid] + red[64 + t.tid]) + red[128 + t.tid]) + red[192 + t.tid];
        const double ssq = vsq[t.tid] + vsq[64 + t.tid];
        const double ys = p.ystd[m];
        macc = fma(wm, p.ybar[m] + ys * mu, macc);
        vacc = fma(wm * wm * ys * ys, os - ssq, vacc);
      }
    }
    if (t.tid < kTB && b0 + t.tid < p.B) {
      if (p.nsplit == 1) {
        p.mean[b0 + t.tid] = macc;
        p.var[b0 + t.tid] = vacc;
      } else {
        p.part[((size_t)spx * 2 + 0) * p.B + b0 + t.tid] = macc;
        p.part[((size_t)spx * 2 + 1) * p.B + b0 + t.tid] = vacc;
      }
    }
  }
}

__global__ void scaml_predict_reduce_kernel(const double* part, double* mean, double* var, int nsplit, int B) {
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
    double m = 0.0, v = 0.0;
    for (int s = 0; s < nsplit; ++s) {
      m += part[((size_t)s * 2 + 0) * B + b];
      v += part[((size_t)s * 2 + 1) * B + b];
    }
    mean[b] = m;
    var[b] = v;
  }
}

template <int KIND>
int launch_predict_k(const PredParams& p, int grid, size_t smem, void* stream) {
#ifdef SCAML_EMU
  (void)stream;
  cuemu::launch(dim3(grid), dim3(256), smem, scaml_predict_kernel<KIND>, p);
  return 0;
#else
  cudaError_t err =
      cudaFuncSetAttribute(scaml_predict_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return (int)err;
  scaml_predict_kernel<KIND><<<grid, 256, smem, (cudaStream_t)stream>>>(p);
  return (int)cudaGetLastError();
#endif
}

inline int launch_predict_weighted(const double* X, const int32_t* n_valid, const double* theta, const double* linv,
                                   const double* alpha, const double* ybar, const double* ystd, const double* w,
                                   const double* Xc, double* mean, double* var, double* workspace, int M, int n_max,
                                   int n_pad, int d, int B, int kernel, int num_sms, void* stream) {
  PredParams p;
  p.X = X, p.n_valid = n_valid, p.theta = theta, p.linv = linv, p.alpha = alpha, p.ybar = ybar, p.ystd = ystd;
  p.w = w, p.Xc = Xc, p.mean = mean, p.var = var, p.part = workspace;
  p.M = M, p.n_max = n_max, p.n_pad = n_pad, p.d = d, p.B = B;
  p.ntile = (B + kTB - 1) / kTB;
  p.nsplit = predict_nsplit(M, B, num_sms);
  const size_t smem = predict_smem_bytes(n_pad, d);
  if (smem > 227 * 1024) return SCAML_E_SMEM;
  long long items = (long long)p.ntile * p.nsplit;
  int grid = (int)(items < num_sms ? items : num_sms);
  int rc;
  switch (kernel) {
    case SCAML_KERNEL_RBF: rc = launch_predict_k<SCAML_KERNEL_RBF>(p, grid, smem, stream); break;
    case SCAML_KERNEL_MATERN12: rc = launch_predict_k<SCAML_KERNEL_MATERN12>(p, grid, smem, stream); break;
    case SCAML_KERNEL_MATERN32: rc = launch_predict_k<SCAML_KERNEL_MATERN32>(p, grid, smem, stream); break;
    default: rc = launch_predict_k<SCAML_KERNEL_MATERN52>(p, grid, smem, stream); break;
  }
  if (rc != 0 || p.nsplit == 1) return rc;
#ifdef SCAML_EMU
  cuemu::launch(dim3(1), dim3(64), 0, scaml_predict_reduce_kernel, (const double*)workspace, mean, var, p.nsplit, B);
  return 0;
#else
  const int rb = (B + 255) / 256;
  scaml_predict_reduce_kernel<<<rb < 1184 ? rb : 1184, 256, 0, (cudaStream_t)stream>>>(workspace, mean, var,
                                                                                      p.nsplit, B);
  return (int)cudaGetLastError();
#endif
}

}  // namespace scaml
