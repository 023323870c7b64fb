// Weighted ScaML-GP prediction (K6-K9): for a tile of CT (64 or 32) candidates a CTA walks over the
// tasks of its split.  Per task it (1) builds k*(X_m, candidates) once in shared memory (padded
// k-major 32x32 tiles) together with the mean partials k*^T alpha, (2) streams the packed L_m^-1
// tiles through a 3-stage cp.async pipeline (one barrier per 32-deep chunk, the pipeline does not
// drain between the super-rows of a task) and forms V = L^-1 k* on the FP64 tensor cores
// (mma.sync m8n8k4 -> DMMA; 8 warps, warp tile 32 x CT/4; the zero half of the diagonal tiles is
// skipped), (3) folds  w_m (ybar_m + ystd_m k*^T alpha_m)  and  w_m^2 ystd_m^2 (s_m - ||V||^2)  into
// per-candidate accumulators that stay in registers for the whole task loop -> the reduction over
// tasks happens inside the kernel in a fixed order (deterministic, no atomics).
// Reference: _compute_target_prior, scamlgp/model.py:108-135; posterior A.7 of SURVEY.md.
#pragma once
#include "scaml_device.cuh"

namespace scaml {

constexpr int kPLd = 36;               // padded row stride of a staged tile (conflict-free DMMA fragment loads)
constexpr int kPTile = kBS * kPLd;     // 1152 doubles
constexpr int kPStages = 3;
constexpr int kPredThreads = 256;

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
  int M, n_max, n_pad, d, B, nsplit, ntile, ct, alias;
  // CROSS variant (conditioning on target data, see scaml_cond.cuh)
  const double* condA;  // [M][n_pad][n_tp]  A_m = K_m^-1 K_m(X_m, X_t)
  double* cxp;          // [nsplit][B][n_tp] partials of sum_m c_m k*_m^T A_m
  int n_tp;
};

// shared memory (doubles): kst | stage | [xst | alp | xcs] (aliased onto stage when alias) | xcr | red | vsq
#ifdef SCAML_EMU
inline
#else
__host__ __device__ inline
#endif
size_t predict_aux_doubles(int n_pad, int d, int ct) { return (size_t)d * n_pad + n_pad + (size_t)d * ct; }
inline size_t predict_smem_bytes(int n_pad, int d, int ct, int alias) {
  const size_t kst = (size_t)(n_pad / kBS) * (ct / kBS) * kPTile;
  const size_t stage = (size_t)kPStages * 2 * kPTile;
  const size_t aux = predict_aux_doubles(n_pad, d, ct);
  return sizeof(double) * (kst + stage + (alias ? 0 : aux) + (size_t)d * ct + 4 * ct + 2 * ct + 2 * kMaxP + 8);
}
// candidate tile width / layout for (n_pad, d): widest tile that fits 227 KB, un-aliased if possible
inline bool predict_config(int n_pad, int d, int* ct, int* alias) {
  for (int c = 64; c >= 32; c -= 32)
    for (int a = 0; a <= 1; ++a) {
      if (a && predict_aux_doubles(n_pad, d, c) > (size_t)kPStages * 2 * kPTile) continue;
      if (predict_smem_bytes(n_pad, d, c, a) <= 227 * 1024) {
        *ct = c, *alias = a;
        return true;
      }
    }
  return false;
}
inline int predict_nsplit(int M, int B, int num_sms, int ct) {
  const int ntile = (B + ct - 1) / ct;
  int ns = (2 * num_sms + ntile - 1) / ntile;
  if (ns < 1) ns = 1;
  if (ns > M) ns = M;
  return ns;
}
inline size_t predict_workspace_bytes(int M, int n_pad, int d, int B, int num_sms) {
  int ct = 64, alias = 0;
  if (!predict_config(n_pad, d, &ct, &alias)) return 0;
  const int ns = predict_nsplit(M, B, num_sms, ct);
  return ns > 1 ? sizeof(double) * 2 * (size_t)ns * (size_t)B : 0;
}

// one dense 32x32 tile (8 KB, contiguous) global -> padded shared rows: 2 x 16 B per thread (256 threads)
SCAML_DEVICE void ptile_async(double* sdst, const double* gsrc, int tid) {
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int c2 = tid + u * kPredThreads;
    const int row = c2 >> 4, j = c2 & 15;
    cp_async16(sdst + row * kPLd + 2 * j, gsrc + row * kBS + 2 * j);
  }
}

// flat chunk q of a task: super-row I, 32-wide column block ck (0 .. 2I+1)
struct PChunk {
  int I, ck;
  SCAML_DEVICE void next() {
    if (++ck > 2 * I + 1) {
      ++I;
      ck = 0;
    }
  }
};
SCAML_DEVICE void pred_issue(const double* Lm, const PChunk& c, double* st, int tid) {
  if (c.ck <= 2 * c.I) ptile_async(st, Lm + (size_t)(tri(2 * c.I) + c.ck) * kTile, tid);
  ptile_async(st + kPTile, Lm + (size_t)(tri(2 * c.I + 1) + c.ck) * kTile, tid);
}

template <int KIND, int CT, bool CROSS>
__global__ void __launch_bounds__(kPredThreads, 1) scaml_predict_kernel(const PredParams p) {
  constexpr int NJ = CT / 32;  // 8-column DMMA tiles per warp (warp tile: 32 rows x 8*NJ candidates)
  constexpr int CBT = CT / 32;  // kst tile columns
  SCAML_DYN_SMEM(double, sm);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  const int rb = warp >> 2, col0 = (warp & 3) * 8 * NJ, cb = col0 >> 5, cin = col0 & 31;
  const int d = p.d, P = d + 2, n_pad = p.n_pad;
  double* kst = sm;                                                // (n_pad/32) x CBT padded tiles [a][c]
  double* stage = kst + (size_t)(n_pad / kBS) * CBT * kPTile;      // kPStages x 2 padded tiles
  double* aux = stage + (p.alias ? 0 : kPStages * 2 * kPTile);
  double* xst = aux;                                               // [d][n_pad] scaled inputs of the task
  double* alp = xst + (size_t)d * n_pad;                           // [n_pad]
  double* xcs = alp + n_pad;                                       // [d][CT] scaled candidates
  double* xcr = stage + kPStages * 2 * kPTile + (p.alias ? 0 : predict_aux_doubles(n_pad, d, CT));  // [d][CT] raw
  double* red = xcr + d * CT;                                      // 4 x CT mean partials
  double* vsq = red + 4 * CT;                                      // 2 x CT
  double* invl2 = vsq + 2 * CT;                                    // 2 x kMaxP reciprocal lengthscales (task parity)
  int tcount = 0;
  const long long lstride = (long long)tri(n_pad / kBS) * kTile;
  const int items = p.ntile * p.nsplit;
  const int mper = (p.M + p.nsplit - 1) / p.nsplit;

  for (int it = blockIdx.x; it < items; it += gridDim.x) {
    const int ct = it % p.ntile, spx = it / p.ntile;
    const int b0 = ct * CT;
    const int m_lo = spx * mper, m_hi = (m_lo + mper < p.M) ? m_lo + mper : p.M;
    __syncthreads();
    for (int i = tid; i < CT * d; i += kPredThreads) {
      const int c = i / d, k = i - c * d;
      xcr[k * CT + c] = (b0 + c < p.B) ? p.Xc[(size_t)(b0 + c) * d + k] : 0.0;
    }
    double macc = 0.0, vacc = 0.0;
    // CROSS: cx[i][jj] = 8x8 block (candidates 32 xt + 8 i + .., target columns 8 (jw + JS jj) + ..) of
    // sum_m c_m k*_m^T A_m, accumulated over ALL tasks of the split by the tensor cores.  CT = 64: the warp's
    // candidate tile column xt = warp & 1 and column blocks jw = warp >> 1, +4, ..; CT = 32: xt = 0, jw = warp, +8, ..
    constexpr int JS = (CT == 64) ? 4 : 8;
    const int xt = (CT == 64) ? (warp & 1) : 0, jw = (CT == 64) ? (warp >> 1) : warp;
    double cx[CROSS ? 4 : 1][CROSS ? 4 : 1][2];
#pragma unroll
    for (int i = 0; i < (CROSS ? 4 : 1); ++i)
#pragma unroll
      for (int jj = 0; jj < (CROSS ? 4 : 1); ++jj) cx[i][jj][0] = cx[i][jj][1] = 0.0;
    // inputs of the NEXT task (point `tid`, first kPreD dimensions, alpha) are fetched into registers while the
    // tensor-core phase of the current task runs, so that the staging below does not wait on global memory
    constexpr int kPreD = 8;
    int pre_m = -1;
    double xpre[kPreD], apre = 0.0;
#pragma unroll
    for (int k = 0; k < kPreD; ++k) xpre[k] = 0.0;
    for (int m = m_lo; m < m_hi; ++m) {
      const double wm = p.w[m];
      if (wm == 0.0) continue;
      const int nv = p.n_valid ? p.n_valid[m] : p.n_max;
      const int NS = (nv + kSB - 1) / kSB, npt = NS * kSB;
      const int L = NS * (NS + 1);  // flat chunks of this task
      const double* th = p.theta + (size_t)m * P;
      const double os = th[d];
      const double* Lm = p.linv + (size_t)m * lstride;
      // x * (1/l) instead of x / l: an FP64 division is ~30 instructions per staged coordinate (<= 1 ulp apart)
      double* invl = invl2 + (tcount & 1) * kMaxP;
      ++tcount;
      if (tid < d) invl[tid] = 1.0 / th[tid];
      __syncthreads();  // previous task fully consumed (kst, stage, aux); invl visible
      PChunk qi{0, 0};
      int issued = 0;
      if (!p.alias) {  // start streaming L^-1 under the exp-heavy assembly below
        for (; issued < 2 && issued < L; ++issued, qi.next()) {
          pred_issue(Lm, qi, stage + (issued % kPStages) * 2 * kPTile, tid);
          cp_async_commit();
        }
      }
      {
        const double* Xm = p.X + (size_t)m * p.n_max * d;
        const bool have = (pre_m == m);
        if (tid < npt) {  // one point per thread: no runtime integer division
#pragma unroll
          for (int k = 0; k < kPreD; ++k)
            if (k < d) {
              const double v = have ? xpre[k] : ((tid < nv) ? Xm[(size_t)tid * d + k] : 0.0);
              xst[k * n_pad + tid] = (tid < nv) ? v * invl[k] : 0.0;
            }
          for (int k = kPreD; k < d; ++k) xst[k * n_pad + tid] = (tid < nv) ? Xm[(size_t)tid * d + k] * invl[k] : 0.0;
          alp[tid] = have ? apre : p.alpha[(size_t)m * n_pad + tid];
        }
        for (int a = tid + kPredThreads; a < npt; a += kPredThreads) {
          for (int k = 0; k < d; ++k) xst[k * n_pad + a] = (a < nv) ? Xm[(size_t)a * d + k] * invl[k] : 0.0;
          alp[a] = p.alpha[(size_t)m * n_pad + a];
        }
        for (int i = tid; i < CT * d; i += kPredThreads) xcs[i] = xcr[i] * invl[i / CT];
      }
      __syncthreads();
      // ---- k*(X_m, candidates) and the mean partials -------------------------------- //
      // a thread owns one candidate and walks the task's points 8 at a time: 8 independent distance / exp
      // chains per thread keep the FP64 pipe busy (a single chain per thread is latency bound).
      {
        constexpr int QN = kPredThreads / CT;  // row phases (4 for CT = 64, 8 for CT = 32)
        // rows per pass = QN * U must divide 64 (npt is a multiple of 64); CROSS keeps 32 more accumulators live
        constexpr int U = (CT == 64 && !CROSS) ? 16 : 8;
        static_assert((kPredThreads / CT) * U <= 64, "assembly pass must not run past the task's rows");
        const int c = tid % CT, q = tid / CT;
        double mu = 0.0;
        for (int a0 = q; a0 < npt; a0 += QN * U) {
          double r2[U];
#pragma unroll
          for (int u = 0; u < U; ++u) r2[u] = 0.0;
          for (int k = 0; k < d; ++k) {
            const double xc = xcs[k * CT + c];
            const double* xr = xst + k * n_pad + a0;
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const double df = xr[u * QN] - xc;
              r2[u] = fma(df, df, r2[u]);
            }
          }
          double kap[U];
          kappa_n<KIND, U, false>(r2, kap, kap);
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int a = a0 + u * QN;
            const double kv = ((a < nv) ? os : 0.0) * kap[u];
            kst[((a >> 5) * CBT + (c >> 5)) * kPTile + (a & 31) * kPLd + (c & 31)] = kv;
            mu = fma(kv, alp[a], mu);
          }
        }
        if (q < 4) red[q * CT + c] = mu;
        if (QN > 4) {  // CT = 32: fold phases 4..7 onto 0..3 in a fixed order
          __syncthreads();
          if (q >= 4) red[(q - 4) * CT + c] += mu;
        }
      }
      __syncthreads();
      if (CROSS) {
        // cx += (c_m k*)^T A_m : A-operand = k* (shared, this warp's 32-candidate tile column), B-operand = A_m
        // straight from L2 (each 4 x 8 fragment is 4 full 64-byte row segments), one k-step prefetched ahead
        const int njt = p.n_tp >> 3;
        const double cm = wm * wm * p.ystd[m] * p.ystd[m];
        const double* Am = p.condA + (size_t)m * n_pad * p.n_tp + (size_t)t4 * p.n_tp + g;
        const double* ks = kst + (size_t)xt * kPTile + t4 * kPLd + g;
        // work items = (32-row block rbk of k*, column block jj of this warp); the 8 B-fragments of an item are
        // fetched while the previous item's 32 DMMAs run (double-buffered registers): L2 latency is hidden
        const int nrb = npt >> 5;
        const int njw = (njt > jw) ? (njt - jw + JS - 1) / JS : 0;  // column blocks jw, jw+JS, .. owned by this warp
        const int nitems = nrb * njw;
        double bcur[8], bnxt[8];
        auto fetch = [&](int item, double (&b)[8]) {
          const int rbk = item / njw, jj = item - rbk * njw;
          const double* src = Am + (size_t)(32 * rbk) * p.n_tp + 8 * (jw + JS * jj);
#pragma unroll
          for (int s = 0; s < 8; ++s) b[s] = __ldg(src + (size_t)(4 * s) * p.n_tp);
        };
        if (nitems > 0) fetch(0, bnxt);
        for (int item = 0; item < nitems; ++item) {
#pragma unroll
          for (int s = 0; s < 8; ++s) bcur[s] = bnxt[s];
          if (item + 1 < nitems) fetch(item + 1, bnxt);
          const int rbk = item / njw, jj = item - rbk * njw;
          const double* kr = ks + (size_t)rbk * CBT * kPTile;
#pragma unroll
          for (int s = 0; s < 8; ++s) {
            const double a[4] = {cm * kr[0], cm * kr[8], cm * kr[16], cm * kr[24]};
            kr += 4 * kPLd;
            // jj is warp-uniform but not a compile-time constant: select the accumulator set by unrolled compare
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (q == jj) {
#pragma unroll
                for (int i = 0; i < 4; ++i) dmma884(cx[i][q], a[i], bcur[s]);
              }
          }
        }
      }
      if (p.alias) {
        for (; issued < 2 && issued < L; ++issued, qi.next()) {
          pred_issue(Lm, qi, stage + (issued % kPStages) * 2 * kPTile, tid);
          cp_async_commit();
        }
      }
      if (m + 1 < m_hi) {  // prefetch for the next task (lands during the product below)
        pre_m = m + 1;
        const int nv1 = p.n_valid ? p.n_valid[m + 1] : p.n_max;
        const double* X1 = p.X + (size_t)(m + 1) * p.n_max * d;
#pragma unroll
        for (int k = 0; k < kPreD; ++k) xpre[k] = (k < d && tid < nv1) ? X1[(size_t)tid * d + k] : 0.0;
        apre = (tid < n_pad) ? p.alpha[(size_t)(m + 1) * n_pad + tid] : 0.0;
      }
      // ---- V = L^-1 k*  on the FP64 tensor cores, column sums of squares ------------- //
      double acc[4][NJ][2];
      double vs[NJ][2];
#pragma unroll
      for (int j = 0; j < NJ; ++j) vs[j][0] = vs[j][1] = 0.0;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
      PChunk qc{0, 0};
      for (int q = 0; q < L; ++q) {
        cp_async_wait<1>();
        __syncthreads();  // chunk q visible to everyone; everyone has finished reading chunk q-1
        if (issued < L) {
          pred_issue(Lm, qi, stage + (issued % kPStages) * 2 * kPTile, tid);
          ++issued;
          qi.next();
        }
        cp_async_commit();  // always commit (possibly empty): keeps the wait<1> accounting uniform
        const int brow = 2 * qc.I + rb;  // 32-row block of L^-1 this warp multiplies
        if (qc.ck <= brow) {
          const bool dg = (qc.ck == brow);  // diagonal tile: L^-1(r, kk) = 0 for kk > r
          const double* ar = stage + ((q % kPStages) * 2 + rb) * kPTile + t4 * kPLd + g;
          const double* br = kst + (qc.ck * CBT + cb) * kPTile + t4 * kPLd + cin + g;
#pragma unroll 2
          for (int s = 0; s < 8; ++s) {
            const double a[4] = {ar[0], ar[8], ar[16], ar[24]};
            double b[NJ];
#pragma unroll
            for (int j = 0; j < NJ; ++j) b[j] = br[8 * j];
            ar += 4 * kPLd;
            br += 4 * kPLd;
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (!dg || s < 2 * i + 2) {
#pragma unroll
                for (int j = 0; j < NJ; ++j) dmma884(acc[i][j], a[i], b[j]);
              }
          }
        }
        if (qc.ck == 2 * qc.I + 1) {  // super-row complete: ||V_I||^2 per candidate column
#pragma unroll
          for (int j = 0; j < NJ; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                vs[j][e] = fma(acc[i][j][e], acc[i][j][e], vs[j][e]);
                acc[i][j][e] = 0.0;
              }
            }
        }
        qc.next();
      }
      cp_async_wait<0>();
#pragma unroll
      for (int j = 0; j < NJ; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          double s = vs[j][e];
          s += __shfl_xor_sync(0xffffffffu, s, 4);
          s += __shfl_xor_sync(0xffffffffu, s, 8);
          s += __shfl_xor_sync(0xffffffffu, s, 16);
          if (g == 0) vsq[rb * CT + col0 + 8 * j + 2 * t4 + e] = s;
        }
      __syncthreads();
      if (tid < CT) {
        const double mu = ((red[tid] + red[CT + tid]) + red[2 * CT + tid]) + red[3 * CT + tid];
        const double ssq = vsq[tid] + vsq[CT + tid];
        const double ys = p.ystd[m];
        macc = fma(wm, p.ybar[m] + ys * mu, macc);
        vacc = fma(wm * wm * ys * ys, os - ssq, vacc);
      }
    }
    if (CROSS) {
      const int njt = p.n_tp >> 3;
      double* dst = p.cxp + (size_t)spx * p.B * p.n_tp;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int b = b0 + 32 * xt + 8 * i + g;
        if (b < p.B) {
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
            if (jw + JS * jj < njt)
              *reinterpret_cast<double2*>(dst + (size_t)b * p.n_tp + 8 * (jw + JS * jj) + 2 * t4) =
                  make_double2(cx[i][jj][0], cx[i][jj][1]);
        }
      }
    }
    if (tid < CT && b0 + tid < p.B) {
      if (p.nsplit == 1) {
        p.mean[b0 + tid] = macc;
        p.var[b0 + tid] = vacc;
      } else {
        p.part[((size_t)spx * 2 + 0) * p.B + b0 + tid] = macc;
        p.part[((size_t)spx * 2 + 1) * p.B + b0 + tid] = vacc;
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

// many splits (few candidates): one warp per candidate, lanes stride over the splits, fixed-order warp tree
__global__ void __launch_bounds__(128) scaml_predict_reduce_wide_kernel(const double* part, double* mean, double* var,
                                                                        int nsplit, int B) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpg = blockDim.x >> 5;
  for (int b = blockIdx.x * wpg + warp; b < B; b += gridDim.x * wpg) {
    double m = 0.0, v = 0.0;
    for (int s = lane; s < nsplit; s += 32) {
      m += part[((size_t)s * 2 + 0) * B + b];
      v += part[((size_t)s * 2 + 1) * B + b];
    }
    m = warp_sum(m);
    v = warp_sum(v);
    if (lane == 0) {
      mean[b] = m;
      var[b] = v;
    }
  }
}

template <int KIND, int CT, bool CROSS>
int launch_predict_kc(const PredParams& p, int grid, size_t smem, void* stream) {
#ifdef SCAML_EMU
  (void)stream;
  cuemu::launch(dim3(grid), dim3(kPredThreads), smem, scaml_predict_kernel<KIND, CT, CROSS>, p);
  return 0;
#else
  cudaError_t err = cudaFuncSetAttribute(scaml_predict_kernel<KIND, CT, CROSS>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return (int)err;
  scaml_predict_kernel<KIND, CT, CROSS><<<grid, kPredThreads, smem, (cudaStream_t)stream>>>(p);
  return (int)cudaGetLastError();
#endif
}
template <int KIND>
int launch_predict_k(const PredParams& p, int grid, size_t smem, void* stream) {
  if (p.condA != nullptr)
    return p.ct == 64 ? launch_predict_kc<KIND, 64, true>(p, grid, smem, stream)
                      : launch_predict_kc<KIND, 32, true>(p, grid, smem, stream);
  return p.ct == 64 ? launch_predict_kc<KIND, 64, false>(p, grid, smem, stream)
                    : launch_predict_kc<KIND, 32, false>(p, grid, smem, stream);
}

// condA != null: also accumulate the cross-covariance partials into cxp (needs 64-candidate tiles: SCAML_E_SMEM
// otherwise); workspace layout: [nsplit][2][B] mean/var partials (nsplit > 1) -- cxp is a separate buffer
inline int launch_predict_weighted(const double* X, const int32_t* n_valid, const double* theta, const double* linv,
                                   const double* alpha, const double* ybar, const double* ystd, const double* w,
                                   const double* Xc, double* mean, double* var, double* workspace, int M, int n_max,
                                   int n_pad, int d, int B, int kernel, int num_sms, void* stream,
                                   const double* condA = nullptr, double* cxp = nullptr, int n_tp = 0) {
  PredParams p;
  p.condA = condA, p.cxp = cxp, p.n_tp = n_tp;
  p.X = X, p.n_valid = n_valid, p.theta = theta, p.linv = linv, p.alpha = alpha, p.ybar = ybar, p.ystd = ystd;
  p.w = w, p.Xc = Xc, p.mean = mean, p.var = var, p.part = workspace;
  p.M = M, p.n_max = n_max, p.n_pad = n_pad, p.d = d, p.B = B;
  if (!predict_config(n_pad, d, &p.ct, &p.alias)) return SCAML_E_SMEM;
  if (condA != nullptr && (n_tp <= 0 || n_tp > 128 || (n_tp & 7))) return SCAML_E_UNSUPPORTED;
  p.ntile = (B + p.ct - 1) / p.ct;
  p.nsplit = predict_nsplit(M, B, num_sms, p.ct);
  const size_t smem = predict_smem_bytes(n_pad, d, p.ct, p.alias);
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
  const bool wide = p.nsplit >= 32;  // by shape only
#ifdef SCAML_EMU
  if (wide) cuemu::launch(dim3(1), dim3(128), 0, scaml_predict_reduce_wide_kernel, (const double*)workspace, mean, var, p.nsplit, B);
  else cuemu::launch(dim3(1), dim3(64), 0, scaml_predict_reduce_kernel, (const double*)workspace, mean, var, p.nsplit, B);
  return 0;
#else
  if (wide) {
    const int wb = (B + 3) / 4;
    scaml_predict_reduce_wide_kernel<<<wb < 1184 ? wb : 1184, 128, 0, (cudaStream_t)stream>>>(workspace, mean, var,
                                                                                             p.nsplit, B);
  } else {
    const int rb = (B + 255) / 256;
    scaml_predict_reduce_kernel<<<rb < 1184 ? rb : 1184, 256, 0, (cudaStream_t)stream>>>(workspace, mean, var,
                                                                                        p.nsplit, B);
  }
  return (int)cudaGetLastError();
#endif
}

}  // namespace scaml
