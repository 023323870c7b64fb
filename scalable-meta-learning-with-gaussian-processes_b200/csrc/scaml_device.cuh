// Device-side helpers shared by the ScaML-GP kernels (sm_100a; also compiles under the
// CPU logic emulation in emu/cuda_emu.h when SCAML_EMU is defined).
#pragma once

#ifdef SCAML_EMU
#include "emu/cuda_emu.h"
#define SCAML_DEVICE inline
#define SCAML_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(cuemu::ctx()->dyn_smem)
struct alignas(16) double2 {
  double x, y;
};
inline double2 make_double2(double x, double y) { return double2{x, y}; }
#else
#include <cuda_runtime.h>
#define SCAML_DEVICE __device__ __forceinline__
#define SCAML_DYN_SMEM(type, name) \
  extern __shared__ __align__(1024) unsigned char name##_raw_[]; \
  type* name = reinterpret_cast<type*>(name##_raw_)
#endif

#include <stdint.h>

#include "../../include/scaml_b200.h"

namespace scaml {

constexpr int kBS = 32;           // tile edge
constexpr int kTile = kBS * kBS;  // doubles per tile (8 KB)
constexpr int kSB = 64;           // super-tile edge (2x2 tiles)
constexpr int kMaxP = 34;  // d <= 32
constexpr double kLog2Pi = 1.8378770664093454835606594728112;

// ----------------------------------------------------------------------------------- //
// async copies (LDGSTS, L2-only caching: the workspace is rewritten in place, L1 must
// never serve a stale line)
// ----------------------------------------------------------------------------------- //
SCAML_DEVICE void cp_async16(double* smem_dst, const double* gsrc) {
#ifdef SCAML_EMU
  std::memcpy(smem_dst, gsrc, 16);
#else
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc) : "memory");
#endif
}
SCAML_DEVICE void cp_async_commit() {
#ifndef SCAML_EMU
  asm volatile("cp.async.commit_group;\n" ::: "memory");
#endif
}
template <int N>
SCAML_DEVICE void cp_async_wait() {
#ifndef SCAML_EMU
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
#endif
}

// FP64 tensor-core MMA (DMMA): D(8x8) += A(8x4) * B(4x8).  Fragments (PTX ISA, mma.m8n8k4 .f64):
//   a    = A[lane>>2][lane&3]        b = B[lane&3][lane>>2]
//   d[e] = D[lane>>2][2*(lane&3)+e]
SCAML_DEVICE void dmma884(double (&d)[2], double a, double b) {
#ifdef SCAML_EMU
  cuemu::dmma884(d, a, b);
#else
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d[0]), "+d"(d[1])
               : "d"(a), "d"(b));
#endif
}

SCAML_DEVICE double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

SCAML_DEVICE int tri(int i) { return (i * (i + 1)) >> 1; }

// ----------------------------------------------------------------------------------- //
// stationary kernels: kappa(r^2) and kd = -2 dkappa/dr^2
// ----------------------------------------------------------------------------------- //
template <int KIND>
SCAML_DEVICE double kappa_of(double r2) {
  if (KIND == SCAML_KERNEL_RBF) return exp(-0.5 * r2);
  const double r = sqrt(r2 < 1e-30 ? 1e-30 : r2);
  if (KIND == SCAML_KERNEL_MATERN12) return exp(-r);
  if (KIND == SCAML_KERNEL_MATERN32) {
    const double s3 = 1.7320508075688772935;
    return (1.0 + s3 * r) * exp(-s3 * r);
  }
  const double s5 = 2.2360679774997896964;
  return (1.0 + s5 * r + (5.0 / 3.0) * r * r) * exp(-s5 * r);
}

template <int KIND>
SCAML_DEVICE void kappa_pair(double r2, double& k, double& kd) {
  if (KIND == SCAML_KERNEL_RBF) {
    k = exp(-0.5 * r2);
    kd = k;
    return;
  }
  const double r = sqrt(r2 < 1e-30 ? 1e-30 : r2);
  if (KIND == SCAML_KERNEL_MATERN12) {
    k = exp(-r);
    kd = (r2 > 0.0) ? k / r : 0.0;
    return;
  }
  if (KIND == SCAML_KERNEL_MATERN32) {
    const double s3 = 1.7320508075688772935;
    const double e = exp(-s3 * r);
    k = (1.0 + s3 * r) * e;
    kd = 3.0 * e;
    return;
  }
  const double s5 = 2.2360679774997896964;
  const double e = exp(-s5 * r);
  k = (1.0 + s5 * r + (5.0 / 3.0) * r * r) * e;
  kd = (5.0 / 3.0) * (1.0 + s5 * r) * e;
}

// runtime-dispatched variant for the non-hot callers
SCAML_DEVICE double kappa_rt(int kind, double r2) {
  switch (kind) {
    case SCAML_KERNEL_RBF: return kappa_of<SCAML_KERNEL_RBF>(r2);
    case SCAML_KERNEL_MATERN12: return kappa_of<SCAML_KERNEL_MATERN12>(r2);
    case SCAML_KERNEL_MATERN32: return kappa_of<SCAML_KERNEL_MATERN32>(r2);
    default: return kappa_of<SCAML_KERNEL_MATERN52>(r2);
  }
}

// ----------------------------------------------------------------------------------- //
// priors (torch.distributions Gamma / LogNormal log densities) and the Interval transform
// ----------------------------------------------------------------------------------- //
SCAML_DEVICE double log_prior(int kind, double p1, double p2, double x) {
  if (kind == SCAML_PRIOR_GAMMA) return p1 * log(p2) + (p1 - 1.0) * log(x) - p2 * x - lgamma(p1);
  if (kind == SCAML_PRIOR_LOGNORMAL) {
    const double lx = log(x);
    const double t = lx - p1;
    return -lx - log(p2) - 0.5 * kLog2Pi - t * t / (2.0 * p2 * p2);
  }
  return 0.0;
}
SCAML_DEVICE double dlog_prior(int kind, double p1, double p2, double x) {
  if (kind == SCAML_PRIOR_GAMMA) return (p1 - 1.0) / x - p2;
  if (kind == SCAML_PRIOR_LOGNORMAL) return -1.0 / x - (log(x) - p1) / (p2 * p2 * x);
  return 0.0;
}
SCAML_DEVICE double sigmoid(double x) { return 1.0 / (1.0 + exp(-x)); }

}  // namespace scaml
