// FP64 roofline denominators measured on the box itself (MEASURED_PEAKS.json has no FP64
// entry): register-resident DFMA chains and DMMA (mma.sync f64) chains.  Diagnostics only.
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double seed) {
  double a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed + threadIdx.x * 1e-9 + i;
  const double b = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], b, c);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d[0]), "+d"(d[1])
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
      "{%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
      : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]),
        "d"(b[2]), "d"(b[3]));
}

__global__ void __launch_bounds__(256) dmma884_kernel(double* out, int iters, double seed) {
  double acc[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = seed + i;
  const double a = 1.0 + threadIdx.x * 1e-12, b = 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma884(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) dmma16816_kernel(double* out, int iters, double seed) {
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = seed + i + j;
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 1.0 + threadIdx.x * 1e-12 + i * 1e-6;
#pragma unroll
  for (int i = 0; i < 4; ++i) b[i] = 1e-3 + i * 1e-6;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) dmma16816(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += acc[i][j];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

extern "C" {
// mode 0: DFMA, 1: DMMA m8n8k4, 2: DMMA m16n8k16.  Launches `blocks` CTAs of 256 threads and
// stores the number of FLOPs one launch performs in *flops.  out: >= blocks*256 doubles.
int scaml_microbench_fp64(int mode, int blocks, int iters, double* out, double* flops, void* stream) {
  if (blocks <= 0 || iters <= 0 || !out || !flops) return -1;
  cudaStream_t s = (cudaStream_t)stream;
  const double threads = (double)blocks * 256.0, warps = threads / 32.0;
  if (mode == 0) {
    dfma_kernel<<<blocks, 256, 0, s>>>(out, iters, 1.0);
    *flops = threads * (double)iters * 16.0 * 2.0;
  } else if (mode == 1) {
    dmma884_kernel<<<blocks, 256, 0, s>>>(out, iters, 1.0);
    *flops = warps * (double)iters * 8.0 * 512.0;
  } else if (mode == 2) {
    dmma16816_kernel<<<blocks, 256, 0, s>>>(out, iters, 1.0);
    *flops = warps * (double)iters * 4.0 * 4096.0;
  } else {
    return -1;
  }
  return (int)cudaGetLastError();
}
}
