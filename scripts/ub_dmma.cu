// Exploration microbenchmark (GPU box): what DMMA rate can N warps per SM sustain, with register-resident or
// shared-memory operands, and do DMMA and DFMA share a pipe?   nvcc -arch=sm_100a -O3 -o scripts/ub_dmma.bin scripts/ub_dmma.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma1684(double (&d)[4], const double (&a)[2], double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}

constexpr int kLd = 36;
// mode 0: m8n8k4 4x4 tiles, register operands.  1: m8n8k4 4x4 tiles, smem operands (fmma-like, 32x32 warp tile)
// 2: m16n8k8, 2x4 tiles (32x32 warp tile), smem operands.  3: m8n8k4 2x4 tiles (16x32 warp tile) smem operands
// 4: DFMA chains (16 independent).  5: warps alternate DMMA(mode 1) / DFMA by warp parity
// 6: m16n8k4 2x4 tiles smem operands
template <int MODE>
__global__ void k(double* out, int iters, int dmma_warps) {
  extern __shared__ double sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t4 = lane & 3;
  for (int i = tid; i < 2 * 32 * kLd * 2; i += blockDim.x) sm[i] = 1e-3 * (i % 7);
  __syncthreads();
  const double* A = sm + (warp & 1) * 32 * kLd;
  const double* B = sm + 2 * 32 * kLd;
  double s = 0;
  if (MODE == 4 || (MODE == 5 && warp >= dmma_warps)) {
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = 1.0 + tid * 1e-9 + i;
    for (int it = 0; it < iters * 8; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fma(a[i], 1.0000001, 1e-9);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
  } else if (MODE == 0) {
    double acc[4][4][2] = {};
    double a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = 1.0 + lane * 1e-6 + i, b[i] = 1e-3 * (i + 1);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int s4 = 0; s4 < 8; ++s4)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma884(acc[i][j], a[i], b[j]);
    }
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j][0] + acc[i][j][1];
  } else if (MODE == 1 || MODE == 5) {
    double acc[4][4][2] = {};
    for (int it = 0; it < iters; ++it) {
      const double* ar = A + t4 * kLd + g;
      const double* br = B + t4 * kLd + g;
#pragma unroll 2
      for (int s4 = 0; s4 < 8; ++s4) {
        const double a[4] = {ar[0], ar[8], ar[16], ar[24]};
        const double b[4] = {br[0], br[8], br[16], br[24]};
        ar += 4 * kLd; br += 4 * kLd;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma884(acc[i][j], a[i], b[j]);
      }
    }
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j][0] + acc[i][j][1];
  } else if (MODE == 7 || MODE == 8 || MODE == 9) {
    // fit-kernel-like step: [cp.async 16 KB (MODE 8/9)] barrier, 4 k-steps (64 DMMA), barrier
    double acc[4][4][2] = {};
    double* stg = sm + 4 * 32 * kLd;  // 2 stages x 4 half tiles x 576 doubles
    for (int it = 0; it < iters * 2; ++it) {
      if (MODE >= 8) {
        double* dst = stg + (it & 1) * 4 * 576;
        const double* gsrc = out + 4096 + (size_t)blockIdx.x * 8192 + (it & 3) * 2048;
#pragma unroll
        for (int h = 0; h < 4; ++h)
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int c2 = tid + u * 128, row = c2 >> 4, j = c2 & 15;
            unsigned sa = (unsigned)__cvta_generic_to_shared(dst + h * 576 + row * kLd + 2 * j);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gsrc + h * 512 + row * 32 + 2 * j) : "memory");
          }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        asm volatile("cp.async.wait_group 1;\n" ::: "memory");
      }
      __syncthreads();
      const double* ar = (MODE == 9 ? stg + ((it + 1) & 1) * 4 * 576 + (warp & 1) * 576 : A) + t4 * kLd + g;
      const double* br = (MODE == 9 ? stg + ((it + 1) & 1) * 4 * 576 + (2 + (warp >> 1)) * 576 : B) + t4 * kLd + g;
#pragma unroll 2
      for (int s4 = 0; s4 < 4; ++s4) {
        const double a[4] = {ar[0], ar[8], ar[16], ar[24]};
        const double b[4] = {br[0], br[8], br[16], br[24]};
        ar += 4 * kLd; br += 4 * kLd;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma884(acc[i][j], a[i], b[j]);
      }
      __syncthreads();
    }
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j][0] + acc[i][j][1];
  } else if (MODE == 2) {
    // m16n8k8: A frag a0:(g, t4) a1:(g+8, t4) a2:(g, t4+4) a3:(g+8, t4+4); B frag b0:(t4, g) b1:(t4+4, g)
    double acc[2][4][4] = {};
    for (int it = 0; it < iters; ++it) {
      const double* ar = A + t4 * kLd + g;
      const double* br = B + t4 * kLd + g;
#pragma unroll 2
      for (int s8 = 0; s8 < 4; ++s8) {
        double a[2][4], b[4][2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          a[i][0] = ar[16 * i]; a[i][1] = ar[16 * i + 8]; a[i][2] = ar[4 * kLd + 16 * i]; a[i][3] = ar[4 * kLd + 16 * i + 8];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) { b[j][0] = br[8 * j]; b[j][1] = br[4 * kLd + 8 * j]; }
        ar += 8 * kLd; br += 8 * kLd;
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma1688(acc[i][j], a[i], b[j]);
      }
    }
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 4; ++j) for (int e = 0; e < 4; ++e) s += acc[i][j][e];
  } else if (MODE == 3) {
    double acc[2][4][2] = {};
    for (int it = 0; it < iters; ++it) {
      const double* ar = A + t4 * kLd + g;
      const double* br = B + t4 * kLd + g;
#pragma unroll 2
      for (int s4 = 0; s4 < 8; ++s4) {
        const double a[2] = {ar[0], ar[8]};
        const double b[4] = {br[0], br[8], br[16], br[24]};
        ar += 4 * kLd; br += 4 * kLd;
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma884(acc[i][j], a[i], b[j]);
      }
    }
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j][0] + acc[i][j][1];
  } else if (MODE == 6) {
    double acc[2][4][4] = {};
    for (int it = 0; it < iters; ++it) {
      const double* ar = A + t4 * kLd + g;
      const double* br = B + t4 * kLd + g;
#pragma unroll 2
      for (int s4 = 0; s4 < 8; ++s4) {
        double a[2][2], b[4];
#pragma unroll
        for (int i = 0; i < 2; ++i) { a[i][0] = ar[16 * i]; a[i][1] = ar[16 * i + 8]; }
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = br[8 * j];
        ar += 4 * kLd; br += 4 * kLd;
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma1684(acc[i][j], a[i], b[j]);
      }
    }
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 4; ++j) for (int e = 0; e < 4; ++e) s += acc[i][j][e];
  }
  if (s == 123.456) out[blockIdx.x * blockDim.x + tid] = s;
}

template <int MODE>
void run(const char* name, int warps, int iters, double flop_per_warp_iter, double* out, int dmma_warps = 0,
         double dfma_flop_per_warp_iter = 0) {
  const int smem = 120 * 1024;  // > 113 KB: one CTA per SM
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k<MODE><<<148, warps * 32, smem>>>(out, iters, dmma_warps);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  double fl;
  if (MODE == 5) fl = 148.0 * iters * (dmma_warps * flop_per_warp_iter + (warps - dmma_warps) * dfma_flop_per_warp_iter);
  else fl = 148.0 * warps * iters * flop_per_warp_iter;
  printf("%-44s warps/SM=%2d  %8.3f ms  %7.2f TFLOP/s", name, warps, best, fl / best / 1e9);
  if (MODE == 5) printf("  (dmma part %.2f, dfma part %.2f)", 148.0 * iters * dmma_warps * flop_per_warp_iter / best / 1e9,
                        148.0 * iters * (warps - dmma_warps) * dfma_flop_per_warp_iter / best / 1e9);
  printf("\n");
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) printf("  CUDA error %s\n", cudaGetErrorString(err));
}

int main() {
  double* out; cudaMalloc(&out, (4096 + 148 * 8192 + 8192) * 8); cudaMemset(out, 0, (4096 + 148 * 8192 + 8192) * 8);
  const int it = 2000;
  const double f884x16x8 = 8 * 16 * 512.0;  // per warp per iter: 8 k-steps x 16 DMMA x 512 flop
  for (int w : {4, 8, 12, 16, 32}) run<0>("m8n8k4 4x4 tiles, register operands", w, it, f884x16x8, out);
  for (int w : {4, 8, 12, 16, 32}) run<1>("m8n8k4 4x4 tiles, smem operands", w, it, f884x16x8, out);
  run<7>("fit-like step: bar, 64 DMMA (smem), bar", 4, it, f884x16x8, out);
  run<8>("fit-like step + cp.async 16KB/step (unused)", 4, it, f884x16x8, out);
  run<9>("fit-like step + cp.async 16KB/step (consumed)", 4, it, f884x16x8, out);
  for (int w : {4, 8, 12, 16, 32}) run<2>("m16n8k8 2x4 tiles, smem operands", w, it, 4 * 8 * 2048.0, out);
  for (int w : {4, 8, 12, 16, 32}) run<6>("m16n8k4 2x4 tiles, smem operands", w, it, 8 * 8 * 1024.0, out);
  for (int w : {4, 8, 12, 16, 32}) run<3>("m8n8k4 2x4 tiles (16x32), smem operands", w, it, 8 * 8 * 512.0, out);
  for (int w : {4, 8, 12, 16, 32}) run<4>("DFMA 16 chains", w, it, 8 * 16 * 64.0, out);
  for (int w : {8, 16, 32}) run<5>("half warps DMMA(smem) / half DFMA", w, it, f884x16x8, out, w / 2, 8 * 16 * 64.0);
  return 0;
}
