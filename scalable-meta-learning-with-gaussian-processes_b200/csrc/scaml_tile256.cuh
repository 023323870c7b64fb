// 256-thread / 4x4-per-thread register tiling of a 64x64 super-tile (prediction and
// kernel-matrix kernels).  acc[4][4] += A[kk][r] * B[kk][c] over 32-deep chunks of
// k-major 32x32 tiles held in shared memory.
#pragma once
#include "scaml_device.cuh"

namespace scaml {

// per-thread coordinates inside a 64x64 super-tile (16x16 threads of 4x4 elements)
struct Thr {
  int tid, warp, lane;
  int rb, cb;    // tile row / col inside the super-tile (0/1) -- warp-uniform
  int rin, cin;  // first row / col inside the tile (multiples of 4)
};
SCAML_DEVICE Thr make_thr() {
  Thr t;
  t.tid = threadIdx.x;
  t.warp = t.tid >> 5;
  t.lane = t.tid & 31;
  const int tyb = t.warp >> 2, txb = t.warp & 3;
  t.rb = tyb;
  t.cb = txb >> 1;
  t.rin = 4 * (t.lane >> 2);
  t.cin = 16 * (txb & 1) + 4 * (t.lane & 3);
  return t;
}

SCAML_DEVICE void acc_zero(double (&acc)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
}

// acc += A[kk][rin..rin+3] (x) B[kk][cin..cin+3] over one 32-deep chunk
SCAML_DEVICE void mma_chunk(double (&acc)[4][4], const double* __restrict__ Ap, const double* __restrict__ Bp) {
#pragma unroll 8
  for (int kk = 0; kk < kBS; ++kk) {
    const double2 a01 = *reinterpret_cast<const double2*>(Ap + kk * kBS);
    const double2 a23 = *reinterpret_cast<const double2*>(Ap + kk * kBS + 2);
    const double2 b01 = *reinterpret_cast<const double2*>(Bp + kk * kBS);
    const double2 b23 = *reinterpret_cast<const double2*>(Bp + kk * kBS + 2);
    const double a[4] = {a01.x, a01.y, a23.x, a23.y};
    const double b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
  }
}

// One 32x32 tile (8 KB, contiguous) global -> shared by a 256-thread CTA: 2 x 16 B per thread.
SCAML_DEVICE void tile_async256(double* sdst, const double* gsrc, int tid) {
  cp_async16(sdst + 2 * tid, gsrc + 2 * tid);
  cp_async16(sdst + 2 * (tid + 256), gsrc + 2 * (tid + 256));
}

}  // namespace scaml
