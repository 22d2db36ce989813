// LSU global-load throughput per SM vs. the shape of one warp-wide LDG.128 (512 B per instruction):
// R rows x (512/R) contiguous bytes, rows `ld` floats apart.  256 threads per CTA, 148 CTAs, each CTA walks its own
// 128-row panel of an L2-resident matrix along k; 8 loads in flight per thread.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256, 1) k(const float* __restrict__ A, int ld, int rows_per_instr, int iters, float* sink,
                                           long long* out) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int R = rows_per_instr, Q = 32 / R;          // Q 16-byte quads per row per instruction
  const int row = (lane / Q), quad = lane % Q;
  // a CTA covers 128 rows: warp w takes rows [16w, 16w+16) in passes of R rows
  const float* base = A + (size_t)(blockIdx.x % 16) * 128 * ld;
  float acc = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const int k0 = (it * Q * 4) % (ld / 2);
    float4 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = (warp * 32 + j * R + row) % 128;
      v[j] = *reinterpret_cast<const float4*>(base + (size_t)r * ld + k0 + quad * 4);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += v[j].x + v[j].y + v[j].z + v[j].w;
  }
  long long t1 = clock64();
  if (acc == 123.456f) sink[0] = acc;
  if (blockIdx.x == 0 && tid == 0) out[0] = t1 - t0;
}
int main() {
  const int ld = 2048, M = 16 * 128;     // 16 MB matrix: L2 resident
  float* A; cudaMalloc(&A, (size_t)M * ld * 4); cudaMemset(A, 0, (size_t)M * ld * 4);
  float* sink; cudaMalloc(&sink, 4);
  long long* out; cudaMallocManaged(&out, 8);
  const int iters = 400;
  for (int R : {1, 2, 4, 8, 16, 32}) {
    for (int rep = 0; rep < 2; ++rep) k<<<148, 256>>>(A, ld, R, iters, sink, out);
    { cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; } }
    const double bytes = 256.0 * 16 * 8 * iters;
    printf("%2d rows x %3d B per warp LDG.128: %6.1f B/clk/SM\n", R, 512 / R, bytes / out[0]);
  }
  return 0;
}
