// Which part of {expect_tx, cp.async.bulk, wait} costs ~640 cycles for the issuing thread?  And can
// several lanes of ONE warp issue copies in parallel?
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace wmrl::ptx;

__global__ void __launch_bounds__(128, 1) k(const uint8_t* src, int bytes, int nlanes, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[32];
  const uint32_t base = smem_u32(smem);
  if (threadIdx.x == 0) { for (int i = 0; i < 32; ++i) mbar_init(smem_u32(&bars[i]), 1); fence_barrier_init(); }
  __syncthreads();
  const int lane = threadIdx.x;
  if (threadIdx.x < 32) {
    long long t_issue = 0, t_wait = 0, t_all0 = clock64();
    const int iters = 256;
    if (lane < nlanes) {
      uint32_t ph = 0;
      const uint32_t bar = smem_u32(&bars[lane]);
      for (int it = 0; it < iters; ++it) {
        long long a = clock64();
        mbar_arrive_expect_tx(bar, bytes);
        bulk_g2s(base + lane * 16384, src + ((size_t)(it * nlanes + lane) * bytes) % (1 << 20), bytes, bar);
        long long b = clock64();
        mbar_wait(bar, ph); ph ^= 1;
        long long c = clock64();
        t_issue += b - a; t_wait += c - b;
      }
    }
    long long t_all1 = clock64();
    if (blockIdx.x == 0 && lane == 0) { out[0] = t_issue / iters; out[1] = t_wait / iters; out[2] = (t_all1 - t_all0) / iters; }
  }
}
int main() {
  uint8_t* src; cudaMalloc(&src, 2 << 20); cudaMemset(src, 1, 2 << 20);
  long long* out; cudaMallocManaged(&out, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int bytes : {4096, 16384})
    for (int nlanes : {1, 2, 4, 8}) {
      if (nlanes * 16384 > 200 * 1024) continue;
      k<<<148, 128, 200 * 1024>>>(src, bytes, nlanes, out);
      k<<<148, 128, 200 * 1024>>>(src, bytes, nlanes, out);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("err\n"); return 1; }
      printf("bytes %5d lanes %d: issue %4lld cyc, wait %4lld cyc, per-iteration %4lld cyc -> %5.1f B/clk/SM\n", bytes, nlanes,
             out[0], out[1], out[2], (double)bytes * nlanes / out[2]);
    }
}
