// cycles per tcgen05.mma for cta_group::2 at M=256 (128 rows/CTA) and M=128 (64 rows/CTA), lean issue loop.
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace wmrl::ptx;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k(int M, int N, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const uint32_t rank = cluster_ctarank();
  const uint32_t base = smem_u32(smem);
  for (int i = threadIdx.x; i < (160 * 1024) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc2(smem_u32(&tslot), 512); tmem_relinquish2(); }
  fence_proxy_async_smem();
  tc_fence_before(); cluster_sync_all(); tc_fence_after();
  const uint32_t tm = tslot;
  if (rank == 0 && threadIdx.x == 0) {
    const int rows = M / 2, nh = N / 2;
    const uint64_t ad0 = make_smem_desc(base, rows * 16, 128), bd0 = make_smem_desc(base + 65536, nh * 16, 128);
    const uint32_t ahi = (uint32_t)(ad0 >> 32), bhi = (uint32_t)(bd0 >> 32);
    const uint32_t alo = (uint32_t)ad0, blo = (uint32_t)bd0;
    const uint32_t astep = (rows * 32) >> 4, bstep = (nh * 32) >> 4;
    const uint32_t idesc = make_idesc_bf16(M, N);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        mma_bf16_lohi2(tm, alo + kk * astep, ahi, blo + kk * bstep, bhi, idesc, (it | kk) ? 1u : 0u);
    }
    tc_commit2(smem_u32(&bar), 3);
    long long t1 = clock64();
    mbar_wait(smem_u32(&bar), 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  if (rank == 1 && threadIdx.x == 0) mbar_wait(smem_u32(&bar), 0);
  tc_fence_before(); cluster_sync_all();
  if (threadIdx.x < 32) tmem_dealloc2(tm, 512);
}
int main() {
  long long* out; cudaMallocManaged(&out, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  for (int M : {256, 128})
    for (int N : {256, 128, 64, 32, 16}) {
      const int iters = 256;
      k<<<148, 128, 160 * 1024>>>(M, N, iters, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("M %d N %d error %s\n", M, N, cudaGetErrorString(e)); return 1; }
      printf("cta_group::2 M %3d N %3d: %6.1f cycles/MMA to complete (MACs/clk/SM %6.0f)\n", M, N, (double)out[1] / (iters * 4),
             (double)M * N * 16 / 2 / ((double)out[1] / (iters * 4)));
    }
}
