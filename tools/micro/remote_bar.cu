// Can cp.async.bulk (global -> own smem) complete_tx on an mbarrier that lives in the PEER CTA?
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace wmrl::ptx;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k(const uint32_t* src, int* ok, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  cluster_sync_all();
  const int bytes = 8192;
  long long t0 = clock64();
  for (int it = 0; it < 64; ++it) {
    if (threadIdx.x == 0) {
      if (rank == 0) {
        mbar_arrive_expect_tx(smem_u32(&bar), 2 * bytes);                 // both halves land on MY barrier
        bulk_g2s(smem_u32(smem), src + it * 4096, bytes, smem_u32(&bar));
      } else {
        bulk_g2s(smem_u32(smem), src + it * 4096 + 2048, bytes, mapa(smem_u32(&bar), 0));   // peer -> leader's barrier
      }
    }
    if (rank == 0 && threadIdx.x == 0) mbar_wait(smem_u32(&bar), it & 1);
    cluster_sync_all();      // peer may now read its smem: leader observed both halves
    if (threadIdx.x < 32) {
      const uint32_t v = reinterpret_cast<uint32_t*>(smem)[threadIdx.x * 64];
      const uint32_t want = src[it * 4096 + (rank ? 2048 : 0) + threadIdx.x * 64];
      if (v != want) atomicAdd(ok, 1);
    }
    cluster_sync_all();
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = (clock64() - t0) / 64;
}
int main() {
  uint32_t* src; cudaMalloc(&src, 2 << 20);
  uint32_t* h = (uint32_t*)malloc(2 << 20);
  for (int i = 0; i < (2 << 20) / 4; ++i) h[i] = i * 2654435761u;
  cudaMemcpy(src, h, 2 << 20, cudaMemcpyHostToDevice);
  int* ok; cudaMallocManaged(&ok, 4); *ok = 0;
  long long* cyc; cudaMallocManaged(&cyc, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  k<<<148, 128, 64 * 1024>>>(src, ok, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s, mismatches %d, cycles per round %lld\n", cudaGetErrorString(e), *ok, cyc[0]);
}
