// What does the per-stage protocol of the folded-pair kernel cost the issuing lane?  cta_group::2, M=128, N=256, 8 MMAs
// per "stage"; variants: (0) MMAs only, (1) + tcgen05.commit (multicast to both CTAs) per stage, (2) + a try_wait on an
// already-complete mbarrier per stage, (3) both, (4) both + the wait placed in the middle of the stage's MMAs.
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace wmrl::ptx;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k(int variant, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, done_bar, ring_bar[4];
  __shared__ uint32_t tslot;
  const uint32_t rank = cluster_ctarank();
  const uint32_t base = smem_u32(smem);
  for (int i = threadIdx.x; i < (160 * 1024) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&done_bar), 1);
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&ring_bar[i]), 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) { tmem_alloc2(smem_u32(&tslot), 512); tmem_relinquish2(); }
  fence_proxy_async_smem();
  tc_fence_before(); cluster_sync_all(); tc_fence_after();
  const uint32_t tm = tslot;
  if (rank == 0 && threadIdx.x == 0) {
    mbar_arrive(smem_u32(&bar));                  // phase 0 of `bar` is complete: waits on parity 0 return at once
    const int rows = 64, nh = 128;
    const uint64_t ad0 = make_smem_desc(base, rows * 16, 128), bd0 = make_smem_desc(base + 65536, nh * 16, 128);
    const uint32_t ahi = (uint32_t)(ad0 >> 32), bhi = (uint32_t)(bd0 >> 32);
    const uint32_t alo = (uint32_t)ad0, blo = (uint32_t)bd0;
    const uint32_t astep = (rows * 32) >> 4, bstep = (nh * 32) >> 4;
    const uint32_t idesc = make_idesc_bf16(128, 256);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if ((variant & 2) && !(variant & 4)) { mbar_wait(smem_u32(&bar), 0); tc_fence_after(); }
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        mma_bf16_lohi2(tm, alo + (kk & 3) * astep, ahi, blo + (kk & 3) * bstep, bhi, idesc, (it | kk) ? 1u : 0u);
        if (kk == 3 && (variant & 4)) { mbar_wait(smem_u32(&bar), 0); tc_fence_after(); }
      }
      if (variant & 1) tc_commit2(smem_u32(&ring_bar[it & 3]), 3);
    }
    tc_commit2(smem_u32(&done_bar), 3);
    mbar_wait(smem_u32(&done_bar), 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) out[0] = t2 - t0;
  }
  if (rank == 1 && threadIdx.x == 0) mbar_wait(smem_u32(&done_bar), 0);
  tc_fence_before(); cluster_sync_all();
  if (threadIdx.x < 32) tmem_dealloc2(tm, 512);
}
int main() {
  long long* out; cudaMallocManaged(&out, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  const char* names[] = {"MMAs only", "+ commit per stage", "+ wait(complete barrier) per stage", "+ commit + wait", "", "", "", "+ commit + wait in mid-stage"};
  for (int variant : {0, 1, 2, 3, 7}) {
    const int iters = 256;
    for (int rep = 0; rep < 2; ++rep) k<<<148, 128, 160 * 1024>>>(variant, iters, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    printf("%-38s %7.1f cycles per 8-MMA stage (ideal 8 x 64.4 = 515)\n", names[variant], (double)out[0] / iters);
  }
}
