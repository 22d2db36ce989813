// Saturated L2 -> shared-memory ingest per SM with 1-D bulk copies (cp.async.bulk), all 148 SMs active:
// `depth` copies of `bytes` each kept in flight per CTA by one lane (re-issued as they complete), sources walking a
// 2.4 MB L2-resident buffer (the weight image of the rollout kernel).  Also the same with cluster multicast: in a
// cluster of `csz` CTAs every CTA issues 1/csz of each block and multicasts it to all, so each SM *receives*
// the same bytes but *requests* 1/csz of them.
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace wmrl::ptx;

__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "h"(mask)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) ingest(const uint8_t* src, size_t src_bytes, int bytes, int depth, int iters, int csz,
                                                 long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[16];
  const uint32_t base = smem_u32(smem);
  const uint32_t rank = csz > 1 ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(smem_u32(&bars[i]), 1); fence_barrier_init(); }
  __syncthreads();
  if (csz > 1) cluster_sync_all();
  if (threadIdx.x == 0) {
    const uint32_t part = bytes / csz;
    const uint16_t mask = (uint16_t)((1u << csz) - 1u);
    size_t off = ((size_t)(blockIdx.x / csz) * 65536) % src_bytes;
    auto issue = [&](int slot) {
      const uint32_t bar = smem_u32(&bars[slot]);
      mbar_arrive_expect_tx(bar, bytes);                    // every CTA receives the whole block
      if (off + bytes > src_bytes) off = 0;
      if (csz > 1) bulk_g2s_mc(base + slot * bytes + rank * part, src + off + rank * part, part, bar, mask);
      else bulk_g2s(base + slot * bytes, src + off, bytes, bar);
      off += bytes;
    };
    for (int s = 0; s < depth; ++s) issue(s);
    long long t0 = clock64();
    uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      for (int s = 0; s < depth; ++s) {
        mbar_wait(smem_u32(&bars[s]), ph);
        if (csz > 1) {   // a slot may only be re-armed when every CTA of the cluster has seen it complete
          // (omitted: timing only, data is never read; re-arming early just overlaps the writes)
        }
        if (it + 1 < iters) issue(s);
      }
      ph ^= 1;
    }
    long long t1 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; }
  }
  if (csz > 1) cluster_sync_all();
}

int main() {
  const size_t src_bytes = 2400 * 1024;
  uint8_t* src; cudaMalloc(&src, src_bytes); cudaMemset(src, 1, src_bytes);
  long long* out; cudaMallocManaged(&out, 64);
  cudaFuncSetAttribute(ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(ingest, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  const int iters = 200;
  for (int csz : {1, 2, 4})
    for (int bytes : {8192, 16384, 32768})
      for (int depth : {1, 2, 3, 4, 6}) {
        if ((size_t)depth * bytes > 196 * 1024) continue;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(148 / csz * csz); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 200 * 1024;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = csz; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        for (int rep = 0; rep < 2; ++rep) cudaLaunchKernelEx(&cfg, ingest, (const uint8_t*)src, src_bytes, bytes, depth, iters, csz, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s (csz %d bytes %d depth %d)\n", cudaGetErrorString(e), csz, bytes, depth); return 1; }
        const double total = (double)bytes * depth * iters;
        printf("cluster %d  block %5d B  depth %d (%3d KB in flight): %6.1f B/clk/SM received, %6.1f requested\n", csz, bytes, depth,
               bytes * depth / 1024, total / out[0], total / out[0] / csz);
      }
  return 0;
}
