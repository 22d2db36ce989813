// Microbenchmark: cycles per tcgen05.mma (M=128, K=8, tf32) for different smem operand layouts.
// Timing only - operand contents are garbage.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../world-model-rl_b200/csrc
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace wmrl::ptx;

__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = make_smem_desc(addr, lbo, sbo);
  d |= (uint64_t)layout << 61;
  return d;
}

// mode 0: no-swizzle [k8][rows][8] (LBO = rows*16, SBO = 128); mode 1: SWIZZLE_128B (rows of 128 B, SBO 1024);
// mode 2: SWIZZLE_64B (rows of 64 B, SBO 512); mode 3: SWIZZLE_32B (rows of 32 B, SBO 256)
__global__ void __launch_bounds__(128, 1) mma_rate(int mode, int N, int iters, int nbuf, int nacc, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const uint32_t base = smem_u32(smem);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tslot), 512); tmem_relinquish(); }
  for (int i = threadIdx.x; i < (200 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  tc_fence_before(); __syncthreads(); tc_fence_after();
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t tm = tslot;
    const uint32_t idesc = ((1u<<4)|(2u<<7)|(2u<<10)|((uint32_t)(N>>3)<<17)|((128u>>4)<<24));
    // operand buffers: A tile region 16 KB each, B tile region 32 KB each, nbuf of them
    // lean issue loop: descriptors built once, only the low word (address field) advances
    const uint64_t ad0 = desc(base, 2080, 128, 0), bd0 = desc(base + 65536, N * 16 + 32, 128, 0);
    const uint32_t ahi = (uint32_t)(ad0 >> 32), bhi = (uint32_t)(bd0 >> 32);
    uint32_t alo = (uint32_t)ad0, blo = (uint32_t)bd0;
    const uint32_t astep = 4160 >> 4, bstep = (N * 32 + 64) >> 4;
    long long t0 = clock64();
    uint32_t col = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t acc = (it > 0 || (int)(col / N) + 0 < 0) ? 1u : (uint32_t)(it * 4 + k >= nacc);
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
            "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
            "setp.ne.b32 p, %6, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}" ::"r"(tm + col),
            "r"(alo + k * astep), "r"(ahi), "r"(blo + k * bstep), "r"(bhi), "r"(idesc), "r"(acc)
            : "memory");
        col += N;
        if (col >= (uint32_t)(nacc * N)) col = 0;
      }
    }
    tc_commit(smem_u32(&bar));
    long long t1 = clock64();
    mbar_wait(smem_u32(&bar), 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tslot, 512);
}

int main() {
  long long* out; cudaMallocManaged(&out, 16);
  cudaFuncSetAttribute(mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const char* names[] = {"no-swizzle", "swizzle128", "swizzle64", "swizzle32"};
  for (int grid : {148})
    for (int N : {256, 128, 64, 32, 16})
      for (int mode = 0; mode < 1; ++mode)
        for (int nacc : {1, 2, 3, 4, 8}) {
          if (nacc * N > 512) continue;
          const int iters = 256, nbuf = 4;
          mma_rate<<<grid, 128, 200 * 1024>>>(mode, N, iters, nbuf, nacc, out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          printf("N %3d %-10s independent accumulators %d: issue %7.1f cyc/MMA, complete %7.1f cyc/MMA (ideal %d)\n", N,
                 names[mode], nacc, (double)out[0] / (iters * 4), (double)out[1] / (iters * 4), N / 2);
        }
  return 0;
}
