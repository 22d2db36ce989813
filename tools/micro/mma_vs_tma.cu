// Does a concurrent bulk-copy stream into shared memory slow tcgen05.mma down?  One lane issues back-to-back MMAs
// (M=128, N=256, K=16, bf16, no-swizzle panels: 12 KB of operand reads per 128-cycle MMA = 96 B/clk); another warp's lane
// keeps `depth` bulk copies of `bytes` in flight into a separate region of the same CTA's shared memory.
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace wmrl::ptx;

__global__ void __launch_bounds__(128, 1) k(const uint8_t* src, int copy_bytes, int depth, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, cbars[8];
  __shared__ uint32_t tslot;
  __shared__ volatile long long n_shared, t_mma_end;
    const uint32_t base = smem_u32(smem);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&cbars[i]), 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tslot), 512); tmem_relinquish(); }
  for (int i = threadIdx.x; i < (96 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  tc_fence_before(); __syncthreads(); tc_fence_after();
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t tm = tslot;
    const uint32_t idesc = make_idesc_bf16(128, 256);
    const uint64_t ad0 = make_smem_desc(base, 2048, 128), bd0 = make_smem_desc(base + 32768, 4096, 128);
    const uint32_t ahi = (uint32_t)(ad0 >> 32), bhi = (uint32_t)(bd0 >> 32);
    const uint32_t alo = (uint32_t)ad0, blo = (uint32_t)bd0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        mma_bf16_lohi(tm, alo + kk * (4096 >> 4), ahi, blo + kk * (8192 >> 4), bhi, idesc, 1u);
    }
    tc_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    t_mma_end = t1;
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[3] = n_shared; out[4] = t0; }
  } else if (threadIdx.x == 32 && depth > 0) {
    // copy stream into smem [96 KB, 96 KB + depth * copy_bytes): a fixed number of rounds that outlasts the MMA loop
    const int rounds = (iters * 4 * 200) / (400 * 1) + 8;
    uint32_t ph = 0;
    long long n = 0, t_first = clock64();
    for (int s = 0; s < depth; ++s) {
      mbar_arrive_expect_tx(smem_u32(&cbars[s]), copy_bytes);
      bulk_g2s(base + 96 * 1024 + s * copy_bytes, src + (size_t)s * copy_bytes, copy_bytes, smem_u32(&cbars[s]));
    }
    for (int r = 0; r < rounds; ++r) {
      for (int s = 0; s < depth; ++s) {
        mbar_wait(smem_u32(&cbars[s]), ph);
        ++n;
        n_shared = n;
        if (r + 1 < rounds) {
          mbar_arrive_expect_tx(smem_u32(&cbars[s]), copy_bytes);
          bulk_g2s(base + 96 * 1024 + s * copy_bytes, src + ((size_t)(n * 7 + s) * copy_bytes) % (1 << 21), copy_bytes,
                   smem_u32(&cbars[s]));
        }
      }
      ph ^= 1;
    }
    if (blockIdx.x == 0) { out[1] = n; out[2] = clock64() - t_first; }
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tslot, 512);
}

int main() {
  uint8_t* src; cudaMalloc(&src, 4 << 20); cudaMemset(src, 1, 4 << 20);
  long long* out; cudaMallocManaged(&out, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
  const int iters = 512;
  for (int bytes : {0, 16384, 32768})
    for (int depth : {1, 2, 3}) {
      if (bytes == 0 && depth > 1) continue;
      out[0] = out[1] = out[2] = out[3] = out[4] = 0;
      for (int rep = 0; rep < 2; ++rep) k<<<148, 128, 224 * 1024>>>(src, bytes, bytes ? depth : 0, iters, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      const double cyc = (double)out[0];
      printf("copy stream %5d B x depth %d: %6.1f cycles per MMA (ideal 128); copies: %5.1f B/clk over their whole run, %5.1f B/clk while the MMAs ran\n",
             bytes, bytes ? depth : 0, cyc / (iters * 4), out[2] ? (double)out[1] * bytes / (double)out[2] : 0.0,
             (double)out[3] * bytes / cyc);
    }
  return 0;
}
