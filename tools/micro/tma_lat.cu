// Microbenchmark: latency / throughput of cp.async.bulk global->shared (L2-resident source shared by all CTAs).
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace wmrl::ptx;

__global__ void __launch_bounds__(128, 1) tma_lat(const uint8_t* src, size_t span, int bytes, int depth, int iters,
                                                  int nissuers, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[32];
  const uint32_t base = smem_u32(smem);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 32; ++i) mbar_init(smem_u32(&bars[i]), 1);
    fence_barrier_init();
  }
  __syncthreads();
  const int issuer = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && issuer < nissuers) {
    // warm L2
    size_t off = (size_t)issuer * 65536;
    const int slot = 98304 / depth / 1024 * 1024;
    long long t0 = clock64();
    // keep `depth` copies in flight
    int issued = 0, done = 0;
    uint32_t phases = 0;
    while (done < iters) {
      while (issued < iters && issued - done < depth) {
        const int s = issued % depth;
        mbar_arrive_expect_tx(smem_u32(&bars[issuer * 8 + s]), bytes);
        bulk_g2s(base + issuer * 98304 + s * slot, src + off, bytes, smem_u32(&bars[issuer * 8 + s]));
        off += bytes; if (off + bytes > span) off = 0;
        ++issued;
      }
      const int s = done % depth;
      mbar_wait(smem_u32(&bars[issuer * 8 + s]), (phases >> s) & 1u);
      phases ^= 1u << s;
      ++done;
    }
    long long t1 = clock64();
    if (blockIdx.x == 0 && issuer == 0) out[0] = t1 - t0;
  }
}

int main() {
  uint8_t* src; const size_t span = 2 << 20;
  cudaMalloc(&src, span); cudaMemset(src, 1, span);
  long long* out; cudaMallocManaged(&out, 16);
  cudaFuncSetAttribute(tma_lat, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int grid : {148})
    for (int nissuers : {1, 2})
      for (int bytes : {8192, 16384, 32768, 49152})
        for (int depth : {1, 2, 3}) {
          if (bytes * depth > 98304) continue;
          const int iters = 512;
          tma_lat<<<grid, 128, 200 * 1024>>>(src, span, bytes, depth, iters, nissuers, out);
          tma_lat<<<grid, 128, 200 * 1024>>>(src, span, bytes, depth, iters, nissuers, out);
          if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
          printf("issuers %d bytes %5d depth %d: %7.0f cycles per copy per issuer, %6.1f B/clk/SM total\n", nissuers, bytes,
                 depth, (double)out[0] / iters, (double)bytes * iters * nissuers / out[0]);
        }
  return 0;
}
