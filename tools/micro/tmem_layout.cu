// Where do the rows of D land in TMEM for tcgen05.mma cta_group::2 with M=128 (64 rows per CTA)?
// A[r][0] = 1 + r (global row id), other A = 0; B[n][0] = 1  ->  D[r][n] = 1 + r.
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace wmrl::ptx;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k(int M, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const uint32_t rank = cluster_ctarank();
  const int rows = M / 2;                 // rows of A held by this CTA
  __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(smem);            // [2 k8][rows][8]
  __nv_bfloat16* B = reinterpret_cast<__nv_bfloat16*>(smem + 8192);     // [2 k8][8 rows (N/2)][8]
  for (int i = threadIdx.x; i < 4096; i += 128) { A[i] = __float2bfloat16(0.f); B[i] = __float2bfloat16(0.f); }
  __syncthreads();
  if (threadIdx.x < rows) A[threadIdx.x * 8] = __float2bfloat16((float)(1 + threadIdx.x + rank * rows));
  if (threadIdx.x < 8) B[threadIdx.x * 8] = __float2bfloat16(1.f);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc2(smem_u32(&tslot), 32); tmem_relinquish2(); }
  fence_proxy_async_smem();
  tc_fence_before(); cluster_sync_all(); tc_fence_after();
  const uint32_t tm = tslot;
  // sentinel: fill this warp's lanes with -7 first (tcgen05.st 32x32b.x16)
  {
    const uint32_t t = tm + ((uint32_t)((threadIdx.x >> 5) * 32) << 16);
    uint32_t v = __float_as_uint(-7.f);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(t), "r"(v) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before(); cluster_sync_all(); tc_fence_after();
  if (rank == 0 && threadIdx.x == 0) {
    const uint64_t ad = make_smem_desc(smem_u32(A), rows * 16, 128);
    const uint64_t bd = make_smem_desc(smem_u32(B), 8 * 16, 128);
    mma_bf16_ss2(tm, ad, bd, make_idesc_bf16(M, 16), 0);
    tc_commit2(smem_u32(&bar), 3);
  }
  if (threadIdx.x == 0) mbar_wait(smem_u32(&bar), 0);
  __syncthreads();
  tc_fence_after();
  float v[16];
  tmem_ld16(tm + ((uint32_t)((threadIdx.x >> 5) * 32) << 16), v);
  tmem_ld_wait();
  out[(rank * 128 + threadIdx.x) * 2 + 0] = v[0];
  out[(rank * 128 + threadIdx.x) * 2 + 1] = v[15];
  tc_fence_before(); cluster_sync_all();
  if (threadIdx.x < 32) tmem_dealloc2(tm, 32);
}
int main() {
  float* out; cudaMallocManaged(&out, 2 * 128 * 2 * sizeof(float));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  for (int M : {256, 128}) {
    k<<<2, 128, 32768>>>(M, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("M=%d status %s\n", M, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    for (int r = 0; r < 2; ++r) {
      printf(" CTA %d: lane -> D[.,0] (col15 in brackets if different)\n  ", r);
      for (int l = 0; l < 128; ++l) {
        float a = out[(r * 128 + l) * 2], b = out[(r * 128 + l) * 2 + 1];
        if (a == b) printf("%g ", a); else printf("%g[%g] ", a, b);
        if (l % 32 == 31) printf("\n  ");
      }
      printf("\n");
    }
  }
}
