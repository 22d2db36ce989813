// Strided GEMM descriptor shared by the SIMT kernels (f32_path.cu) and the tcgen05 3xTF32 kernel
// (tf32_gemm.cu):  C[M,N] (+)= act([A1 | A2] * B + bias), every operand addressed by element strides.
#pragma once
#include <cuda_runtime.h>

namespace wmrl {

struct GemmArgs {
  const float* A1; long long a1_si, a1_sk; int K1;   // A(i,k) = A1[i*a1_si + k*a1_sk], k <  K1
  const float* A2; long long a2_si, a2_sk; int K2;   //          A2[i*a2_si + (k-K1)*a2_sk], k >= K1
  const float* Bm; long long b_sk, b_sj;             // B(k,j) = Bm[k*b_sk + j*b_sj]
  float* C; long long ldc;
  const float* bias;  // per output column, added by z-slice 0; may be null
  int M, N;
  int act;         // 0 none, 1 ELU (only with a single K slice)
  int accumulate;  // C += result
  int kper;        // K elements per z-slice (multiple of 16)
};

// tcgen05 path: fp32 operands split into TF32 hi + lo parts, three MMAs per product (fp32-level
// accuracy), fp32 accumulation in TMEM.  splitk > 1 adds the partial tiles into C with atomics.
int tf32x3_gemm_launch(cudaStream_t st, const GemmArgs& g, int splitk);

}  // namespace wmrl
