// fp32 SIMT path of libwmrl: the validation-mode forward (rtol 1e-5 against the
// reference) and the hand-written BPTT / full backward for both RSSM variants.
//
// Every step of the reference loops (rssm.py:123-140 imagine, :93-121 observe)
// is issued as a short chain of kernels on the caller's stream: one generic
// tiled SGEMM (2-segment A operand = torch.cat-free Linear, arbitrary strides so
// it reads/writes the [B,T,.] sequence tensors in place, split-K atomics for
// the weight gradients) plus fused element-wise kernels (LayerNorm+ELU, GRU
// gates, sampling).  No cuBLAS, no host synchronisation.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "gemm.cuh"

namespace wmrl {

// ---------------------------------------------------------------------------
// Generic SGEMM:  C[M,N] (+)= A[M,K] * B[K,N]  with  A(i,k) = A[i*si + k*sk]
// (two K-segments), B(k,j) = B[k*sk + j*sj].
// ---------------------------------------------------------------------------
constexpr int BM = 64, BN = 64, BK = 16;

__device__ __forceinline__ float elu_f(float x) { return x > 0.f ? x : expm1f(x); }

__global__ void __launch_bounds__(256) sgemm_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int K = g.K1 + g.K2;
  const int kbeg = blockIdx.z * g.kper;
  const int kend = min(K, kbeg + g.kper);
  const bool a_fastk = (g.a1_sk == 1);
  const bool b_fastj = (g.b_sj == 1);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
    for (int e = tid; e < BM * BK; e += 256) {
      int i, k;
      if (a_fastk) { k = e % BK; i = e / BK; } else { i = e % BM; k = e / BM; }
      const int gm = m0 + i, gk = k0 + k;
      float v = 0.f;
      if (gm < g.M && gk < kend) {
        v = (gk < g.K1) ? g.A1[(long long)gm * g.a1_si + (long long)gk * g.a1_sk]
                        : g.A2[(long long)gm * g.a2_si + (long long)(gk - g.K1) * g.a2_sk];
      }
      As[k][i] = v;
    }
#pragma unroll
    for (int e = tid; e < BN * BK; e += 256) {
      int j, k;
      if (b_fastj) { j = e % BN; k = e / BN; } else { k = e % BK; j = e / BK; }
      const int gn = n0 + j, gk = k0 + k;
      float v = 0.f;
      if (gn < g.N && gk < kend) v = g.Bm[(long long)gk * g.b_sk + (long long)gn * g.b_sj];
      Bs[k][j] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const bool split = gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= g.N) continue;
      float v = acc[i][j];
      if (g.bias && blockIdx.z == 0) v += g.bias[gn];
      float* c = g.C + (long long)gm * g.ldc + gn;
      if (split) {
        atomicAdd(c, v);
      } else {
        if (g.act == 1) v = elu_f(v);
        *c = g.accumulate ? (*c + v) : v;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Skinny GEMM for M <= 64 rows (posterior filtering at B ~ 50).  The 64x64-tile kernel above would
// launch N/64 blocks that each walk all of K serially; at these sizes the weight matrix (a few MB,
// L2-resident) is the traffic and latency is the cost, so the work is spread over every SM instead:
//   grid = (N/32 column strips, split), cluster = (1, split, 1): the CTAs of a cluster share a strip
//   and each takes 1/split of K.  A CTA stages its [64 x kt] A tile and [32 x kt] B tile with 4-byte
//   cp.async (any strides / alignment), two tiles in flight, and computes 4 rows x 2 columns per
//   thread with 128-bit shared loads along k.  Partial tiles are reduced through distributed shared
//   memory: rank r sums rows [r*64/split, ...) over all ranks, adds bias / activation and stores.
// No atomics, so the result is deterministic and C needs no zero fill.
// ---------------------------------------------------------------------------
constexpr int SK_M = 64, SK_N = 32, SK_KT = 128, SK_KS = SK_KT + 4;
constexpr int SK_STAGE_FLOATS = (SK_M + SK_N) * SK_KS;
constexpr int SK_SMEM_BYTES = (2 * SK_STAGE_FLOATS + SK_M * SK_N) * (int)sizeof(float);

__device__ __forceinline__ void cp_async4(float* dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(ptx::smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory");
}
__device__ __forceinline__ float ld_dsmem(uint32_t addr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

__device__ __forceinline__ void sk_stage_tile(const GemmArgs& g, float* As, float* Bs, int k0,
                                              int kend, int n0, int tid) {
  const bool a_fastk = (g.a1_sk == 1);
  const bool b_fastk = (g.b_sk == 1);
#pragma unroll 4
  for (int e = tid; e < SK_M * SK_KT; e += 256) {
    int m, k;
    if (a_fastk) { k = e % SK_KT; m = e / SK_KT; } else { m = e % SK_M; k = e / SK_M; }
    const int gk = k0 + k;
    float* dst = As + m * SK_KS + k;
    if (m < g.M && gk < kend) {
      const float* src = (gk < g.K1) ? g.A1 + (long long)m * g.a1_si + (long long)gk * g.a1_sk
                                     : g.A2 + (long long)m * g.a2_si + (long long)(gk - g.K1) * g.a2_sk;
      cp_async4(dst, src);
    } else {
      *dst = 0.f;
    }
  }
#pragma unroll 4
  for (int e = tid; e < SK_N * SK_KT; e += 256) {
    int j, k;
    if (b_fastk) { k = e % SK_KT; j = e / SK_KT; } else { j = e % SK_N; k = e / SK_N; }
    const int gk = k0 + k, gn = n0 + j;
    float* dst = Bs + j * SK_KS + k;
    if (gn < g.N && gk < kend) cp_async4(dst, g.Bm + (long long)gk * g.b_sk + (long long)gn * g.b_sj);
    else *dst = 0.f;
  }
}

__global__ void __launch_bounds__(256) sgemm_skinny_kernel(GemmArgs g) {
  extern __shared__ __align__(16) float sk_smem[];
  float* red = sk_smem + 2 * SK_STAGE_FLOATS;
  const int tid = threadIdx.x;
  const int tx = tid & 15;   // columns tx, tx + 16
  const int ty = tid >> 4;   // rows ty*4 .. ty*4+3
  const int n0 = blockIdx.x * SK_N;
  const int split = gridDim.y;
  const int rank = blockIdx.y;   // == %cluster_ctarank for cluster dims (1, split, 1)
  const int K = g.K1 + g.K2;
  const int kbeg = min(K, rank * g.kper), kend = min(K, kbeg + g.kper);
  const int ntiles = (kend - kbeg + SK_KT - 1) / SK_KT;

  float acc[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = 0.f;

  if (ntiles > 0) sk_stage_tile(g, sk_smem, sk_smem + SK_M * SK_KS, kbeg, kend, n0, tid);
  cp_async_commit();
  for (int t = 0; t < ntiles; ++t) {
    float* As = sk_smem + (t & 1) * SK_STAGE_FLOATS;
    float* Bs = As + SK_M * SK_KS;
    if (t + 1 < ntiles) {
      float* An = sk_smem + ((t + 1) & 1) * SK_STAGE_FLOATS;
      sk_stage_tile(g, An, An + SK_M * SK_KS, kbeg + (t + 1) * SK_KT, kend, n0, tid);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const int klen = min(SK_KT, kend - (kbeg + t * SK_KT));
    const int k4n = (klen + 3) >> 2;   // the tail of the last float4 is zero-filled
    const float* ap = As + (ty * 4) * SK_KS;
    const float* bp = Bs + tx * SK_KS;
#pragma unroll 2
    for (int k4 = 0; k4 < k4n; ++k4) {
      const float4 b0 = *reinterpret_cast<const float4*>(bp + k4 * 4);
      const float4 b1 = *reinterpret_cast<const float4*>(bp + 16 * SK_KS + k4 * 4);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 a = *reinterpret_cast<const float4*>(ap + i * SK_KS + k4 * 4);
        acc[i][0] = fmaf(a.x, b0.x, fmaf(a.y, b0.y, fmaf(a.z, b0.z, fmaf(a.w, b0.w, acc[i][0]))));
        acc[i][1] = fmaf(a.x, b1.x, fmaf(a.y, b1.y, fmaf(a.z, b1.z, fmaf(a.w, b1.w, acc[i][1]))));
      }
    }
    __syncthreads();   // the buffer is refilled by the staging of tile t + 2
  }

  if (split == 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = ty * 4 + i;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int n = n0 + tx + 16 * j;
        if (m < g.M && n < g.N) {
          float v = acc[i][j] + (g.bias ? g.bias[n] : 0.f);
          if (g.act == 1) v = elu_f(v);
          float* c = g.C + (long long)m * g.ldc + n;
          *c = g.accumulate ? (*c + v) : v;
        }
      }
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    red[(ty * 4 + i) * SK_N + tx] = acc[i][0];
    red[(ty * 4 + i) * SK_N + tx + 16] = acc[i][1];
  }
  ptx::cluster_sync_all();
  const int rows = SK_M / split;   // split is 2, 4 or 8
  for (int o = tid; o < rows * SK_N; o += 256) {
    const int m = rank * rows + o / SK_N, j = o % SK_N, n = n0 + j;
    const uint32_t local = ptx::smem_u32(red + m * SK_N + j);
    float v = 0.f;
    for (int r = 0; r < split; ++r) v += ld_dsmem(ptx::mapa(local, (uint32_t)r));
    if (m < g.M && n < g.N) {
      if (g.bias) v += g.bias[n];
      if (g.act == 1) v = elu_f(v);
      float* c = g.C + (long long)m * g.ldc + n;
      *c = g.accumulate ? (*c + v) : v;
    }
  }
  ptx::cluster_sync_all();   // peers may still be reading this CTA's partial tile
}

static int launch_skinny(cudaStream_t st, GemmArgs g) {
  static bool attr_set = false;
  if (!attr_set) {
    WMRL_CUDA_OK(cudaFuncSetAttribute(sgemm_skinny_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      SK_SMEM_BYTES));
    attr_set = true;
  }
  const int K = g.K1 + g.K2;
  const int strips = (g.N + SK_N - 1) / SK_N;
  int split = 1;
  while (split < 8 && strips * split < 148 && K / (split * 2) >= 32) split *= 2;
  g.kper = ((K + split - 1) / split + 3) / 4 * 4;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(strips, split, 1);
  cfg.blockDim = dim3(256, 1, 1);
  cfg.dynamicSmemBytes = SK_SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = split; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  WMRL_CUDA_OK(cudaLaunchKernelEx(&cfg, sgemm_skinny_kernel, g));
  return 0;
}

static int launch_gemm(cudaStream_t st, GemmArgs g, int splitk) {
  if (g.M <= 0 || g.N <= 0) return 0;
  if (splitk <= 1 && g.M <= SK_M && (g.K1 + g.K2) >= 128) return launch_skinny(st, g);
  // More than 64 rows: tensor cores (3xTF32, fp32-accurate).  WMRL_F32_GEMM=simt keeps the SIMT tile
  // kernel below for A/B comparisons.
  static const bool use_simt = [] {
    const char* e = getenv("WMRL_F32_GEMM");
    return e && e[0] == 's';
  }();
  if (!use_simt && g.M > SK_M) {
    if (splitk > 1) {
      // split-K (weight gradients: K = batch rows): aim at two 128-row tiles per SM, >= 256 k per slice
      const int K = g.K1 + g.K2;
      const int tiles = ((g.M + 127) / 128) * ((g.N + 255) / 256);
      splitk = (296 + tiles - 1) / tiles;
      const int maxsplit = (K + 255) / 256;
      if (splitk > maxsplit) splitk = maxsplit;
      if (splitk < 2) splitk = 2;   // callers rely on += semantics through the atomic epilogue
    }
    return tf32x3_gemm_launch(st, g, splitk);
  }
  const int K = g.K1 + g.K2;
  if (splitk < 1) splitk = 1;
  int kper = ((K + splitk - 1) / splitk + BK - 1) / BK * BK;
  if (kper < BK) kper = BK;
  splitk = (K + kper - 1) / kper;
  if (splitk < 1) splitk = 1;
  g.kper = kper;
  dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, splitk);
  sgemm_kernel<<<grid, 256, 0, st>>>(g);
  WMRL_CUDA_OK(cudaGetLastError());
  return 0;
}

// Y[M,N] = act([X1 | X2] W^T + b), W[N, K1+K2] row-major (torch Linear).
static int linear(cudaStream_t st, const float* X1, long long ld1, int K1, const float* X2,
                  long long ld2, int K2, const float* W, const float* bias, float* Y,
                  long long ldy, int M, int N, int act) {
  GemmArgs g{};
  g.A1 = X1; g.a1_si = ld1; g.a1_sk = 1; g.K1 = K1;
  g.A2 = X2; g.a2_si = ld2; g.a2_sk = 1; g.K2 = K2;
  g.Bm = W; g.b_sk = 1; g.b_sj = K1 + K2;
  g.C = Y; g.ldc = ldy; g.bias = bias; g.M = M; g.N = N; g.act = act; g.accumulate = 0;
  return launch_gemm(st, g, 1);
}

// dX[M,Kout] (+)= dY[M,N] * W[:, koff:koff+Kout], W[N,Ktot] row-major.
static int dgrad(cudaStream_t st, const float* dY, long long lddy, int N, const float* W, int Ktot,
                 int koff, int Kout, float* dX, long long lddx, int M, int accumulate) {
  GemmArgs g{};
  g.A1 = dY; g.a1_si = lddy; g.a1_sk = 1; g.K1 = N;
  g.A2 = nullptr; g.K2 = 0; g.a2_si = 0; g.a2_sk = 1;
  g.Bm = W + koff; g.b_sk = Ktot; g.b_sj = 1;
  g.C = dX; g.ldc = lddx; g.bias = nullptr; g.M = M; g.N = Kout; g.act = 0;
  g.accumulate = accumulate;
  return launch_gemm(st, g, 1);
}

// dW[:, koff:koff+Kx] += dY[M,N]^T * X[M,Kx]   (dW is [N,Ktot] row-major), split over M.
static int wgrad(cudaStream_t st, const float* dY, long long lddy, int N, const float* X,
                 long long ldx, int Kx, float* dW, int Ktot, int koff, int M) {
  if (!dW) return 0;
  GemmArgs g{};
  g.A1 = dY; g.a1_si = 1; g.a1_sk = lddy; g.K1 = M;   // A(i=n, k=m) = dY[m][n]
  g.A2 = nullptr; g.K2 = 0; g.a2_si = 0; g.a2_sk = 1;
  g.Bm = X; g.b_sk = ldx; g.b_sj = 1;                  // B(k=m, j) = X[m][j]
  g.C = dW + koff; g.ldc = Ktot; g.bias = nullptr; g.M = N; g.N = Kx; g.act = 0;
  g.accumulate = 1;
  const int tiles = ((N + BM - 1) / BM) * ((Kx + BN - 1) / BN);
  int splitk = (296 + tiles - 1) / tiles;
  const int maxsplit = (M + 127) / 128;
  if (splitk > maxsplit) splitk = maxsplit;
  if (splitk < 2) splitk = 2;   // always the atomic path: += semantics
  if (M <= BK) splitk = 1;
  if (splitk == 1) return launch_gemm(st, g, 1);
  return launch_gemm(st, g, splitk);
}

// ---------------------------------------------------------------------------
// Element-wise kernels
// ---------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

// out[b,j] = (a? a[b*lda+j] : 0) + (c? c[b*ldc+j] : 0)
__global__ void combine_kernel(const float* a, long long lda, const float* c, long long ldc,
                               float* out, long long ldo, int Bn, int W) {
  const long long n = (long long)Bn * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / W), j = (int)(i % W);
    float v = 0.f;
    if (a) v += a[b * lda + j];
    if (c) v += c[b * ldc + j];
    out[b * ldo + j] = v;
  }
}

// LayerNorm(eps 1e-5, affine) + activation, one warp per row (actor.py:27-36 with ELU; the heads next to
// the path use SiLU - models.py:3-37 - and ReLU - trainers/value/critic.py:4-29).
// kAct: 0 ELU, 1 SiLU, 2 ReLU.
template <int kAct>
__device__ __forceinline__ float act_f(float z) {
  if (kAct == 0) return z > 0.f ? z : expm1f(z);
  if (kAct == 1) return z / (1.f + expf(-z));
  return fmaxf(z, 0.f);
}
template <int kAct>
__device__ __forceinline__ float dact_f(float z) {
  if (kAct == 0) return z > 0.f ? 1.f : expf(z);
  if (kAct == 1) { const float sg = 1.f / (1.f + expf(-z)); return sg * (1.f + z * (1.f - sg)); }
  return z > 0.f ? 1.f : 0.f;
}

template <int kAct>
__global__ void ln_act_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gam,
                                  const float* __restrict__ bet, float* __restrict__ xhat,
                                  float* __restrict__ rstd, float* __restrict__ y, int Bn, int W) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= Bn) return;
  const float* xr = x + (long long)warp * W;
  float s = 0.f;
  for (int j = lane; j < W; j += 32) s += xr[j];
  const float mean = warp_sum(s) / W;
  float v = 0.f;
  for (int j = lane; j < W; j += 32) { const float d = xr[j] - mean; v += d * d; }
  const float rs = rsqrtf(warp_sum(v) / W + 1e-5f);
  if (lane == 0) rstd[warp] = rs;
  for (int j = lane; j < W; j += 32) {
    const float xh = (xr[j] - mean) * rs;
    xhat[(long long)warp * W + j] = xh;
    y[(long long)warp * W + j] = act_f<kAct>(xh * gam[j] + bet[j]);
  }
}

// Backward of the block above: dy -> dx, and per-column dgamma / dbeta.  A lane always handles the same
// columns (lane, lane+32, ...), so for W <= 32*kLnCols it keeps their gamma / beta and the running column
// sums in registers over all the rows its warp walks and adds once at the end (shared, then global).
constexpr int kLnCols = 16;
template <int kAct, bool kRegs>
__global__ void __launch_bounds__(256, 2) ln_act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ xhat,
                                  const float* __restrict__ rstd, const float* __restrict__ gam,
                                  const float* __restrict__ bet, float* __restrict__ dx,
                                  float* dgam, float* dbet, int Bn, int W) {
  extern __shared__ float sm[];  // [2*W] column partial sums
  float* sg = sm;
  float* sb = sm + W;
  for (int j = threadIdx.x; j < 2 * W; j += blockDim.x) sm[j] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float g_r[kLnCols], b_r[kLnCols], ag[kLnCols], ab[kLnCols];
  if (kRegs) {
#pragma unroll
    for (int i = 0; i < kLnCols; ++i) {
      const int j = lane + 32 * i;
      g_r[i] = j < W ? gam[j] : 0.f;
      b_r[i] = j < W ? bet[j] : 0.f;
      ag[i] = 0.f; ab[i] = 0.f;
    }
  }
  for (int row = blockIdx.x * wpb + wib; row < Bn; row += gridDim.x * wpb) {
    const long long o = (long long)row * W;
    const float rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
    if (kRegs) {
#pragma unroll
      for (int i = 0; i < kLnCols; ++i) {
        const int j = lane + 32 * i;
        if (j < W) {
          const float xh = xhat[o + j];
          const float dz = dy[o + j] * dact_f<kAct>(xh * g_r[i] + b_r[i]);
          const float dxh = dz * g_r[i];
          s1 += dxh;
          s2 += dxh * xh;
          ag[i] += dz * xh;
          ab[i] += dz;
        }
      }
      s1 = warp_sum(s1) / W;
      s2 = warp_sum(s2) / W;
#pragma unroll
      for (int i = 0; i < kLnCols; ++i) {
        const int j = lane + 32 * i;
        if (j < W) {
          const float xh = xhat[o + j];
          const float dz = dy[o + j] * dact_f<kAct>(xh * g_r[i] + b_r[i]);
          dx[o + j] = rs * (dz * g_r[i] - s1 - xh * s2);
        }
      }
    } else {
      for (int j = lane; j < W; j += 32) {
        const float xh = xhat[o + j];
        const float dz = dy[o + j] * dact_f<kAct>(xh * gam[j] + bet[j]);
        const float dxh = dz * gam[j];
        s1 += dxh;
        s2 += dxh * xh;
        if (dgam) atomicAdd(&sg[j], dz * xh);
        if (dbet) atomicAdd(&sb[j], dz);
      }
      s1 = warp_sum(s1) / W;
      s2 = warp_sum(s2) / W;
      for (int j = lane; j < W; j += 32) {
        const float xh = xhat[o + j];
        const float dz = dy[o + j] * dact_f<kAct>(xh * gam[j] + bet[j]);
        dx[o + j] = rs * (dz * gam[j] - s1 - xh * s2);
      }
    }
  }
  if (kRegs) {
#pragma unroll
    for (int i = 0; i < kLnCols; ++i) {
      const int j = lane + 32 * i;
      if (j < W) {
        if (dgam) atomicAdd(&sg[j], ag[i]);
        if (dbet) atomicAdd(&sb[j], ab[i]);
      }
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < W; j += blockDim.x) {
    if (dgam) atomicAdd(&dgam[j], sg[j]);
    if (dbet) atomicAdd(&dbet[j], sb[j]);
  }
}

static int ln_act_fwd(cudaStream_t st, int act, const float* x, const float* gam, const float* bet, float* xhat,
                      float* rstd, float* y, int Bn, int W) {
  if (Bn <= 0) return 0;
  const int nb = (int)(((long long)Bn * 32 + 255) / 256);
  if (act == 0) ln_act_fwd_kernel<0><<<nb, 256, 0, st>>>(x, gam, bet, xhat, rstd, y, Bn, W);
  else if (act == 1) ln_act_fwd_kernel<1><<<nb, 256, 0, st>>>(x, gam, bet, xhat, rstd, y, Bn, W);
  else ln_act_fwd_kernel<2><<<nb, 256, 0, st>>>(x, gam, bet, xhat, rstd, y, Bn, W);
  WMRL_CUDA_OK(cudaGetLastError());
  return 0;
}
static int ln_act_bwd(cudaStream_t st, int act, const float* dy, const float* xhat, const float* rstd,
                      const float* gam, const float* bet, float* dx, float* dgam, float* dbet, int Bn, int W) {
  if (Bn <= 0) return 0;
  int nb = (Bn + 7) / 8;              // one row per warp until the GPU is full (4 blocks per SM), then warps loop
  if (nb > 592) nb = 592;             // and the register column sums start to pay off
  if (nb < 1) nb = 1;
  const size_t sh = 2 * (size_t)W * sizeof(float);
  const bool regs = W <= 32 * kLnCols && Bn >= 4 * 592 * 8;   // only when a warp walks >= 4 rows
#define WMRL_LN_BWD(A, R) ln_act_bwd_kernel<A, R><<<nb, 256, sh, st>>>(dy, xhat, rstd, gam, bet, dx, dgam, dbet, Bn, W)
  if (act == 0) { if (regs) WMRL_LN_BWD(0, true); else WMRL_LN_BWD(0, false); }
  else if (act == 1) { if (regs) WMRL_LN_BWD(1, true); else WMRL_LN_BWD(1, false); }
  else { if (regs) WMRL_LN_BWD(2, true); else WMRL_LN_BWD(2, false); }
#undef WMRL_LN_BWD
  WMRL_CUDA_OK(cudaGetLastError());
  return 0;
}

// a = tanh(mu/scale)*bound  (actor.py:50); th saved for the backward.
__global__ void tanh_squash_kernel(const float* mu, const float* bound, float scale, float* act,
                                   long long lda, float* th, int Bn, int A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Bn * A) return;
  const int b = i / A, j = i % A;
  const float t = tanhf(mu[i] / scale);
  th[i] = t;
  act[b * lda + j] = t * bound[j];
}
__global__ void tanh_squash_bwd_kernel(const float* g_a, long long ldg, const float* th,
                                       const float* bound, float scale, float* g_mu, int Bn, int A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Bn * A) return;
  const int b = i / A, j = i % A;
  const float t = th[i];
  g_mu[i] = g_a[b * ldg + j] * bound[j] * (1.f - t * t) / scale;
}

// torch.nn.GRUCell gates (rssm.py:66): r,z,n order; b_hn inside r*(.)
__global__ void gru_fwd_kernel(const float* __restrict__ gi, const float* __restrict__ gh,
                               const float* __restrict__ hprev, long long ldh, float* hnew,
                               long long ldn, float* r_, float* z_, float* n_, float* hn_, int Bn,
                               int D) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)Bn * D) return;
  const int b = (int)(i / D), j = (int)(i % D);
  const float* gib = gi + (long long)b * 3 * D;
  const float* ghb = gh + (long long)b * 3 * D;
  const float r = sigmoid_f(gib[j] + ghb[j]);
  const float z = sigmoid_f(gib[D + j] + ghb[D + j]);
  const float hn = ghb[2 * D + j];
  const float n = tanhf(gib[2 * D + j] + r * hn);
  const float h = hprev[b * ldh + j];
  hnew[b * ldn + j] = (1.f - z) * n + z * h;
  r_[i] = r; z_[i] = z; n_[i] = n; hn_[i] = hn;
}
__global__ void gru_bwd_kernel(const float* __restrict__ g_h, const float* __restrict__ r_,
                               const float* __restrict__ z_, const float* __restrict__ n_,
                               const float* __restrict__ hn_, const float* __restrict__ hprev,
                               long long ldh, float* g_gi, float* g_gh, float* g_hprev, int Bn,
                               int D) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)Bn * D) return;
  const int b = (int)(i / D), j = (int)(i % D);
  const float g = g_h[i], r = r_[i], z = z_[i], n = n_[i], hn = hn_[i];
  const float h = hprev[b * ldh + j];
  const float g_n = g * (1.f - z);
  const float g_z = g * (h - n);
  const float g_an = g_n * (1.f - n * n);
  const float g_ar = g_an * hn * r * (1.f - r);
  const float g_az = g_z * z * (1.f - z);
  float* gib = g_gi + (long long)b * 3 * D;
  float* ghb = g_gh + (long long)b * 3 * D;
  gib[j] = g_ar; gib[D + j] = g_az; gib[2 * D + j] = g_an;
  ghb[j] = g_ar; ghb[D + j] = g_az; ghb[2 * D + j] = g_an * r;
  g_hprev[i] = g * z;
}

// g *= elu'(.) given the post-activation y:  y>0 ? 1 : y+1
__global__ void elu_bwd_kernel(float* g, const float* y, long long n) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = y[i];
  g[i] *= (v > 0.f ? 1.f : v + 1.f);
}

// out[j] += sum_m dY[m*ld + j]
// out[j] += sum_m dY[m][j].  Block = 32 columns x 8 row phases; a block covers `rows_per` rows with
// four independent loads in flight per thread, folds the 8 phases in shared memory and adds once.
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ dY, long long ld, int M, int N,
                                                     int rows_per, float* out) {
  __shared__ float part[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + tx;
  const int m0 = blockIdx.y * rows_per, m1 = min(M, m0 + rows_per);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (j < N) {
    const float* p = dY + j;
    int m = m0 + ty;
    for (; m + 24 < m1; m += 32) {
      s0 += p[(long long)m * ld];
      s1 += p[(long long)(m + 8) * ld];
      s2 += p[(long long)(m + 16) * ld];
      s3 += p[(long long)(m + 24) * ld];
    }
    for (; m < m1; m += 8) s0 += p[(long long)m * ld];
  }
  part[ty][tx] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (ty == 0 && j < N) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += part[i][tx];
    atomicAdd(&out[j], s);
  }
}

// continuous head (rssm.py:179-181)
__global__ void sample_cont_kernel(const float* out, const float* eps, float* stoch, long long lds,
                                   float* mean, float* stdv, long long ldm, int Bn, int S) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Bn * S) return;
  const int b = i / S, j = i % S;
  const float m = out[(long long)b * 2 * S + j];
  const float raw = out[(long long)b * 2 * S + S + j];
  // F.softplus (beta 1, threshold 20)
  const float sp = raw > 20.f ? raw : log1pf(expf(raw));
  const float sd = sp + 0.1f;
  mean[b * ldm + j] = m;
  stdv[b * ldm + j] = sd;
  stoch[b * lds + j] = m + sd * eps[i];
}
__global__ void sample_cont_bwd_kernel(const float* g_stoch, const float* g_mean, const float* g_std,
                                       long long ldg, const float* eps, const float* stdv,
                                       long long ldm, float* g_out, int Bn, int S) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Bn * S) return;
  const int b = i / S, j = i % S;
  const float gs = g_stoch[i];
  const float gm = (g_mean ? g_mean[b * ldg + j] : 0.f) + gs;
  const float gsd = (g_std ? g_std[b * ldg + j] : 0.f) + gs * eps[i];
  const float sp = stdv[b * ldm + j] - 0.1f;
  // d softplus / d raw = sigmoid(raw) = 1 - exp(-softplus(raw))
  const float dsp = 1.f - expf(-sp);
  g_out[(long long)b * 2 * S + j] = gm;
  g_out[(long long)b * 2 * S + S + j] = gsd * dsp;
}

// discrete head: one warp per (row, group).  logits -> Categorical.probs
// (normalise by logsumexp, then softmax: torch.distributions order) ->
// index = first argmax of p/q -> one-hot (state/discrete.py:27-32).
__global__ void sample_disc_kernel(const float* out, const float* q, float* logits, long long ldl,
                                   float* stoch, long long lds, int32_t* idx, long long ldi, int Bn,
                                   int C, int S) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= Bn * C) return;
  const int b = warp / C, c = warp % C;
  const float* l = out + (long long)b * C * S + (long long)c * S;
  const float* qq = q + (long long)b * C * S + (long long)c * S;
  float mx = -INFINITY;
  for (int j = lane; j < S; j += 32) mx = fmaxf(mx, l[j]);
  mx = warp_max(mx);
  float se = 0.f;
  for (int j = lane; j < S; j += 32) se += expf(l[j] - mx);
  const float lse = mx + logf(warp_sum(se));
  // softmax of the normalised logits
  float mx2 = -INFINITY;
  for (int j = lane; j < S; j += 32) mx2 = fmaxf(mx2, l[j] - lse);
  mx2 = warp_max(mx2);
  float se2 = 0.f;
  for (int j = lane; j < S; j += 32) se2 += expf((l[j] - lse) - mx2);
  se2 = warp_sum(se2);
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int j = lane; j < S; j += 32) {
    const float p = expf((l[j] - lse) - mx2) / se2;
    const float v = p / qq[j];
    if (v > best) { best = v; bi = j; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  for (int j = lane; j < S; j += 32) {
    logits[b * ldl + (long long)c * S + j] = l[j];
    stoch[b * lds + (long long)c * S + j] = (j == bi) ? 1.f : 0.f;
  }
  if (idx && lane == 0) idx[b * ldi + c] = bi;
}
// straight-through backward: dlogits = p*(g - sum p g) + g_logits
__global__ void sample_disc_bwd_kernel(const float* g_stoch, const float* g_logits, long long ldg,
                                       const float* logits, long long ldl, float* g_out, int Bn,
                                       int C, int S) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= Bn * C) return;
  const int b = warp / C, c = warp % C;
  const float* l = logits + b * ldl + (long long)c * S;
  const float* g = g_stoch + (long long)b * C * S + (long long)c * S;
  float mx = -INFINITY;
  for (int j = lane; j < S; j += 32) mx = fmaxf(mx, l[j]);
  mx = warp_max(mx);
  float se = 0.f, sg = 0.f;
  for (int j = lane; j < S; j += 32) {
    const float e = expf(l[j] - mx);
    se += e;
    sg += e * g[j];
  }
  se = warp_sum(se);
  sg = warp_sum(sg) / se;
  for (int j = lane; j < S; j += 32) {
    const float p = expf(l[j] - mx) / se;
    float v = p * (g[j] - sg);
    if (g_logits) v += g_logits[b * ldg + (long long)c * S + j];
    g_out[(long long)b * C * S + (long long)c * S + j] = v;
  }
}

// ---------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------
static inline int blocks_for(long long n, int t = 256) { return (int)((n + t - 1) / t); }

static int combine(cudaStream_t st, const float* a, long long lda, const float* c, long long ldc,
                   float* out, long long ldo, int Bn, int W) {
  if (Bn * (long long)W == 0) return 0;
  int nb = blocks_for((long long)Bn * W);
  if (nb > 4096) nb = 4096;
  combine_kernel<<<nb, 256, 0, st>>>(a, lda, c, ldc, out, ldo, Bn, W);
  WMRL_CUDA_OK(cudaGetLastError());
  return 0;
}
static int bgrad(cudaStream_t st, const float* dY, long long ld, int M, int N, float* db) {
  if (!db) return 0;
  const int gx = (N + 31) / 32;
  int gy = (M + 63) / 64;                      // >= 64 rows per block ...
  const int want = (148 * 8 + gx - 1) / gx;    // ... but no more blocks than fill the GPU a few times
  if (gy > want) gy = want;
  if (gy < 1) gy = 1;
  const int rows_per = (M + gy - 1) / gy;
  gy = (M + rows_per - 1) / rows_per;
  colsum_kernel<<<dim3(gx, gy), 256, 0, st>>>(dY, ld, M, N, rows_per, db);
  WMRL_CUDA_OK(cudaGetLastError());
  return 0;
}

#define TRY(x)           \
  do {                   \
    int _r = (x);        \
    if (_r) return _r;   \
  } while (0)

// saved-activation layout --------------------------------------------------
struct DynSaved { float *x, *r, *z, *n, *hn; };
struct ImSaved {
  float* xhat[WMRL_MAX_ACTOR_LAYERS];
  float* y[WMRL_MAX_ACTOR_LAYERS];
  float* rstd[WMRL_MAX_ACTOR_LAYERS];
  float* th;
  DynSaved dyn;
  float* hid;
};
static size_t im_step_floats(const Dims& d) {
  return (size_t)d.B * ((size_t)d.La * (2 * d.Ha + 1) + d.A + 5 * (size_t)d.D + d.Hd);
}
static ImSaved im_saved_at(const Dims& d, float* base) {
  ImSaved s;
  float* p = base;
  const size_t B = d.B;
  for (int l = 0; l < d.La; ++l) {
    s.xhat[l] = p; p += B * d.Ha;
    s.y[l] = p; p += B * d.Ha;
    s.rstd[l] = p; p += B;
  }
  s.th = p; p += B * d.A;
  s.dyn.x = p; p += B * d.D;
  s.dyn.r = p; p += B * d.D;
  s.dyn.z = p; p += B * d.D;
  s.dyn.n = p; p += B * d.D;
  s.dyn.hn = p; p += B * d.D;
  s.hid = p;
  return s;
}
struct ObSaved { DynSaved dyn; float *hid_pri, *hid_post; };
static size_t ob_step_floats(const Dims& d) { return (size_t)d.B * (5 * (size_t)d.D + 2 * (size_t)d.Hd); }
static ObSaved ob_saved_at(const Dims& d, float* base) {
  ObSaved s;
  float* p = base;
  const size_t B = d.B;
  s.dyn.x = p; p += B * d.D;
  s.dyn.r = p; p += B * d.D;
  s.dyn.z = p; p += B * d.D;
  s.dyn.n = p; p += B * d.D;
  s.dyn.hn = p; p += B * d.D;
  s.hid_pri = p; p += B * d.Hd;
  s.hid_post = p;
  return s;
}

size_t f32_saved_bytes(const Dims& d, int op) {
  if (op == 0) return im_step_floats(d) * (size_t)d.T * sizeof(float);
  return ob_step_floats(d) * (size_t)(d.T > 0 ? d.T - 1 : 0) * sizeof(float);
}
static size_t bwd_ws_floats(const Dims& d) {
  // g_h, carry_h, g_x (3D) ; g_s, carry_s (2 S_in); g_out (P); g_hid (Hd); g_gi,g_gh (6D);
  // g_a, g_mu (2A); g_y, g_lin (2Ha)
  return (size_t)d.B * (9 * (size_t)d.D + 2 * (size_t)d.S_in + d.P + d.Hd + 2 * (size_t)d.A +
                        2 * (size_t)d.Ha);
}
static size_t fwd_ws_floats(const Dims& d) {
  // tmp_lin (Ha) ; gi, gh (6D); out (P); mu (A)
  return (size_t)d.B * ((size_t)d.Ha + 6 * (size_t)d.D + d.P + d.A);
}
size_t f32_workspace_bytes(const Dims& d, int op) {
  const size_t step = (op == 0) ? im_step_floats(d) : ob_step_floats(d);
  size_t f = fwd_ws_floats(d) + step;  // single-step scratch when nothing is saved
  const size_t b = bwd_ws_floats(d);
  return (f > b ? f : b) * sizeof(float) + 256;
}

// ---------------------------------------------------------------------------
// shared step pieces
// ---------------------------------------------------------------------------
// x = ELU(W_e [s|a] + b); h' = GRUCell(x, h)   (rssm.py:64-66)
static int dyn_fwd(cudaStream_t st, const Dims& d, const wmrl_rssm_params* w, const float* s,
                   long long lds, const float* a, long long lda, const float* h, long long ldh,
                   float* hnew, long long ldn, const DynSaved& sv, float* gi, float* gh) {
  TRY(linear(st, s, lds, d.S_in, a, lda, d.A, w->embed_w, w->embed_b, sv.x, d.D, d.B, d.D, 1));
  TRY(linear(st, sv.x, d.D, d.D, nullptr, 0, 0, w->w_ih, w->b_ih, gi, 3 * d.D, d.B, 3 * d.D, 0));
  TRY(linear(st, h, ldh, d.D, nullptr, 0, 0, w->w_hh, w->b_hh, gh, 3 * d.D, d.B, 3 * d.D, 0));
  gru_fwd_kernel<<<blocks_for((long long)d.B * d.D), 256, 0, st>>>(gi, gh, h, ldh, hnew, ldn, sv.r,
                                                                   sv.z, sv.n, sv.hn, d.B, d.D);
  WMRL_CUDA_OK(cudaGetLastError());
  return 0;
}

// out = W2 ELU(W0 [in1|in2] + b0) + b2     (rssm.py:33-42, 67, 81)
static int head_fwd(cudaStream_t st, const Dims& d, const float* in1, long long ld1, int K1,
                    const float* in2, long long ld2, int K2, const float* W0, const float* b0,
                    const float* W2, const float* b2, float* hid, float* out) {
  TRY(linear(st, in1, ld1, K1, in2, ld2, K2, W0, b0, hid, d.Hd, d.B, d.Hd, 1));
  TRY(linear(st, hid, d.Hd, d.Hd, nullptr, 0, 0, W2, b2, out, d.P, d.B, d.P, 0));
  return 0;
}

static int sample_fwd(cudaStream_t st, const Dims& d, const float* out, const float* noise_t,
                      float* stoch, long long lds, float* da, float* db, long long ldd,
                      int32_t* idx, long long ldi) {
  if (d.variant == WMRL_CONTINUOUS) {
    sample_cont_kernel<<<blocks_for((long long)d.B * d.S), 256, 0, st>>>(out, noise_t, stoch, lds,
                                                                         da, db, ldd, d.B, d.S);
  } else {
    const long long nthreads = (long long)d.B * d.C * 32;
    sample_disc_kernel<<<blocks_for(nthreads), 256, 0, st>>>(out, noise_t, da, ldd, stoch, lds, idx,
                                                             ldi, d.B, d.C, d.S);
  }
  WMRL_CUDA_OK(cudaGetLastError());
  return 0;
}

// g_out from the gradients arriving at (stoch, dist) of one sampled state.
static int sample_bwd(cudaStream_t st, const Dims& d, const float* g_stoch, const float* g_da,
                      const float* g_db, long long ldg, const float* noise_t, const float* da,
                      const float* db, long long ldd, float* g_out) {
  if (d.variant == WMRL_CONTINUOUS) {
    sample_cont_bwd_kernel<<<blocks_for((long long)d.B * d.S), 256, 0, st>>>(
        g_stoch, g_da, g_db, ldg, noise_t, db, ldd, g_out, d.B, d.S);
  } else {
    const long long nthreads = (long long)d.B * d.C * 32;
    sample_disc_bwd_kernel<<<blocks_for(nthreads), 256, 0, st>>>(g_stoch, g_da, ldg, da, ldd, g_out,
                                                                 d.B, d.C, d.S);
  }
  WMRL_CUDA_OK(cudaGetLastError());
  return 0;
}

// backward of head_fwd.  g_in1 accumulates (+=); g_in2 (if any) is overwritten.
static int head_bwd(cudaStream_t st, const Dims& d, float* g_out, const float* hid, const float* in1,
                    long long ld1, int K1, const float* in2, long long ld2, int K2, const float* W0,
                    const float* W2, float* gW0, float* gb0, float* gW2, float* gb2, float* g_hid,
                    float* g_in1, long long ldg1, float* g_in2, long long ldg2) {
  TRY(wgrad(st, g_out, d.P, d.P, hid, d.Hd, d.Hd, gW2, d.Hd, 0, d.B));
  TRY(bgrad(st, g_out, d.P, d.B, d.P, gb2));
  TRY(dgrad(st, g_out, d.P, d.P, W2, d.Hd, 0, d.Hd, g_hid, d.Hd, d.B, 0));
  elu_bwd_kernel<<<blocks_for((long long)d.B * d.Hd), 256, 0, st>>>(g_hid, hid, (long long)d.B * d.Hd);
  WMRL_CUDA_OK(cudaGetLastError());
  TRY(wgrad(st, g_hid, d.Hd, d.Hd, in1, ld1, K1, gW0, K1 + K2, 0, d.B));
  if (K2 > 0) TRY(wgrad(st, g_hid, d.Hd, d.Hd, in2, ld2, K2, gW0, K1 + K2, K1, d.B));
  TRY(bgrad(st, g_hid, d.Hd, d.B, d.Hd, gb0));
  TRY(dgrad(st, g_hid, d.Hd, d.Hd, W0, K1 + K2, 0, K1, g_in1, ldg1, d.B, 1));
  if (K2 > 0 && g_in2) TRY(dgrad(st, g_hid, d.Hd, d.Hd, W0, K1 + K2, K1, K2, g_in2, ldg2, d.B, 0));
  return 0;
}

// backward of dyn_fwd.  g_h: grad wrt h' [B,D] (dense).  Produces carry_h (grad
// wrt h), carry_s (grad wrt s) and g_a (grad wrt action, row stride ldga).
static int dyn_bwd(cudaStream_t st, const Dims& d, const wmrl_rssm_params* w, const wmrl_rssm_grads* gw,
                   const float* g_h, const DynSaved& sv, const float* s, long long lds,
                   const float* a, long long lda, const float* h, long long ldh, float* g_gi,
                   float* g_gh, float* g_x, float* carry_h, float* carry_s, float* g_a,
                   long long ldga) {
  gru_bwd_kernel<<<blocks_for((long long)d.B * d.D), 256, 0, st>>>(g_h, sv.r, sv.z, sv.n, sv.hn, h,
                                                                   ldh, g_gi, g_gh, carry_h, d.B, d.D);
  WMRL_CUDA_OK(cudaGetLastError());
  TRY(dgrad(st, g_gi, 3 * d.D, 3 * d.D, w->w_ih, d.D, 0, d.D, g_x, d.D, d.B, 0));
  TRY(dgrad(st, g_gh, 3 * d.D, 3 * d.D, w->w_hh, d.D, 0, d.D, carry_h, d.D, d.B, 1));
  if (gw) {
    TRY(wgrad(st, g_gi, 3 * d.D, 3 * d.D, sv.x, d.D, d.D, gw->w_ih, d.D, 0, d.B));
    TRY(wgrad(st, g_gh, 3 * d.D, 3 * d.D, h, ldh, d.D, gw->w_hh, d.D, 0, d.B));
    TRY(bgrad(st, g_gi, 3 * d.D, d.B, 3 * d.D, gw->b_ih));
    TRY(bgrad(st, g_gh, 3 * d.D, d.B, 3 * d.D, gw->b_hh));
  }
  elu_bwd_kernel<<<blocks_for((long long)d.B * d.D), 256, 0, st>>>(g_x, sv.x, (long long)d.B * d.D);
  WMRL_CUDA_OK(cudaGetLastError());
  const int Ke = d.S_in + d.A;
  if (gw) {
    TRY(wgrad(st, g_x, d.D, d.D, s, lds, d.S_in, gw->embed_w, Ke, 0, d.B));
    TRY(wgrad(st, g_x, d.D, d.D, a, lda, d.A, gw->embed_w, Ke, d.S_in, d.B));
    TRY(bgrad(st, g_x, d.D, d.B, d.D, gw->embed_b));
  }
  TRY(dgrad(st, g_x, d.D, d.D, w->embed_w, Ke, 0, d.S_in, carry_s, d.S_in, d.B, 0));
  if (g_a) TRY(dgrad(st, g_x, d.D, d.D, w->embed_w, Ke, d.S_in, d.A, g_a, ldga, d.B, 0));
  return 0;
}

// ---------------------------------------------------------------------------
// imagine
// ---------------------------------------------------------------------------
int f32_imagine_fwd(const Dims& d, const wmrl_rssm_params* w, const wmrl_actor_params* aw,
                    const float* deter0, const float* stoch0, const float* noise, float* deter_seq,
                    float* stoch_seq, float* dist_a, float* dist_b, float* actions, int32_t* idx,
                    float* saved, float* ws, cudaStream_t st) {
  const int B = d.B, H = d.T, T1 = H + 1;
  const long long ldD = (long long)T1 * d.D, ldS = (long long)T1 * d.S_in;
  const long long ldP = (long long)T1 * (d.variant == WMRL_CONTINUOUS ? d.S : d.S_in);
  const long long ldA = (long long)H * d.A, ldI = (long long)H * d.C;
  const size_t step = im_step_floats(d);
  float* p = ws;
  float* tmp = p; p += (size_t)B * d.Ha;
  float* gi = p; p += (size_t)B * 3 * d.D;
  float* gh = p; p += (size_t)B * 3 * d.D;
  float* out = p; p += (size_t)B * d.P;
  float* mu = p; p += (size_t)B * d.A;
  float* scratch_saved = p;

  TRY(combine(st, deter0, d.D, nullptr, 0, deter_seq, ldD, B, d.D));
  TRY(combine(st, stoch0, d.S_in, nullptr, 0, stoch_seq, ldS, B, d.S_in));
  for (int t = 0; t < H; ++t) {
    ImSaved sv = im_saved_at(d, saved ? saved + (size_t)t * step : scratch_saved);
    const float* h_t = deter_seq + (size_t)t * d.D;
    const float* s_t = stoch_seq + (size_t)t * d.S_in;
    float* a_t = actions + (size_t)t * d.A;
    // actor (actor.py:47-52), input detached (rssm.py:135)
    for (int l = 0; l < d.La; ++l) {
      if (l == 0)
        TRY(linear(st, h_t, ldD, d.D, s_t, ldS, d.S_in, aw->lin_w[0], aw->lin_b[0], tmp, d.Ha, B, d.Ha, 0));
      else
        TRY(linear(st, sv.y[l - 1], d.Ha, d.Ha, nullptr, 0, 0, aw->lin_w[l], aw->lin_b[l], tmp, d.Ha, B, d.Ha, 0));
      TRY(ln_act_fwd(st, 0, tmp, aw->ln_g[l], aw->ln_b[l], sv.xhat[l], sv.rstd[l], sv.y[l], B, d.Ha));
    }
    TRY(linear(st, sv.y[d.La - 1], d.Ha, d.Ha, nullptr, 0, 0, aw->mu_w, aw->mu_b, mu, d.A, B, d.A, 0));
    tanh_squash_kernel<<<blocks_for((long long)B * d.A), 256, 0, st>>>(mu, aw->bound, d.action_scale,
                                                                       a_t, ldA, sv.th, B, d.A);
    WMRL_CUDA_OK(cudaGetLastError());
    // prior (rssm.py:53-68)
    float* h_n = deter_seq + (size_t)(t + 1) * d.D;
    TRY(dyn_fwd(st, d, w, s_t, ldS, a_t, ldA, h_t, ldD, h_n, ldD, sv.dyn, gi, gh));
    TRY(head_fwd(st, d, h_n, ldD, d.D, nullptr, 0, 0, w->prior0_w, w->prior0_b, w->prior2_w,
                 w->prior2_b, sv.hid, out));
    const size_t pw = (d.variant == WMRL_CONTINUOUS) ? d.S : d.S_in;
    TRY(sample_fwd(st, d, out, noise + (size_t)t * B * d.S_in, stoch_seq + (size_t)(t + 1) * d.S_in,
                   ldS, dist_a + (size_t)(t + 1) * pw, dist_b ? dist_b + (size_t)(t + 1) * pw : nullptr,
                   ldP, idx ? idx + (size_t)t * d.C : nullptr, ldI));
  }
  return 0;
}

int f32_imagine_bwd(const Dims& d, const wmrl_rssm_params* w, const wmrl_actor_params* aw,
                    const float* noise, const float* deter_seq, const float* stoch_seq,
                    const float* dist_a, const float* dist_b, const float* actions,
                    const float* saved, const float* g_deter_seq, const float* g_stoch_seq,
                    const float* g_dist_a, const float* g_dist_b, const wmrl_actor_grads* gaw,
                    const wmrl_rssm_grads* gw, float* g_deter0, float* g_stoch0, float* ws,
                    cudaStream_t st) {
  const int B = d.B, H = d.T, T1 = H + 1;
  const long long ldD = (long long)T1 * d.D, ldS = (long long)T1 * d.S_in;
  const size_t pw = (d.variant == WMRL_CONTINUOUS) ? d.S : d.S_in;
  const long long ldP = (long long)T1 * pw;
  const long long ldA = (long long)H * d.A;
  const size_t step = im_step_floats(d);
  if (!(d.flags & WMRL_FLAG_WM_WGRAD)) gw = nullptr;
  float* p = ws;
  float* g_h = p; p += (size_t)B * d.D;
  float* carry_h = p; p += (size_t)B * d.D;
  float* g_x = p; p += (size_t)B * d.D;
  float* g_s = p; p += (size_t)B * d.S_in;
  float* carry_s = p; p += (size_t)B * d.S_in;
  float* g_out = p; p += (size_t)B * d.P;
  float* g_hid = p; p += (size_t)B * d.Hd;
  float* g_gi = p; p += (size_t)B * 3 * d.D;
  float* g_gh = p; p += (size_t)B * 3 * d.D;
  float* g_a = p; p += (size_t)B * d.A;
  float* g_mu = p; p += (size_t)B * d.A;
  float* g_y = p; p += (size_t)B * d.Ha;
  float* g_lin = p; p += (size_t)B * d.Ha;

  WMRL_CUDA_OK(cudaMemsetAsync(carry_h, 0, sizeof(float) * (size_t)B * d.D, st));
  WMRL_CUDA_OK(cudaMemsetAsync(carry_s, 0, sizeof(float) * (size_t)B * d.S_in, st));
  for (int t = H - 1; t >= 0; --t) {
    ImSaved sv = im_saved_at(d, const_cast<float*>(saved) + (size_t)t * step);
    const float* h_t = deter_seq + (size_t)t * d.D;
    const float* s_t = stoch_seq + (size_t)t * d.S_in;
    const float* h_n = deter_seq + (size_t)(t + 1) * d.D;
    const float* a_t = actions + (size_t)t * d.A;
    TRY(combine(st, g_deter_seq ? g_deter_seq + (size_t)(t + 1) * d.D : nullptr, ldD, carry_h, d.D,
                g_h, d.D, B, d.D));
    TRY(combine(st, g_stoch_seq ? g_stoch_seq + (size_t)(t + 1) * d.S_in : nullptr, ldS, carry_s,
                d.S_in, g_s, d.S_in, B, d.S_in));
    TRY(sample_bwd(st, d, g_s, g_dist_a ? g_dist_a + (size_t)(t + 1) * pw : nullptr,
                   g_dist_b ? g_dist_b + (size_t)(t + 1) * pw : nullptr, ldP,
                   noise + (size_t)t * B * d.S_in, dist_a + (size_t)(t + 1) * pw,
                   dist_b ? dist_b + (size_t)(t + 1) * pw : nullptr, ldP, g_out));
    TRY(head_bwd(st, d, g_out, sv.hid, h_n, ldD, d.D, nullptr, 0, 0, w->prior0_w, w->prior2_w,
                 gw ? gw->prior0_w : nullptr, gw ? gw->prior0_b : nullptr, gw ? gw->prior2_w : nullptr,
                 gw ? gw->prior2_b : nullptr, g_hid, g_h, d.D, nullptr, 0));
    TRY(dyn_bwd(st, d, w, gw, g_h, sv.dyn, s_t, ldS, a_t, ldA, h_t, ldD, g_gi, g_gh, g_x, carry_h,
                carry_s, g_a, d.A));
    // actor backward: a = tanh(mu/scale)*bound
    tanh_squash_bwd_kernel<<<blocks_for((long long)B * d.A), 256, 0, st>>>(g_a, d.A, sv.th, aw->bound,
                                                                           d.action_scale, g_mu, B, d.A);
    WMRL_CUDA_OK(cudaGetLastError());
    TRY(wgrad(st, g_mu, d.A, d.A, sv.y[d.La - 1], d.Ha, d.Ha, gaw->mu_w, d.Ha, 0, B));
    TRY(bgrad(st, g_mu, d.A, B, d.A, gaw->mu_b));
    TRY(dgrad(st, g_mu, d.A, d.A, aw->mu_w, d.Ha, 0, d.Ha, g_y, d.Ha, B, 0));
    for (int l = d.La - 1; l >= 0; --l) {
      TRY(ln_act_bwd(st, 0, g_y, sv.xhat[l], sv.rstd[l], aw->ln_g[l], aw->ln_b[l], g_lin, gaw->ln_g[l], gaw->ln_b[l], B,
                     d.Ha));
      if (l == 0) {
        TRY(wgrad(st, g_lin, d.Ha, d.Ha, h_t, ldD, d.D, gaw->lin_w[0], d.F, 0, B));
        TRY(wgrad(st, g_lin, d.Ha, d.Ha, s_t, ldS, d.S_in, gaw->lin_w[0], d.F, d.D, B));
      } else {
        TRY(wgrad(st, g_lin, d.Ha, d.Ha, sv.y[l - 1], d.Ha, d.Ha, gaw->lin_w[l], d.Ha, 0, B));
        TRY(dgrad(st, g_lin, d.Ha, d.Ha, aw->lin_w[l], d.Ha, 0, d.Ha, g_y, d.Ha, B, 0));
      }
      TRY(bgrad(st, g_lin, d.Ha, B, d.Ha, gaw->lin_b[l]));
    }
  }
  if (g_deter0) TRY(combine(st, g_deter_seq, ldD, carry_h, d.D, g_deter0, d.D, B, d.D));
  if (g_stoch0) TRY(combine(st, g_stoch_seq, ldS, carry_s, d.S_in, g_stoch0, d.S_in, B, d.S_in));
  return 0;
}

// ---------------------------------------------------------------------------
// observe
// ---------------------------------------------------------------------------
int f32_observe_fwd(const Dims& d, const wmrl_rssm_params* w, const float* obs, const float* act,
                    const float* noise_pri, const float* noise_post, float* pri_deter,
                    float* pri_stoch, float* pri_da, float* pri_db, float* post_deter,
                    float* post_stoch, float* post_da, float* post_db, float* saved, float* ws,
                    cudaStream_t st) {
  const int B = d.B, T = d.T;
  const long long ldD = (long long)T * d.D, ldS = (long long)T * d.S_in;
  const size_t pw = (d.variant == WMRL_CONTINUOUS) ? d.S : d.S_in;
  const long long ldP = (long long)T * pw;
  const long long ldA = (long long)T * d.A, ldE = (long long)T * d.E;
  const size_t step = ob_step_floats(d);
  float* p = ws;
  p += (size_t)B * d.Ha;  // keep the imagine layout (tmp unused here)
  float* gi = p; p += (size_t)B * 3 * d.D;
  float* gh = p; p += (size_t)B * 3 * d.D;
  float* out = p; p += (size_t)B * d.P;
  p += (size_t)B * d.A;
  float* scratch_saved = p;
  // zero initial state (rssm.py:166-172 / 215-220)
  TRY(combine(st, nullptr, 0, nullptr, 0, pri_deter, ldD, B, d.D));
  TRY(combine(st, nullptr, 0, nullptr, 0, post_deter, ldD, B, d.D));
  TRY(combine(st, nullptr, 0, nullptr, 0, pri_stoch, ldS, B, d.S_in));
  TRY(combine(st, nullptr, 0, nullptr, 0, post_stoch, ldS, B, d.S_in));
  for (int t = 0; t + 1 < T; ++t) {
    ObSaved sv = ob_saved_at(d, saved ? saved + (size_t)t * step : scratch_saved);
    const float* h_t = post_deter + (size_t)t * d.D;
    const float* s_t = post_stoch + (size_t)t * d.S_in;
    float* h_n = pri_deter + (size_t)(t + 1) * d.D;
    TRY(dyn_fwd(st, d, w, s_t, ldS, act + (size_t)t * d.A, ldA, h_t, ldD, h_n, ldD, sv.dyn, gi, gh));
    TRY(combine(st, h_n, ldD, nullptr, 0, post_deter + (size_t)(t + 1) * d.D, ldD, B, d.D));
    TRY(head_fwd(st, d, h_n, ldD, d.D, nullptr, 0, 0, w->prior0_w, w->prior0_b, w->prior2_w,
                 w->prior2_b, sv.hid_pri, out));
    TRY(sample_fwd(st, d, out, noise_pri + (size_t)t * B * d.S_in, pri_stoch + (size_t)(t + 1) * d.S_in,
                   ldS, pri_da + (size_t)(t + 1) * pw, pri_db ? pri_db + (size_t)(t + 1) * pw : nullptr,
                   ldP, nullptr, 0));
    TRY(head_fwd(st, d, h_n, ldD, d.D, obs + (size_t)(t + 1) * d.E, ldE, d.E, w->post0_w, w->post0_b,
                 w->post2_w, w->post2_b, sv.hid_post, out));
    TRY(sample_fwd(st, d, out, noise_post + (size_t)t * B * d.S_in,
                   post_stoch + (size_t)(t + 1) * d.S_in, ldS, post_da + (size_t)(t + 1) * pw,
                   post_db ? post_db + (size_t)(t + 1) * pw : nullptr, ldP, nullptr, 0));
  }
  return 0;
}

int f32_observe_bwd(const Dims& d, const wmrl_rssm_params* w, const float* obs, const float* act,
                    const float* noise_pri, const float* noise_post, const float* pri_deter,
                    const float* pri_stoch, const float* pri_da, const float* pri_db,
                    const float* post_deter, const float* post_stoch, const float* post_da,
                    const float* post_db, const float* saved, const float* g_pri_deter,
                    const float* g_pri_stoch, const float* g_pri_da, const float* g_pri_db,
                    const float* g_post_deter, const float* g_post_stoch, const float* g_post_da,
                    const float* g_post_db, const wmrl_rssm_grads* gw, float* g_obs, float* g_act,
                    float* ws, cudaStream_t st) {
  (void)pri_stoch;
  const int B = d.B, T = d.T;
  const long long ldD = (long long)T * d.D, ldS = (long long)T * d.S_in;
  const size_t pw = (d.variant == WMRL_CONTINUOUS) ? d.S : d.S_in;
  const long long ldP = (long long)T * pw;
  const long long ldA = (long long)T * d.A, ldE = (long long)T * d.E;
  const size_t step = ob_step_floats(d);
  float* p = ws;
  float* g_h = p; p += (size_t)B * d.D;
  float* carry_h = p; p += (size_t)B * d.D;
  float* g_x = p; p += (size_t)B * d.D;
  float* g_s = p; p += (size_t)B * d.S_in;
  float* carry_s = p; p += (size_t)B * d.S_in;
  float* g_out = p; p += (size_t)B * d.P;
  float* g_hid = p; p += (size_t)B * d.Hd;
  float* g_gi = p; p += (size_t)B * 3 * d.D;
  float* g_gh = p; p += (size_t)B * 3 * d.D;

  WMRL_CUDA_OK(cudaMemsetAsync(carry_h, 0, sizeof(float) * (size_t)B * d.D, st));
  WMRL_CUDA_OK(cudaMemsetAsync(carry_s, 0, sizeof(float) * (size_t)B * d.S_in, st));
  if (g_obs) WMRL_CUDA_OK(cudaMemsetAsync(g_obs, 0, sizeof(float) * (size_t)B * T * d.E, st));
  if (g_act) WMRL_CUDA_OK(cudaMemsetAsync(g_act, 0, sizeof(float) * (size_t)B * T * d.A, st));
  for (int t = T - 2; t >= 0; --t) {
    ObSaved sv = ob_saved_at(d, const_cast<float*>(saved) + (size_t)t * step);
    const float* h_t = post_deter + (size_t)t * d.D;
    const float* s_t = post_stoch + (size_t)t * d.S_in;
    const float* h_n = pri_deter + (size_t)(t + 1) * d.D;
    // h' feeds both sequences' deter fields (rssm.py:81-82 keeps deter_state)
    TRY(combine(st, g_pri_deter ? g_pri_deter + (size_t)(t + 1) * d.D : nullptr, ldD, carry_h, d.D,
                g_h, d.D, B, d.D));
    if (g_post_deter)
      TRY(combine(st, g_post_deter + (size_t)(t + 1) * d.D, ldD, g_h, d.D, g_h, d.D, B, d.D));
    // posterior branch (its sample is the one fed forward)
    TRY(combine(st, g_post_stoch ? g_post_stoch + (size_t)(t + 1) * d.S_in : nullptr, ldS, carry_s,
                d.S_in, g_s, d.S_in, B, d.S_in));
    TRY(sample_bwd(st, d, g_s, g_post_da ? g_post_da + (size_t)(t + 1) * pw : nullptr,
                   g_post_db ? g_post_db + (size_t)(t + 1) * pw : nullptr, ldP,
                   noise_post + (size_t)t * B * d.S_in, post_da + (size_t)(t + 1) * pw,
                   post_db ? post_db + (size_t)(t + 1) * pw : nullptr, ldP, g_out));
    TRY(head_bwd(st, d, g_out, sv.hid_post, h_n, ldD, d.D, obs + (size_t)(t + 1) * d.E, ldE, d.E,
                 w->post0_w, w->post2_w, gw->post0_w, gw->post0_b, gw->post2_w, gw->post2_b, g_hid,
                 g_h, d.D, g_obs ? g_obs + (size_t)(t + 1) * d.E : nullptr, ldE));
    // prior branch (sample drawn and stored, never fed forward)
    TRY(combine(st, g_pri_stoch ? g_pri_stoch + (size_t)(t + 1) * d.S_in : nullptr, ldS, nullptr, 0,
                g_s, d.S_in, B, d.S_in));
    TRY(sample_bwd(st, d, g_s, g_pri_da ? g_pri_da + (size_t)(t + 1) * pw : nullptr,
                   g_pri_db ? g_pri_db + (size_t)(t + 1) * pw : nullptr, ldP,
                   noise_pri + (size_t)t * B * d.S_in, pri_da + (size_t)(t + 1) * pw,
                   pri_db ? pri_db + (size_t)(t + 1) * pw : nullptr, ldP, g_out));
    TRY(head_bwd(st, d, g_out, sv.hid_pri, h_n, ldD, d.D, nullptr, 0, 0, w->prior0_w, w->prior2_w,
                 gw->prior0_w, gw->prior0_b, gw->prior2_w, gw->prior2_b, g_hid, g_h, d.D, nullptr, 0));
    TRY(dyn_bwd(st, d, w, gw, g_h, sv.dyn, s_t, ldS, act + (size_t)t * d.A, ldA, h_t, ldD, g_gi,
                g_gh, g_x, carry_h, carry_s, g_act ? g_act + (size_t)t * d.A : nullptr, ldA));
  }
  return 0;
}

// ---------------------------------------------------------------------------
// lambda-return (value_trainer.py:62-134), O(H) scan per row - see
// oracle/rssm_oracle.py::lambda_return_scan for the derivation.
// ---------------------------------------------------------------------------
__global__ void lambda_fwd_kernel(const float* __restrict__ V, const float* __restrict__ r,
                                  const float* __restrict__ dn, float gamma, float lam, int Bn, int H,
                                  float* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= Bn) return;
  const long long o = (long long)b * H;
  float vp = V[o + H - 1] * (1.f - dn[o + H - 1]);
  float Std = vp, M = vp;
  out[o + H - 1] = vp;
  double lp = 1.0;  // lam^(n-2), n = H - i
  for (int i = H - 2; i >= 0; --i) {
    const float rp = r[o + i] * (1.f - dn[o + i]);
    Std = rp + gamma * ((1.f - lam) * vp + lam * Std);
    M = rp + gamma * M;
    const float c = (float)(lp - lp * (double)lam);
    out[o + i] = Std - c * M;
    lp *= (double)lam;
    vp = V[o + i] * (1.f - dn[o + i]);
  }
}
__global__ void lambda_bwd_kernel(const float* __restrict__ g_out, const float* __restrict__ dn,
                                  float gamma, float lam, int Bn, int H, float* __restrict__ g_V,
                                  float* __restrict__ g_r) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= Bn) return;
  const long long o = (long long)b * H;
  float gStd = 0.f, gM = 0.f;
  float gv_next = 0.f;  // contribution to V'_{i} from step i-1
  for (int i = 0; i <= H - 2; ++i) {
    const int n = H - i;
    const float c = (float)(pow((double)lam, n - 2) - pow((double)lam, n - 1));
    const float g = g_out[o + i];
    gStd = g + gamma * lam * gStd;
    gM = -c * g + gamma * gM;
    const float keep = 1.f - dn[o + i];
    g_r[o + i] = (gStd + gM) * keep;
    g_V[o + i] = gv_next * keep;
    gv_next = gamma * (1.f - lam) * gStd;
  }
  const float keep = 1.f - dn[o + H - 1];
  const float tail = (H >= 2) ? (gv_next + gamma * lam * gStd + gamma * gM) : 0.f;
  g_V[o + H - 1] = (g_out[o + H - 1] + tail) * keep;
  g_r[o + H - 1] = 0.f;
}
__global__ void done_mask_kernel(const float* __restrict__ dones, int Bn, int H, float* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= Bn) return;
  const long long o = (long long)b * H;
  float m = 0.f;
  out[o] = 0.f;
  for (int i = 1; i < H; ++i) {
    m = fmaxf(m, dones[o + i - 1] > 0.5f ? 1.f : 0.f);
    out[o + i] = m;
  }
}

}  // namespace wmrl

extern "C" {
int wmrl_lambda_return_fwd(const float* V, const float* r, const float* d, float gamma, float lam,
                           int32_t B, int32_t H, float* out, void* stream) {
  WMRL_REQUIRE(B >= 0 && H >= 1, "lambda_return_fwd: need H >= 1");
  if (B == 0) return 0;
  wmrl::lambda_fwd_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(V, r, d, gamma, lam, B, H, out);
  WMRL_CUDA_OK(cudaGetLastError());
  return 0;
}
int wmrl_lambda_return_bwd(const float* g_out, const float* d, float gamma, float lam, int32_t B,
                           int32_t H, float* g_V, float* g_r, void* stream) {
  WMRL_REQUIRE(B >= 0 && H >= 1, "lambda_return_bwd: need H >= 1");
  if (B == 0) return 0;
  wmrl::lambda_bwd_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(g_out, d, gamma, lam, B, H, g_V, g_r);
  WMRL_CUDA_OK(cudaGetLastError());
  return 0;
}
int wmrl_done_mask(const float* dones, int32_t B, int32_t H, float* out, void* stream) {
  WMRL_REQUIRE(B >= 0 && H >= 1, "done_mask: need H >= 1");
  if (B == 0) return 0;
  wmrl::done_mask_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(dones, B, H, out);
  WMRL_CUDA_OK(cudaGetLastError());
  return 0;
}
int wmrl_ln_act_fwd(int32_t M, int32_t W, const float* x, const float* gamma, const float* beta, int32_t act,
                    float* xhat, float* rstd, float* y, void* stream) {
  WMRL_REQUIRE(M >= 0 && W > 0, "ln_act_fwd: bad shape");
  WMRL_REQUIRE(act >= 0 && act <= 2, "ln_act_fwd: act must be 0 (ELU), 1 (SiLU) or 2 (ReLU)");
  if (M == 0) return 0;
  WMRL_REQUIRE(x && gamma && beta && xhat && rstd && y, "ln_act_fwd: NULL tensor");
  return wmrl::ln_act_fwd((cudaStream_t)stream, act, x, gamma, beta, xhat, rstd, y, M, W);
}
int wmrl_ln_act_bwd(int32_t M, int32_t W, const float* dy, const float* xhat, const float* rstd,
                    const float* gamma, const float* beta, int32_t act, float* dx, float* dgamma, float* dbeta,
                    void* stream) {
  WMRL_REQUIRE(M >= 0 && W > 0, "ln_act_bwd: bad shape");
  WMRL_REQUIRE(act >= 0 && act <= 2, "ln_act_bwd: act must be 0 (ELU), 1 (SiLU) or 2 (ReLU)");
  if (M == 0) return 0;
  WMRL_REQUIRE(dy && xhat && rstd && gamma && beta && dx, "ln_act_bwd: NULL tensor");
  return wmrl::ln_act_bwd((cudaStream_t)stream, act, dy, xhat, rstd, gamma, beta, dx, dgamma, dbeta, M, W);
}
int wmrl_linear_fwd(int32_t M, int32_t N, int32_t K, const float* X, int64_t ldx, const float* W,
                    const float* bias, float* Y, int64_t ldy, int32_t act, void* stream) {
  WMRL_REQUIRE(M >= 0 && N > 0 && K > 0, "linear_fwd: bad shape");
  WMRL_REQUIRE(act == 0 || act == 1, "linear_fwd: act must be 0 (none) or 1 (ELU)");
  if (M == 0) return 0;
  WMRL_REQUIRE(X && W && Y, "linear_fwd: NULL tensor");
  return wmrl::linear((cudaStream_t)stream, X, ldx, K, nullptr, 0, 0, W, bias, Y, ldy, M, N, act);
}
int wmrl_linear_dgrad(int32_t M, int32_t N, int32_t K, const float* dY, int64_t lddy, const float* W,
                      float* dX, int64_t lddx, int32_t accumulate, void* stream) {
  WMRL_REQUIRE(M >= 0 && N > 0 && K > 0, "linear_dgrad: bad shape");
  if (M == 0) return 0;
  WMRL_REQUIRE(dY && W && dX, "linear_dgrad: NULL tensor");
  return wmrl::dgrad((cudaStream_t)stream, dY, lddy, N, W, K, 0, K, dX, lddx, M, accumulate);
}
int wmrl_linear_wgrad(int32_t M, int32_t N, int32_t K, const float* dY, int64_t lddy, const float* X,
                      int64_t ldx, float* dW, void* stream) {
  WMRL_REQUIRE(M >= 0 && N > 0 && K > 0, "linear_wgrad: bad shape");
  if (M == 0) return 0;
  WMRL_REQUIRE(dY && X && dW, "linear_wgrad: NULL tensor");
  return wmrl::wgrad((cudaStream_t)stream, dY, lddy, N, X, ldx, K, dW, K, 0, M);
}
}
