// fp32-accurate GEMM on the 5th-gen tensor cores (tcgen05, kind::tf32, 3xTF32 split).
//
// Every Linear / dgrad / wgrad of the fp32 chain (f32_path.cu: actor `actor.py:47-52`, prior
// `rssm.py:53-68`, posterior `rssm.py:70-82` and their backward) goes through this kernel when the
// operand has more than 64 rows.  An fp32 value x is split into hi = x with the low 13 mantissa bits
// cleared (exactly what the tensor core reads of an fp32 word under kind::tf32) and lo = x - hi
// (exact in fp32); a*b ~ hi*hi + hi*lo + lo*hi drops only the lo*lo term (~2^-22 relative), so the
// result matches an fp32 SIMT GEMM to fp32 rounding while running on the tensor pipe.
//
// Data flow per CTA (one 128 x BN output tile, BN <= 256, optional K slice):
//   global fp32 (any strides, two A segments)  --LDG (prefetched one k-block ahead)-->  registers
//   --split hi/lo, STS.128-->  shared, canonical K-major no-swizzle core-matrix panels
//   --tcgen05.mma x 6 per 16-wide k-block (issued by one thread, async)-->  TMEM fp32 accumulators
//   --tcgen05.ld-->  registers --> shared transpose tile --> coalesced global store
//   (bias / ELU / += / atomic add for split-K fused in the store).
// Two stages of operand panels; tcgen05.commit on a per-stage mbarrier releases a stage when the
// MMAs that read it have retired.  Two CTAs fit per SM (97 KB smem, 256 TMEM columns each), so one
// CTA's epilogue and staging overlap the other's MMAs.
#include <stdint.h>
#include <stdlib.h>

#include "common.cuh"
#include "gemm.cuh"
#include "tc_ptx.cuh"

namespace wmrl {

namespace {

constexpr int TG_BM = 128, TG_BK = 16, TG_STAGES = 2, TG_THREADS = 256;
constexpr int TG_QUADS = TG_BK / 4;                 // 16-byte k chunks (4 tf32) per k-block
// panel(row, quad) = quad * LBO + row * 16 bytes; LBO = rows*16 + 32 keeps 8-lane STS.128 phases
// (2 rows x 4 quads) on distinct bank groups.
constexpr int TG_LBO_A = TG_BM * 16 + 32;           // 2080
constexpr int TG_LBO_B = 256 * 16 + 32;             // 4128 (sized for the widest tile)
constexpr int TG_A_BYTES = TG_QUADS * TG_LBO_A;     // one of hi / lo
constexpr int TG_B_BYTES = TG_QUADS * TG_LBO_B;
constexpr int TG_STAGE_BYTES = 2 * TG_A_BYTES + 2 * TG_B_BYTES;   // 49,664
constexpr int TG_SMEM_BYTES = TG_STAGES * TG_STAGE_BYTES;
constexpr int TG_TILE_LD = 65;                      // epilogue transpose tile [128][64 + 1] floats
static_assert(TG_BM * TG_TILE_LD * 4 <= TG_SMEM_BYTES, "epilogue tile must fit in the operand stages");

struct Operand {   // element (row, k) of a two-segment operand
  const float* p1; long long si1, sk1; int K1;
  const float* p2; long long si2, sk2;
};

__device__ __forceinline__ float4 load_quad(const Operand& o, bool row_ok, long long row, int k,
                                            int kend) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!row_ok || k >= kend) return v;
  const float* p;
  long long sk;
  int lim;
  if (k < o.K1) {
    p = o.p1 + row * o.si1 + (long long)k * o.sk1; sk = o.sk1; lim = min(o.K1, kend) - k;
  } else {
    p = o.p2 + row * o.si2 + (long long)(k - o.K1) * o.sk2; sk = o.sk2; lim = kend - k;
  }
  if (lim >= 4) {
    if (sk == 1 && (reinterpret_cast<uintptr_t>(p) & 15) == 0)
      return *reinterpret_cast<const float4*>(p);
    v.x = p[0]; v.y = p[sk]; v.z = p[2 * sk]; v.w = p[3 * sk];
    return v;
  }
  // the quad straddles the segment boundary or the end of the K slice
  float e[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int kk = k + j;
    e[j] = 0.f;
    if (kk < kend)
      e[j] = kk < o.K1 ? o.p1[row * o.si1 + (long long)kk * o.sk1]
                       : o.p2[row * o.si2 + (long long)(kk - o.K1) * o.sk2];
  }
  return make_float4(e[0], e[1], e[2], e[3]);
}

__device__ __forceinline__ void store_split(uint32_t hi_addr, uint32_t lo_addr, float4 v) {
  const uint32_t m = 0xffffe000u;
  const float hx = __uint_as_float(__float_as_uint(v.x) & m), hy = __uint_as_float(__float_as_uint(v.y) & m),
              hz = __uint_as_float(__float_as_uint(v.z) & m), hw = __uint_as_float(__float_as_uint(v.w) & m);
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(hi_addr), "f"(hx), "f"(hy), "f"(hz), "f"(hw)
               : "memory");
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(lo_addr), "f"(v.x - hx), "f"(v.y - hy),
               "f"(v.z - hz), "f"(v.w - hw)
               : "memory");
}

// instruction descriptor, kind::tf32: D = f32, A = B = tf32, both K-major.
__device__ __forceinline__ uint32_t make_idesc_tf32(uint32_t M, uint32_t N) {
  uint32_t d = 0;
  d |= 1u << 4;            // [4,6)   D format: f32
  d |= 2u << 7;            // [7,10)  A format: tf32
  d |= 2u << 10;           // [10,13) B format: tf32
  d |= (N >> 3) << 17;     // [17,23) N >> 3
  d |= (M >> 4) << 24;     // [24,29) M >> 4
  return d;
}
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ float elu_ref(float x) { return x > 0.f ? x : expm1f(x); }

// item -> (row, quad) of a [rows x 4 quads] panel.  k-contiguous operands put the 4 quads of a row on
// adjacent lanes (64 B global segments, conflict-free STS with the padded LBO); row-contiguous
// operands put consecutive rows on adjacent lanes (coalesced scalar loads, contiguous STS.128).
__device__ __forceinline__ void item_rc(int it, int rows, bool kfast, int& row, int& quad) {
  if (kfast) { quad = it & 3; row = it >> 2; } else { row = it % rows; quad = it / rows; }
}

// predicated, branch-free loads into pre-zeroed registers
__device__ __forceinline__ void ldg128_if(float4& v, const float* p, bool pred) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t"
      "@q ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
      : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)
      : "l"(p), "r"((int)pred));
}
__device__ __forceinline__ void ldg32_if(float& v, const float* p, bool pred) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t"
      "@q ld.global.nc.f32 %0, [%1];\n\t}"
      : "+f"(v)
      : "l"(p), "r"((int)pred));
}
__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// 256 staging / epilogue threads + one MMA-issuing warp.
constexpr int TG_BLOCK = TG_THREADS + 32;

// <P, S>: P k-blocks of operand loads in flight per thread (register sets), S shared-memory stages.
// <1, 2> (two CTAs per SM) is the shipped instantiation.
template <int P, int S>
__global__ void __launch_bounds__(TG_BLOCK, (P == 1 ? 2 : 1)) tf32x3_gemm_kernel(GemmArgs g, int BN, int tmem_cols) {
  extern __shared__ __align__(128) uint8_t tg_smem[];
  __shared__ __align__(8) uint64_t bars[2 * S + 1];   // full[S] | empty[S] | accumulator ready
  __shared__ uint32_t tmem_slot;
  using namespace ptx;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * TG_BM;
  const int K = g.K1 + g.K2;
  const int kbeg = blockIdx.z * g.kper, kend = min(K, kbeg + g.kper);
  const int nkb = kend > kbeg ? (kend - kbeg + TG_BK - 1) / TG_BK : 0;
  const uint32_t smem0 = smem_u32(tg_smem);
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * S, bar_acc = bar_full + 16 * S;

  if (tid == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(bar_full + 8 * i, TG_THREADS);
      mbar_init(bar_empty + 8 * i, 1);
    }
    mbar_init(bar_acc, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_slot), (uint32_t)tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == TG_THREADS / 32) {
    // ------------------------------------------------------------------ MMA issuer (one lane)
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(TG_BM, (uint32_t)BN);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % S;
        const uint32_t st = smem0 + s * TG_STAGE_BYTES;
        mbar_wait(bar_full + 8 * s, (kb / S) & 1);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < TG_BK / 8; ++ks) {
          const uint32_t a_hi = st + ks * 2 * TG_LBO_A, a_lo = a_hi + TG_A_BYTES;
          const uint32_t b_hi = st + 2 * TG_A_BYTES + ks * 2 * TG_LBO_B, b_lo = b_hi + TG_B_BYTES;
          const uint64_t dah = make_smem_desc(a_hi, TG_LBO_A, 128), dal = make_smem_desc(a_lo, TG_LBO_A, 128);
          const uint64_t dbh = make_smem_desc(b_hi, TG_LBO_B, 128), dbl = make_smem_desc(b_lo, TG_LBO_B, 128);
          mma_tf32_ss(tmem, dal, dbh, idesc, (kb | ks) != 0);   // small terms first
          mma_tf32_ss(tmem, dah, dbl, idesc, 1);
          mma_tf32_ss(tmem, dah, dbh, idesc, 1);
        }
        tc_commit(bar_empty + 8 * s);
        if (kb == nkb - 1) tc_commit(bar_acc);
      }
    }
  } else {
    // ------------------------------------------------------------------ staging + epilogue threads
    const Operand oa{g.A1, g.a1_si, g.a1_sk, g.K1, g.A2, g.a2_si, g.a2_sk};
    const Operand ob{g.Bm, g.b_sj, g.b_sk, K, g.Bm, g.b_sj, g.b_sk};   // B(k, j): row = j
    const bool a_kfast = (g.a1_sk == 1), b_kfast = (g.b_sk == 1);
    const int b_items = BN * TG_QUADS;
    // How a whole 16-wide k-block of an operand can be loaded - uniform over the CTA, so the common cases
    // are branch-free predicated loads off per-item row pointers (the generic path costs ~50 instructions
    // per quad, which at 8 warps per SM was the bottleneck):
    //   vec: k-contiguous, 16-byte aligned rows            -> one LDG.128 per quad
    //   str: row-contiguous (dgrad's W, wgrad's dY^T / X)   -> 4 coalesced LDG.32 per quad
    const bool a_vec1 = g.a1_sk == 1 && aligned16(g.A1) && (g.a1_si & 3) == 0;
    const bool a_vec2 = g.K2 > 0 && g.a2_sk == 1 && aligned16(g.A2) && (g.a2_si & 3) == 0 && (g.K1 & 3) == 0;
    const bool a_str = g.K2 == 0 && g.a1_si == 1 && g.a1_sk != 1;
    const bool b_vec = g.b_sk == 1 && aligned16(g.Bm) && (g.b_sj & 3) == 0;
    const bool b_str = g.b_sj == 1 && g.b_sk != 1;

    int a_row[2], a_kq[2], b_row[4], b_kq[4];
    bool a_ok[2], b_ok[4];
    uint32_t a_soff[2], b_soff[4];
    const float* a_p1[2];
    const float* a_p2[2];
    const float* b_p[4];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      int q;
      item_rc(tid + TG_THREADS * j, TG_BM, a_kfast, a_row[j], q);
      a_kq[j] = q * 4;
      a_ok[j] = (m0 + a_row[j]) < g.M;
      a_soff[j] = q * TG_LBO_A + a_row[j] * 16;
      const long long row = m0 + a_row[j];
      a_p1[j] = a_str ? g.A1 + row + (long long)a_kq[j] * g.a1_sk : g.A1 + row * g.a1_si + a_kq[j];
      a_p2[j] = g.K2 > 0 ? g.A2 + row * g.a2_si + a_kq[j] - g.K1 : a_p1[j];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int it = tid + TG_THREADS * j;
      int q;
      item_rc(it < b_items ? it : 0, BN, b_kfast, b_row[j], q);
      b_kq[j] = q * 4;
      b_ok[j] = it < b_items && (n0 + b_row[j]) < g.N;
      b_soff[j] = 2 * TG_A_BYTES + q * TG_LBO_B + b_row[j] * 16;
      const long long row = n0 + b_row[j];
      b_p[j] = b_str ? g.Bm + row + (long long)b_kq[j] * g.b_sk : g.Bm + row * g.b_sj + b_kq[j];
      if (it >= b_items) b_row[j] = -1;   // nothing to stage
    }

    float4 rav[P][2], rbv[P][4];
    auto fetch = [&](int kb, float4 (&ra)[2], float4 (&rb)[4]) {
      const int k0 = kbeg + kb * TG_BK;
      const bool whole = k0 + TG_BK <= kend;
      if (whole && a_vec1 && k0 + TG_BK <= g.K1) {
#pragma unroll
        for (int j = 0; j < 2; ++j) { ra[j] = make_float4(0.f, 0.f, 0.f, 0.f); ldg128_if(ra[j], a_p1[j] + k0, a_ok[j]); }
      } else if (whole && a_vec2 && k0 >= g.K1) {
#pragma unroll
        for (int j = 0; j < 2; ++j) { ra[j] = make_float4(0.f, 0.f, 0.f, 0.f); ldg128_if(ra[j], a_p2[j] + k0, a_ok[j]); }
      } else if (whole && a_str) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const float* p = a_p1[j] + (long long)k0 * g.a1_sk;
          ra[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          ldg32_if(ra[j].x, p, a_ok[j]); ldg32_if(ra[j].y, p + g.a1_sk, a_ok[j]);
          ldg32_if(ra[j].z, p + 2 * g.a1_sk, a_ok[j]); ldg32_if(ra[j].w, p + 3 * g.a1_sk, a_ok[j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 2; ++j) ra[j] = load_quad(oa, a_ok[j], m0 + a_row[j], k0 + a_kq[j], kend);
      }
      if (whole && b_vec) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { rb[j] = make_float4(0.f, 0.f, 0.f, 0.f); ldg128_if(rb[j], b_p[j] + k0, b_ok[j]); }
      } else if (whole && b_str) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float* p = b_p[j] + (long long)k0 * g.b_sk;
          rb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          ldg32_if(rb[j].x, p, b_ok[j]); ldg32_if(rb[j].y, p + g.b_sk, b_ok[j]);
          ldg32_if(rb[j].z, p + 2 * g.b_sk, b_ok[j]); ldg32_if(rb[j].w, p + 3 * g.b_sk, b_ok[j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) rb[j] = load_quad(ob, b_ok[j], n0 + b_row[j], k0 + b_kq[j], kend);
      }
    };
#pragma unroll
    for (int u = 0; u < P; ++u)
      if (u < nkb) fetch(u, rav[u], rbv[u]);

    for (int kb0 = 0; kb0 < nkb; kb0 += P) {
#pragma unroll
      for (int u = 0; u < P; ++u) {
        const int kb = kb0 + u;
        if (kb < nkb) {
          const int s = kb % S;
          const uint32_t st = smem0 + s * TG_STAGE_BYTES;
          if (kb >= S) mbar_wait(bar_empty + 8 * s, ((kb / S) - 1) & 1);
#pragma unroll
          for (int j = 0; j < 2; ++j) store_split(st + a_soff[j], st + TG_A_BYTES + a_soff[j], rav[u][j]);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (b_row[j] >= 0) store_split(st + b_soff[j], st + TG_B_BYTES + b_soff[j], rbv[u][j]);
          if (kb + P < nkb) fetch(kb + P, rav[u], rbv[u]);
          fence_proxy_async_smem();
          mbar_arrive(bar_full + 8 * s);
        }
      }
    }

    if (nkb > 0) {
      mbar_wait(bar_acc, 0);
      tc_fence_after();
    }
    // Epilogue: 64 columns per round.  Warp w owns TMEM lanes 32*(w%4).., columns (w/4)*32.. of the round.
    float* tile = reinterpret_cast<float*>(tg_smem);
    const bool atomic = gridDim.z > 1;
    const int ncols_tile = min(BN, g.N - n0);
    for (int c0 = 0; c0 < BN; c0 += 64) {
      const int cw = c0 + (warp >> 2) * 32;
      if (cw < BN) {
        float v[32];
        if (nkb > 0) {
          tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)cw, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
        float* trow = tile + ((warp & 3) * 32 + lane) * TG_TILE_LD + (warp >> 2) * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j) trow[j] = v[j];
      }
      asm volatile("bar.sync 1, %0;" ::"r"(TG_THREADS) : "memory");
      for (int idx = tid; idx < TG_BM * 64; idx += TG_THREADS) {
        const int r = idx >> 6, c = idx & 63;
        const int gm = m0 + r, col = c0 + c;
        if (gm < g.M && col < ncols_tile) {
          const int gn = n0 + col;
          float v = tile[r * TG_TILE_LD + c];
          if (g.bias && blockIdx.z == 0) v += g.bias[gn];
          float* cp = g.C + (long long)gm * g.ldc + gn;
          if (atomic) {
            atomicAdd(cp, v);
          } else {
            if (g.act == 1) v = elu_ref(v);
            *cp = g.accumulate ? (*cp + v) : v;
          }
        }
      }
      asm volatile("bar.sync 1, %0;" ::"r"(TG_THREADS) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, (uint32_t)tmem_cols);
}

}  // namespace

int tf32x3_gemm_launch(cudaStream_t st, const GemmArgs& g_in, int splitk) {
  GemmArgs g = g_in;
  static bool attr_set = false;
  if (!attr_set) {
    WMRL_CUDA_OK(cudaFuncSetAttribute(tf32x3_gemm_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      2 * TG_STAGE_BYTES));
    attr_set = true;
  }
  const int K = g.K1 + g.K2;
  if (splitk < 1) splitk = 1;
  int kper = ((K + splitk - 1) / splitk + TG_BK - 1) / TG_BK * TG_BK;
  if (kper < TG_BK) kper = TG_BK;
  splitk = (K + kper - 1) / kper;
  if (splitk < 1) splitk = 1;
  g.kper = kper;
  // Column tiles: as few as possible, narrowed (>= 32 wide) until there are two CTAs per SM.  (A one-CTA-per-SM
  // instantiation with three k-blocks of loads in flight, <3, 4>, was measured 1.5x slower at 2 500 rows: the
  // staging is throughput-bound, not latency-bound, and a second resident CTA overlaps it better.)
  const int mt = (g.M + TG_BM - 1) / TG_BM;
  int nt = (g.N + 255) / 256;
  while ((long long)mt * nt * splitk < 296 && (g.N + nt) / (nt + 1) >= 32) ++nt;
  int BN = ((g.N + nt - 1) / nt + 15) / 16 * 16;
  if (BN < 16) BN = 16;
  nt = (g.N + BN - 1) / BN;
  int tmem_cols = 32;
  while (tmem_cols < BN) tmem_cols *= 2;
  dim3 grid(nt, mt, splitk);
  tf32x3_gemm_kernel<1, 2><<<grid, TG_BLOCK, 2 * TG_STAGE_BYTES, st>>>(g, BN, tmem_cols);
  WMRL_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace wmrl
