// Inline-PTX wrappers for the sm_100a features the fused RSSM kernel uses:
// mbarrier, bulk async copy (TMA engine, 1-D), tcgen05 MMA / TMEM / commit.
// Descriptor bit layouts follow the PTX ISA "tcgen05 matrix / instruction
// descriptor" tables (cross-checked against cute/arch/mma_sm100_desc.hpp).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace wmrl {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Spin with a watchdog: a protocol bug must not hang the GPU box - after ~2^28
// polls the kernel traps (reported to the host as a launch failure).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) {
      printf("wmrl: mbarrier watchdog (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x,
             threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// Long waits (epilogue warps): WMRL_WAIT_HINT > 0 passes a suspend-time hint (ns) to try_wait so the hardware parks the
// warp until the phase completes instead of polling; WMRL_WAIT_SLEEP (ns) is the back-off between polls (0: none).
#ifndef WMRL_WAIT_HINT
#define WMRL_WAIT_HINT 0
#endif
#ifndef WMRL_WAIT_SLEEP
#define WMRL_WAIT_SLEEP 64
#endif
__device__ __forceinline__ bool mbar_try_wait_long(uint32_t bar, uint32_t parity) {
#if WMRL_WAIT_HINT > 0
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"((uint32_t)WMRL_WAIT_HINT)
      : "memory");
  return ok != 0;
#else
  return mbar_try_wait(bar, parity);
#endif
}
__device__ __forceinline__ void mbar_long_backoff() {
#if WMRL_WAIT_SLEEP > 0
  __nanosleep(WMRL_WAIT_SLEEP);
#endif
}
// Same, for waiters that expect to wait long (epilogue warps): sleep between polls so the
// polling does not steal issue slots from the producer / MMA warps.
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_long(bar, parity)) {
    mbar_long_backoff();
    if (++spins > (1u << 24)) {
      printf("wmrl: mbarrier watchdog (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x,
             threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// Same without the printf (a real call: every register live across the wait would be spilled around it) - for
// waiters that hold prefetched operands in registers.  A protocol bug still traps.
__device__ __forceinline__ void mbar_wait_backoff_quiet(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_long(bar, parity)) {
    mbar_long_backoff();
    if (++spins > (1u << 24)) __trap();
  }
}

// ---------------------------------------------------------------- cluster (CTA pair)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// arrive on an mbarrier that lives in another CTA of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// same with cluster-scope release: generic-proxy stores made (here or, ordered by __syncwarp, by the other lanes) before
// the arrive are visible to a cluster-scope acquire wait on that barrier
__device__ __forceinline__ void mbar_arrive_remote_release(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void st_cluster_f32x2(uint32_t cluster_addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1,%2};" ::"r"(cluster_addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
// wait on a local mbarrier whose arrivals may come from the peer CTA (cluster-scope acquire)
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (++spins > (1u << 26)) {
      printf("wmrl: mbarrier watchdog/cluster (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x,
             threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// same without the printf (see mbar_wait_backoff_quiet): a protocol bug still traps
__device__ __forceinline__ void mbar_wait_cluster_quiet(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity))
    if (++spins > (1u << 26)) __trap();
}

// generic-proxy smem writes -> visible to the async proxy (tensor core / TMA)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ------------------------------------------------------- bulk copy (TMA 1-D)
// global -> shared, completion signalled on an mbarrier as transaction bytes.
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes,
                                         uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          dst_smem),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

// ------------------------------------------------------------------ tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread complete
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}

// ---- cta_group::2 variants (CTA pair: M=256 split over two SMs, each SM holds half of B)
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// commit -> arrive on the mbarrier at this smem offset in every CTA of `mask`
__device__ __forceinline__ void tc_commit2(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          bar),
      "h"(mask)
      : "memory");
}
__device__ __forceinline__ void mma_bf16_ss2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// K-major, no-swizzle shared-memory matrix descriptor.  The operand panel is
// stored as [k8 chunk][row][8 x bf16]: a core matrix (8 rows x 16 B) is 128
// contiguous bytes; SBO = 128 B between 8-row groups, LBO = rows*16 B between
// the two 8-wide K chunks of one K=16 MMA.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes,
                                                   uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);            // [0,14)  start address >> 4
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;       // [16,30) leading byte offset >> 4
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;       // [32,46) stride byte offset >> 4
  d |= (uint64_t)1 << 46;                                 // [46,48) descriptor version 1 (sm_100)
  // base offset 0, lbo mode 0, layout type [61,64) = 0: SWIZZLE_NONE
  return d;
}

// instruction descriptor, kind::f16: D=f32, A=B=bf16, both K-major, M=128.
__device__ __forceinline__ uint32_t make_idesc_bf16(uint32_t M, uint32_t N) {
  uint32_t d = 0;
  d |= 1u << 4;            // [4,6)   D format: f32
  d |= 1u << 7;            // [7,10)  A format: bf16
  d |= 1u << 10;           // [10,13) B format: bf16
  // [13],[14] negate = 0; [15] A major = K (0); [16] B major = K (0)
  d |= (N >> 3) << 17;     // [17,23) N >> 3
  d |= (M >> 4) << 24;     // [24,29) M >> 4
  return d;
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
__device__ __forceinline__ void mma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Lean issue forms: descriptors are built once per stage; per K-step only the low word (the
// 14-bit address field) advances.  The issuing thread is a single lane - every extra integer
// instruction between two MMAs costs ~4 cycles of issue latency.
__device__ __forceinline__ void mma_bf16_lohi(uint32_t d_tmem, uint32_t alo, uint32_t ahi, uint32_t blo,
                                              uint32_t bhi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
      "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_bf16_lohi2(uint32_t d_tmem, uint32_t alo, uint32_t ahi, uint32_t blo,
                                               uint32_t bhi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
      "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// TMEM -> registers: each thread of the warp reads its own lane (row), N
// consecutive 32-bit columns.  The warp may only touch lanes 32*(warp%4)..+31.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ------------------------------------------------------------------- math
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// ELU without predicates (a select per element makes the compiler pack 32 predicates into
// registers): elu(z) = max(z,0) + (exp(min(z,0)) - 1).
// elu_scaled takes zs = z*log2(e):  ln2*max(zs,0) + (2^min(zs,0) - 1)
__device__ __forceinline__ float elu_scaled(float zs) {
  const float e = ex2_fast(fminf(zs, 0.f)) - 1.f;
  return fmaf(fmaxf(zs, 0.f), 0.6931471805599453f, e);
}
__device__ __forceinline__ float elu_fast(float z) { return elu_scaled(z * 1.4426950408889634f); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
// store 8 consecutive K-values of one row as a 16-byte chunk of an A panel
__device__ __forceinline__ void st_chunk(uint32_t smem_addr, const float* v) {
  const uint32_t a = pack_bf16x2(v[0], v[1]), b = pack_bf16x2(v[2], v[3]);
  const uint32_t c = pack_bf16x2(v[4], v[5]), d = pack_bf16x2(v[6], v[7]);
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(smem_addr), "r"(a), "r"(b), "r"(c),
               "r"(d)
               : "memory");
}
// same, and the identical 16 bytes to a global-memory stash panel (training: activations recorded for the BPTT)
__device__ __forceinline__ void st_chunk_sg(uint32_t smem_addr, void* gptr, const float* v) {
  const uint32_t a = pack_bf16x2(v[0], v[1]), b = pack_bf16x2(v[2], v[3]);
  const uint32_t c = pack_bf16x2(v[4], v[5]), d = pack_bf16x2(v[6], v[7]);
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(smem_addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
  *reinterpret_cast<uint4*>(gptr) = make_uint4(a, b, c, d);
}
__device__ __forceinline__ void st_chunk_g(void* gptr, const float* v) {
  *reinterpret_cast<uint4*>(gptr) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                               pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}
// streaming variants (st.global.cs: evict-first) for write-once data that is read back only by a later kernel, so that
// it does not push the operands the running kernel has prefetched out of L2
__device__ __forceinline__ void st_g_cs(void* gptr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.cs.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(gptr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_chunk_g_cs(void* gptr, const float* v) {
  st_g_cs(gptr, pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}
__device__ __forceinline__ void st_chunk_sg_cs(uint32_t smem_addr, void* gptr, const float* v) {
  const uint32_t a = pack_bf16x2(v[0], v[1]), b = pack_bf16x2(v[2], v[3]);
  const uint32_t c = pack_bf16x2(v[4], v[5]), d = pack_bf16x2(v[6], v[7]);
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(smem_addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
  st_g_cs(gptr, a, b, c, d);
}
// 8 bf16 of a stash panel chunk -> fp32
__device__ __forceinline__ void ld_chunk_g(const void* gptr, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(gptr);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
}

// Packed fp32x2 arithmetic (sm_100: FADD2 / FFMA2, two fp32 per issue slot) for the issue-bound epilogues.
__device__ __forceinline__ void add2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{\n\t.reg .b64 a, b, d;\n\tmov.b64 a, {%2,%3};\n\tmov.b64 b, {%4,%5};\n\tadd.rn.f32x2 d, a, b;\n\tmov.b64 {%0,%1}, d;\n\t}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void fma2(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0,
                                     float c1) {
  asm("{\n\t.reg .b64 a, b, c, d;\n\tmov.b64 a, {%2,%3};\n\tmov.b64 b, {%4,%5};\n\tmov.b64 c, {%6,%7};\n\t"
      "fma.rn.f32x2 d, a, b, c;\n\tmov.b64 {%0,%1}, d;\n\t}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}

// two ELUs at once: the -1 and the final multiply-add are packed (4 issue slots per element instead of 5)
__device__ __forceinline__ void elu_scaled2(float& y0, float& y1, float zs0, float zs1) {
  float e0 = ex2_fast(fminf(zs0, 0.f)), e1 = ex2_fast(fminf(zs1, 0.f));
  add2(e0, e1, e0, e1, -1.f, -1.f);
  fma2(y0, y1, fmaxf(zs0, 0.f), fmaxf(zs1, 0.f), 0.6931471805599453f, 0.6931471805599453f, e0, e1);
}

}  // namespace ptx
}  // namespace wmrl
