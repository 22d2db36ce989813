// bf16 tensor-core BPTT of RSSMBase.imagine_rollout (autograd of rssm.py:123-140 through actor.py:47-52 and
// rssm.py:53-68 with the world model frozen, world_model.py:172, and the actor input detached, rssm.py:135):
//
//   tc_bptt_kernel     ONE persistent kernel per call, the rollout kernel's structure run backwards in time: CTA
//                      pairs (tcgen05.mma.cta_group::2, M=128, 64 rows per CTA, accumulators folded over the TMEM
//                      lane halves), gradient panels in shared memory as bf16 MMA A operands, TRANSPOSED weight
//                      stages streamed L2 -> smem by the TMA engine through a ring, fused epilogues:
//                        Gaussian-sample backward -> prior head dgrad (x ELU') -> GRU gate backward -> W_ih^T / W_hh^T
//                        dgrad -> embed dgrad -> tanh-squash backward -> actor dgrad with LayerNorm+ELU backward.
//                      The carried dL/dh lives in an fp32 scratch row owned by the thread that owns the columns,
//                      dL/ds in shared memory.  Forward activations come from the bf16 stash the rollout kernel
//                      recorded (StashLayout); dL/d(pre-LayerNorm), dL/d(LayerNorm out) and dL/d(mu) go to the
//                      gradient stash in the same panel layout.
//   tc_wgrad_kernel    actor weight gradients as big-K GEMMs over the two stashes after the scan:
//                      dW[out][in] += sum_blocks G_blk^T Y_blk, both operands read as MN-major tcgen05 operands
//                      straight from the panel layout (no transpose pass), split-K over the SMs, fp32 atomics.
//   colsum_kernel      bias / LayerNorm-affine gradients: column sums of gradient panels (HBM-bound).
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "tc_plan.cuh"
#include "tc_ptx.cuh"

namespace wmrl {
using namespace ptx;

extern std::mutex g_tc_mutex;
unsigned long long tc_pack_generation(const void* packed);

constexpr int kMaxStagesB = 144, kMaxJobsB = 16, kMaxEparamsB = 8192;
constexpr int kBarBytesB = 512;
constexpr uint32_t kPSB = 1024;   // bytes per 8-column chunk of a 64-row panel
constexpr int kRowsB = 64;
constexpr int kThreadsB = kEpiThreads + 96;   // 16 epilogue warps + producer + MMA + ready-counter warps
constexpr int kReadyWarpB = kMmaWarp + 1;
__constant__ uint4 cb_stages[kMaxStagesB];
__constant__ JobDesc cb_jobs[kMaxJobsB];
__constant__ __align__(16) float cb_eparams[kMaxEparamsB];

enum : uint8_t { JB_ELUB = 1, JB_GRUB, JB_CARRY, JB_EMB, JB_LNB };

struct BwdParams {
  const uint8_t* stage_data;
  const uint8_t* stash_f;
  uint8_t* stash_g;
  float* gh_scratch;        // fp32 carried dL/dh (and the GRU's direct z-path term in between), one PANEL per (tile, CTA):
                            // [Dp/4][64 rows][4] - a warp's float4 access to 4 columns of its 32 rows is 512 contiguous
                            // bytes (the row-major form cost 32 LSU wavefronts per instruction)
  const float *deter_seq, *std_seq, *actions, *noise;
  const float *g_deter, *g_stoch, *g_mean, *g_std;   // upstream gradients of the sequence outputs (any may be null)
  float *g_deter0, *g_stoch0;
  int nstages, njobs;
  int B, H, D, S, A, Dp, Sp, Ha, Hdp, La;
  float inv_action_scale;
  uint32_t off_go, off_gmu, off_gx, off_big, off_carry, off_red, ring_off, bar_off, slot_bytes, bound_off;
  int nslots, ntiles;
  StashLayout sl;
  long long* trace;   // WMRL_FLAG_TRACE: SM clocks of CTA 0's first tile: [backward step][job][4] =
                      // {first MMA issued, accumulator committed, epilogue woke, epilogue done}
};

// --------------------------------------------------------------------------
// plan
// --------------------------------------------------------------------------
Plan tc_make_bwd_plan(const Dims& d, const wmrl_rssm_params* w, const wmrl_actor_params* aw) {
  Plan pl;
  pl.cta = 2; pl.fold = true; pl.pp = false; pl.ps = kPSB;
  if (d.variant != WMRL_CONTINUOUS) return pl;
  if (d.Ha % 64 != 0 || d.Ha > 512 || d.A > 8 || d.S > 64 || d.La < 1) return pl;
  pl.Dp = ceil16(d.D); pl.Sp = ceil16(d.S); pl.Ap = 16; pl.Ha = d.Ha; pl.Hdp = ceil16(d.Hd);
  if (pl.Dp > 256 || pl.Hdp > 256) return pl;   // one accumulator group per dynamics GEMM, two static TMEM quarters
  pl.sl = make_stash_layout(pl.Dp, pl.Sp, d.Ha, pl.Hdp, d.La);
  const uint32_t PS = kPSB;
  const uint32_t off_go = 0;
  const uint32_t off_gmu = off_go + 2 * pl.Sp / 8 * PS;
  const uint32_t off_gx = off_gmu + 16 / 8 * PS;
  const uint32_t off_big = off_gx + std::max(pl.Dp, pl.Hdp) / 8 * PS;
  const uint32_t off_carry = off_big + std::max(4 * pl.Dp, d.Ha) / 8 * PS;
  const uint32_t off_red = off_carry + kRowsB * pl.Sp * 4;
  pl.panels_bytes = off_red + 2 * kQ * kRowsB * 8;
  pl.ring_off = (pl.panels_bytes + 1023) / 1024 * 1024;
  pl.boff[0] = off_go; pl.boff[1] = off_gmu; pl.boff[2] = off_gx; pl.boff[3] = off_big; pl.boff[4] = off_carry;
  pl.boff[5] = off_red;
  const int avail = kMaxSmem - (int)pl.ring_off - kBarBytesB - kMaxStagesB * 16;
  pl.stage_bytes = (avail / 32768 >= 3) ? 65536 : 32768;
  pl.nslots = avail / (pl.stage_bytes / 2);
  if (pl.nslots < 2) return pl;
  if (pl.nslots > 16) pl.nslots = 16;
  pl.bar_off = pl.ring_off + pl.nslots * (pl.stage_bytes / 2);
  pl.smem_bytes = pl.bar_off + kBarBytesB + kMaxStagesB * 16;
  const float* nul = nullptr;
  auto job = [&](uint8_t type, int n, int tmem_col, uint32_t dst, uint32_t st0, uint32_t st1, uint32_t layer,
                 uint32_t p_off, uint32_t aux) {
    JobDesc jb{};
    jb.type = type; jb.n = (uint16_t)n; jb.tmem_col = (uint16_t)tmem_col; jb.dst_off = dst; jb.st0 = st0; jb.st1 = st1;
    jb.layer = layer; jb.p_off = p_off; jb.aux = aux;
    pl.jobs.push_back(jb);
  };
  const int kDynA = 256, kDynB = 384;   // TMEM columns of the two dynamics accumulators (actor accumulators at 0)
  // ---- J2: dL/d(prior hidden) = [g_mean | g_raw] W2  (x ELU'(hid))
  {
    job(JB_ELUB, pl.Hdp, kDynA, off_gx, pl.sl.hid, kNoStash, 0, 0, 0);
    Seg s{off_go, 2 * pl.Sp, w ? w->prior2_w : nul, d.Hd, 0, 0};
    s.trans = 1; s.add_k(0, d.S, pl.Sp); s.add_k(d.S, d.S, pl.Sp);
    add_acc_group(pl, RowSpec(0, d.Hd, pl.Hdp), {s}, kDynA, false, true, 1, true);
  }
  // ---- J3: dL/dh' += g_hid W1, then the GRU gate backward
  {
    job(JB_GRUB, pl.Dp, kDynB, off_big, kNoStash, kNoStash, 0, 0, 0);
    Seg s{off_gx, pl.Hdp, w ? w->prior0_w : nul, d.D, 0, d.Hd};
    s.trans = 1;
    add_acc_group(pl, RowSpec(0, d.D, pl.Dp), {s}, kDynB, false, true, 1, true);
  }
  // ---- J4: dL/dx = [g_n | g_r | g_z] W_ih (x ELU'(x));  J5: dL/dh += [g_r | g_z | g_n r] W_hh.  J5's MMAs need
  // nothing from J4's epilogue: they are issued right behind J4's and run under it.
  {
    job(JB_ELUB, pl.Dp, kDynA, off_gx, pl.sl.xe, kNoStash, 0, 0, 0);
    Seg s{off_big, 3 * pl.Dp, w ? w->w_ih : nul, d.D, 0, 0};
    s.trans = 1; s.add_k(2 * d.D, d.D, pl.Dp); s.add_k(0, d.D, pl.Dp); s.add_k(d.D, d.D, pl.Dp);
    add_acc_group(pl, RowSpec(0, d.D, pl.Dp), {s}, kDynA, false, true, 1, true);
    job(JB_CARRY, pl.Dp, kDynB, 0, kNoStash, kNoStash, 0, 0, 0);
    Seg h{off_big + (uint32_t)pl.Dp / 8 * PS, 3 * pl.Dp, w ? w->w_hh : nul, d.D, 0, 0};
    h.trans = 1; h.add_k(0, d.D, pl.Dp); h.add_k(d.D, d.D, pl.Dp); h.add_k(2 * d.D, d.D, pl.Dp);
    add_acc_group(pl, RowSpec(0, d.D, pl.Dp), {h}, kDynB, false, true, 0, true);
  }
  // ---- J6: [dL/ds | dL/da] = g_x W_e; tanh-squash backward; then the Gaussian-sample backward of the previous step
  {
    job(JB_EMB, pl.Sp + 16, kDynA, off_gmu, kNoStash, kNoStash, 0, 0, 0);
    RowSpec rs;
    rs.add(0, d.S, pl.Sp);
    rs.add(d.S_in, d.A, 16);
    Seg s{off_gx, pl.Dp, w ? w->embed_w : nul, d.S_in + d.A, 0, d.D};
    s.trans = 1;
    add_acc_group(pl, rs, {s}, kDynA, false, true, 2, true);
  }
  // ---- actor: dL/dy_{L-1} = g_mu W_mu, then per block LayerNorm+ELU backward and dL/dy_{l-1} = g_pre_l W_l
  for (int l = d.La - 1; l >= 0; --l) {
    const uint32_t pb = add_vec(pl, aw ? aw->ln_g[l] : nul, nul, 0, d.Ha, d.Ha, 1.4426950408889634f);
    add_vec(pl, aw ? aw->ln_b[l] : nul, nul, 0, d.Ha, d.Ha, 1.4426950408889634f);
    add_vec(pl, aw ? aw->ln_g[l] : nul, nul, 0, d.Ha, d.Ha);
    job(JB_LNB, d.Ha, 0, off_big, pl.sl.xh[l], kNoStash, (uint32_t)l, pb, l > 0 ? 1u : 0u);
    const int nchunks = (d.Ha + 255) / 256;
    for (int nc = 0; nc < nchunks; ++nc) {
      const int rows = std::min(256, d.Ha - nc * 256);
      Seg s = (l == d.La - 1) ? Seg{off_gmu, 16, aw ? aw->mu_w : nul, d.Ha, 0, d.A}
                              : Seg{off_big, d.Ha, aw ? aw->lin_w[l + 1] : nul, d.Ha, 0, d.Ha};
      s.trans = 1;
      add_acc_group(pl, RowSpec(nc * 256, rows, rows), {s}, nc * 128, false, nc == 0, 1, nc == nchunks - 1);
    }
  }
  pl.boff[6] = add_vec(pl, aw ? aw->bound : nul, nul, 0, d.A, 16);
  for (const StageDesc& st : pl.stages)
    if ((int)st.bytes16 * 16 > pl.stage_bytes) return pl;
  if ((int)pl.stages.size() > kMaxStagesB || (int)pl.jobs.size() > kMaxJobsB || (int)pl.eparams_floats > kMaxEparamsB)
    return pl;
  auto al = [](size_t x) { return (x + 255) / 256 * 256; };
  pl.off_stage_tab = 256;
  pl.off_job_tab = al(pl.off_stage_tab + pl.stages.size() * sizeof(StageDesc));
  pl.off_packops = al(pl.off_job_tab + pl.jobs.size() * sizeof(JobDesc));
  pl.off_eparams = al(pl.off_packops + pl.pack_ops.size() * sizeof(PackOp) + pl.vec_ops.size() * sizeof(VecOp));
  pl.off_data = al(pl.off_eparams + pl.eparams_floats * sizeof(float));
  pl.total_bytes = al(pl.off_data + pl.data_bytes);
  pl.ok = true;
  return pl;
}

bool tc_train_supported(const Dims& d) { return tc_bwd_image_offset(d) != 0; }
static size_t n_blocks(const Dims& d) { return (size_t)((d.B + kTileM - 1) / kTileM) * 2 * (size_t)d.T; }
size_t tc_saved_bytes(const Dims& d) {
  Plan bp = tc_make_bwd_plan(d, nullptr, nullptr);
  return bp.ok ? n_blocks(d) * bp.sl.blk_f : 0;
}
// gradient stash + the dL/dh scratch rows
size_t tc_bwd_workspace_bytes(const Dims& d) {
  Plan bp = tc_make_bwd_plan(d, nullptr, nullptr);
  if (!bp.ok) return 0;
  return n_blocks(d) * bp.sl.blk_g + (size_t)((d.B + kTileM - 1) / kTileM) * kTileM * bp.Dp * sizeof(float) + 256 + 65536;
}

// --------------------------------------------------------------------------
// epilogues
// --------------------------------------------------------------------------
struct BCtx {
  const BwdParams* p;
  uint32_t smem_base, tmem_lane;
  int row, hf, q, sidx, quarter;
  long long b;          // global row
  long long prow;       // padded row index (tile*128 + cta*64 + row): scratch addressing
  bool valid;
  int t;                // forward step being differentiated: state t -> state t+1
  bool first;           // first backward step of the tile (t == H-1): the carried gradients are zero
  const uint8_t* fblk;  // forward stash block of (tile, CTA, t), offset by row*16
  uint8_t* gblk;        // gradient stash block of (tile, CTA, t), offset by row*16
};

// this thread's 4 floats of columns [col, col + 4) (col % 4 == 0) of its row in the fp32 scratch panel of its (tile, CTA)
__device__ __forceinline__ float* scr4(const BCtx& c, int col) {
  const BwdParams& p = *c.p;
  return p.gh_scratch + (((size_t)(c.prow >> 6) * (size_t)(p.Dp >> 2) + (size_t)(col >> 2)) * kRowsB + (size_t)c.row) * 4;
}
__device__ __forceinline__ float4 ldcb4(uint32_t off) { return reinterpret_cast<const float4*>(cb_eparams)[off >> 2]; }
__device__ __forceinline__ void rowpair_sync_b(int quarter) {
  asm volatile("bar.sync %0, %1;" ::"r"((quarter & 1) + 1), "r"(2 * kQ * 32) : "memory");
}
__device__ __forceinline__ void epi_sync_all() { asm volatile("bar.sync 5, %0;" ::"r"(kEpiThreads) : "memory"); }

// Gaussian-sample backward (autograd of rssm.py:179-181) for the step that produced state `ts` (1..H):
//   g_mean = dL/dmean[ts] + g_s,  g_raw = (dL/dstd[ts] + g_s eps[ts-1]) sigmoid(raw),  g_s = carried dL/ds + dL/dstoch[ts]
// sigmoid(raw) = 1 - exp(-softplus(raw)) = 1 - exp(-(std - 0.1)).  Writes the [g_mean | g_raw] panel.  The five
// per-row inputs are loaded ahead of time (SamplePre) by the 8-column owners: threads sidx < Sp/8.
struct SamplePre { float gs[8], gmu[8], gsd[8], e[8], sd[8]; };
__device__ __forceinline__ void sample_load(const BCtx& c, int ts, SamplePre& pf) {
  const BwdParams& p = *c.p;
  if (c.sidx >= p.Sp / 8) return;
  const int col = c.sidx * 8;
  const long long o = (c.b * (p.H + 1) + ts) * p.S + col;
  const float* eps = p.noise + ((long long)(ts - 1) * p.B + c.b) * p.S + col;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool ok = c.valid && col + i < p.S;
    pf.gs[i] = (ok && p.g_stoch) ? p.g_stoch[o + i] : 0.f;
    pf.gmu[i] = (ok && p.g_mean) ? p.g_mean[o + i] : 0.f;
    pf.gsd[i] = (ok && p.g_std) ? p.g_std[o + i] : 0.f;
    pf.e[i] = ok ? eps[i] : 0.f;
    pf.sd[i] = ok ? p.std_seq[o + i] : 1.f;
  }
}
__device__ __forceinline__ void sample_bwd(const BCtx& c, bool use_carry, const SamplePre& pf) {
  const BwdParams& p = *c.p;
  if (c.sidx >= p.Sp / 8) return;
  const int col = c.sidx * 8;
  float gm[8], gr[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool ok = c.valid && col + i < p.S;
    float gs = 0.f;
    if (use_carry)
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(gs) : "r"(c.smem_base + p.off_carry + ((col + i) * kRowsB + c.row) * 4));
    gs += pf.gs[i];
    const float sig = 1.f - __expf(-(pf.sd[i] - 0.1f));
    gm[i] = ok ? pf.gmu[i] + gs : 0.f;
    gr[i] = ok ? (pf.gsd[i] + gs * pf.e[i]) * sig : 0.f;
  }
  st_chunk(c.smem_base + p.off_go + c.sidx * kPSB + c.row * 16, gm);
  st_chunk(c.smem_base + p.off_go + (p.Sp / 8 + c.sidx) * kPSB + c.row * 16, gr);
}

// same arithmetic with the five per-row inputs read in place (two columns at a time, packed at once): used by the embed
// job, where the carried dL/ds is in shared memory and the rows are L2-resident
__device__ __forceinline__ void sample_bwd_direct(const BCtx& c, int ts) {
  const BwdParams& p = *c.p;
  if (c.sidx >= p.Sp / 8) return;
  const int col = c.sidx * 8;
  const long long o = (c.b * (p.H + 1) + ts) * p.S + col;
  const float* eps = p.noise + ((long long)(ts - 1) * p.B + c.b) * p.S + col;
  uint32_t pm[4], pr[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float gm2[2], gr2[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int i = 2 * j + e;
      const bool ok = c.valid && col + i < p.S;
      float gs;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(gs) : "r"(c.smem_base + p.off_carry + ((col + i) * kRowsB + c.row) * 4));
      gs += (ok && p.g_stoch) ? p.g_stoch[o + i] : 0.f;
      const float sd = ok ? p.std_seq[o + i] : 1.f;
      const float sig = 1.f - __expf(-(sd - 0.1f));
      gm2[e] = ok ? ((p.g_mean ? p.g_mean[o + i] : 0.f) + gs) : 0.f;
      gr2[e] = ok ? ((p.g_std ? p.g_std[o + i] : 0.f) + gs * eps[i]) * sig : 0.f;
    }
    pm[j] = pack_bf16x2(gm2[0], gm2[1]);
    pr[j] = pack_bf16x2(gr2[0], gr2[1]);
  }
  const uint32_t d0 = c.smem_base + p.off_go + c.sidx * kPSB + c.row * 16;
  const uint32_t d1 = c.smem_base + p.off_go + (p.Sp / 8 + c.sidx) * kPSB + c.row * 16;
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(d0), "r"(pm[0]), "r"(pm[1]), "r"(pm[2]), "r"(pm[3]) : "memory");
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(d1), "r"(pr[0]), "r"(pr[1]), "r"(pr[2]), "r"(pr[3]) : "memory");
}

__device__ __forceinline__ void prefetch_l2(const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }
__device__ __forceinline__ void bulk_prefetch_l2(const void* ptr, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ uint4 ldg16(const void* ptr) { return *reinterpret_cast<const uint4*>(ptr); }

// Every epilogue needs recorded activations from HBM.  The epilogue warps idle while the job's MMAs run, so each job
// has a PREFETCH half that issues its global loads BEFORE the accumulator wait (the round trip overlaps the MMAs),
// and the producer lane pulls the next step's stash block into L2 one step ahead.
constexpr int kSubMax = 4;   // 8-column sub-groups per thread of a <= 256-wide accumulator (13 per lane half / kQ warps)

// ---- g = acc * ELU'(a), a = the recorded post-ELU activation (ELU'(z) = 1 for z > 0, else a + 1) -> bf16 panel
struct EluPre { uint4 a[kSubMax]; };
__device__ __forceinline__ void elu_prefetch(const BCtx& c, const JobDesc& jb, EluPre& pf) {
  const int hw = jb.n / 2;
#pragma unroll
  for (int o = 0; o < kSubMax; ++o) {
    const int sub = c.q + o * kQ;
    if (sub * 8 < hw) pf.a[o] = ldg16(c.fblk + jb.st0 + ((c.hf * hw) / 8 + sub) * kStashChunk);
  }
}
__device__ __forceinline__ void epib_elu(const BCtx& c, const JobDesc& jb, const EluPre& pf) {
  const int hw = jb.n / 2;
#pragma unroll
  for (int o = 0; o < kSubMax; ++o) {
    const int sub = c.q + o * kQ;
    if (sub * 8 < hw) {
      const int f0 = c.hf * hw + sub * 8;
      float v[8], a[8];
      tmem_ld8(c.tmem_lane + jb.tmem_col + sub * 8, v);
      unpack8(pf.a[o], a);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] *= (a[i] > 0.f) ? 1.f : (a[i] + 1.f);
      st_chunk(c.smem_base + jb.dst_off + (f0 / 8) * kPSB + c.row * 16, v);
    }
  }
}

// ---- GRUCell backward (torch.nn.GRUCell as used at rssm.py:66; h' = (1-z) n + z h, n = tanh(i_n + r h_n)):
// accumulator = g_hid W1 (prior-head path of dL/dh'); the scratch row holds the rest of dL/dh' (carried gradient +
// upstream dL/ddeter[t+1], put there by the previous backward step / the tile initialisation).  Writes the gate
// gradient panels [g_n | g_r | g_z | g_n r] and parks the direct term dL/dh' z in the scratch row.
// L2 prefetch (no register cost) of the fp32 rows the LATER jobs of this step read: the upstream deter gradient of
// index t (carry job), the sample-backward inputs of index t and the actions (embed job) - issued ~15-30K cycles ahead
__device__ __forceinline__ void prefetch_step_rows(const BCtx& c) {
  const BwdParams& p = *c.p;
  if (!c.valid) return;
  const long long T1 = p.H + 1;
  if (p.g_deter) {
    const int hw = p.Dp / 2;
    for (int sub = c.q; sub * 8 < hw; sub += kQ) {
      const int col = c.hf * hw + sub * 8;
      if (col < p.D) prefetch_l2(p.g_deter + (c.b * T1 + c.t) * p.D + col);
    }
  }
  if (c.sidx < p.Sp / 8 && c.sidx * 8 < p.S && c.t > 0) {
    const long long o = (c.b * T1 + c.t) * p.S + c.sidx * 8;
    prefetch_l2(p.std_seq + o);
    if (p.g_stoch) prefetch_l2(p.g_stoch + o);
    if (p.g_mean) prefetch_l2(p.g_mean + o);
    if (p.g_std) prefetch_l2(p.g_std + o);
    prefetch_l2(p.noise + ((long long)(c.t - 1) * p.B + c.b) * p.S + c.sidx * 8);
  }
  if (c.sidx == 0) prefetch_l2(p.actions + (c.b * p.H + c.t) * p.A);
}
__device__ __forceinline__ uint32_t u4c(const uint4& u, int j) { return j == 0 ? u.x : (j == 1 ? u.y : (j == 2 ? u.z : u.w)); }
__device__ __forceinline__ float f4c(const float4& u, int j) { return j == 0 ? u.x : (j == 1 ? u.y : (j == 2 ? u.z : u.w)); }
// Register-lean form: nothing is held across the accumulator wait (everything this epilogue reads is L2-resident: the
// stash block was bulk-prefetched a step ahead, the scratch row was written by this thread one step earlier, the h row
// prefetched by the previous carry epilogue), the recorded gates stay PACKED (bf16 pairs) until the column pair that
// needs them (h_t comes from the recorded bf16 [h | s] panel), and every output pair is packed as soon as it exists - the previous form kept 56 unpacked inputs and 40
// outputs live and spilled ~100 registers per sub-group to local memory (L2 round trips: the L1D is ~1 KB here).
__device__ __forceinline__ void epib_gru(const BCtx& c, const JobDesc& jb) {
  const BwdParams& p = *c.p;
  const int hw = p.Dp / 2;
  prefetch_step_rows(c);
  const uint32_t gate = c.smem_base + jb.dst_off + c.row * 16;
  const uint32_t gstride = (uint32_t)(p.Dp / 8) * kPSB;
#pragma unroll 1
  for (int sub = c.q; sub * 8 < hw; sub += kQ) {
    const int col = c.hf * hw + sub * 8;
    float acc[8];
    tmem_ld8(c.tmem_lane + jb.tmem_col + sub * 8, acc);
    const uint8_t* f = c.fblk + (col / 8) * kStashChunk;
    const uint4 ur = ldg16(f + p.sl.gr), uz = ldg16(f + p.sl.gz), un = ldg16(f + p.sl.gn), uh = ldg16(f + p.sl.ghn);
    float* const scr_a = scr4(c, col);
    float* const scr_b = scr4(c, col + 4);
    const float4 cr0 = *reinterpret_cast<const float4*>(scr_a), cr1 = *reinterpret_cast<const float4*>(scr_b);
    const uint4 up = ldg16(f + p.sl.feat);   // h_t of these columns from the recorded [h | s] panel (bf16), not the fp32 rows
    tmem_ld_wait();
    uint32_t pn[4], pr[4], pz[4], ph[4];
    float od[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {     // columns 2j, 2j + 1
      const uint32_t wr = u4c(ur, j), wz = u4c(uz, j), wn = u4c(un, j), wh = u4c(uh, j), wp = u4c(up, j);
      float o_n[2], o_r[2], o_z[2], o_h[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int i = 2 * j + e;
        const float r = __uint_as_float(e ? (wr & 0xFFFF0000u) : (wr << 16));
        const float z = __uint_as_float(e ? (wz & 0xFFFF0000u) : (wz << 16));
        const float n = __uint_as_float(e ? (wn & 0xFFFF0000u) : (wn << 16));
        const float hn = __uint_as_float(e ? (wh & 0xFFFF0000u) : (wh << 16));
        const float cr = i < 4 ? f4c(cr0, i) : f4c(cr1, i - 4);
        const float hp = __uint_as_float(e ? (wp & 0xFFFF0000u) : (wp << 16));
        const float gh = acc[i] + cr;
        const float gnp = gh * (1.f - z) * (1.f - n * n);
        od[i] = gh * z;
        o_n[e] = gnp;
        o_z[e] = gh * (hp - n) * z * (1.f - z);
        o_h[e] = gnp * r;
        o_r[e] = gnp * hn * r * (1.f - r);
      }
      pn[j] = pack_bf16x2(o_n[0], o_n[1]); pr[j] = pack_bf16x2(o_r[0], o_r[1]);
      pz[j] = pack_bf16x2(o_z[0], o_z[1]); ph[j] = pack_bf16x2(o_h[0], o_h[1]);
    }
    const uint32_t dst = gate + (col / 8) * kPSB;
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(pn[0]), "r"(pn[1]), "r"(pn[2]), "r"(pn[3]) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst + gstride), "r"(pr[0]), "r"(pr[1]), "r"(pr[2]), "r"(pr[3]) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst + 2 * gstride), "r"(pz[0]), "r"(pz[1]), "r"(pz[2]), "r"(pz[3]) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst + 3 * gstride), "r"(ph[0]), "r"(ph[1]), "r"(ph[2]), "r"(ph[3]) : "memory");
    *reinterpret_cast<float4*>(scr_a) = make_float4(od[0], od[1], od[2], od[3]);
    *reinterpret_cast<float4*>(scr_b) = make_float4(od[4], od[5], od[6], od[7]);
  }
}

// upstream gradient of the deter sequence at index `ts` for 8 columns of this row (0 where absent / padding)
__device__ __forceinline__ void load_g_deter(const BCtx& c, int ts, int col, float (&up)[8]) {
  const BwdParams& p = *c.p;
#pragma unroll
  for (int i = 0; i < 8; ++i) up[i] = 0.f;
  if (!p.g_deter || !c.valid) return;
  const float* g = p.g_deter + (c.b * (p.H + 1) + ts) * p.D + col;
  if (col + 8 <= p.D && (p.D & 3) == 0) {
    const float4 a = *reinterpret_cast<const float4*>(g), b4 = *reinterpret_cast<const float4*>(g + 4);
    up[0] = a.x; up[1] = a.y; up[2] = a.z; up[3] = a.w; up[4] = b4.x; up[5] = b4.y; up[6] = b4.z; up[7] = b4.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (col + i < p.D) up[i] = g[i];
  }
}

// tile start: the scratch rows take the upstream gradient of the LAST deter state (the carried part is zero)
__device__ __forceinline__ void carry_init(const BCtx& c) {
  const BwdParams& p = *c.p;
  const int hw = p.Dp / 2;
  for (int sub = c.q; sub * 8 < hw; sub += kQ) {
    const int col = c.hf * hw + sub * 8;
    float up[8];
    load_g_deter(c, p.H, col, up);
    *reinterpret_cast<float4*>(scr4(c, col)) = make_float4(up[0], up[1], up[2], up[3]);
    *reinterpret_cast<float4*>(scr4(c, col + 4)) = make_float4(up[4], up[5], up[6], up[7]);
  }
}

// ---- dL/dh_t = [g_r | g_z | g_n r] W_hh + dL/dh' z (+ the upstream gradient of deter[t]): carried to the previous
// step through the scratch row, or - at t = 0 - the gradient of the start state
struct CarryPre { float4 s[kSubMax][2]; };
__device__ __forceinline__ void carry_prefetch(const BCtx& c, CarryPre& pf) {
  const BwdParams& p = *c.p;
  const int hw = p.Dp / 2;
#pragma unroll
  for (int o = 0; o < kSubMax; ++o) {
    const int sub = c.q + o * kQ;
    if (sub * 8 < hw) {
      pf.s[o][0] = *reinterpret_cast<const float4*>(scr4(c, c.hf * hw + sub * 8));
      pf.s[o][1] = *reinterpret_cast<const float4*>(scr4(c, c.hf * hw + sub * 8 + 4));
    }
  }
}
__device__ __forceinline__ void epib_carry(const BCtx& c, const JobDesc& jb, const CarryPre& pf) {
  const BwdParams& p = *c.p;
  const int hw = p.Dp / 2;
#pragma unroll
  for (int o = 0; o < kSubMax; ++o) {
    const int sub = c.q + o * kQ;
    if (sub * 8 < hw) {
      const int col = c.hf * hw + sub * 8;
      float acc[8], up[8];
      tmem_ld8(c.tmem_lane + jb.tmem_col + sub * 8, acc);
      load_g_deter(c, c.t, col, up);
      tmem_ld_wait();
      acc[0] += pf.s[o][0].x + up[0]; acc[1] += pf.s[o][0].y + up[1]; acc[2] += pf.s[o][0].z + up[2]; acc[3] += pf.s[o][0].w + up[3];
      acc[4] += pf.s[o][1].x + up[4]; acc[5] += pf.s[o][1].y + up[5]; acc[6] += pf.s[o][1].z + up[6]; acc[7] += pf.s[o][1].w + up[7];
      if (c.t > 0) {
        *reinterpret_cast<float4*>(scr4(c, col)) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        *reinterpret_cast<float4*>(scr4(c, col + 4)) = make_float4(acc[4], acc[5], acc[6], acc[7]);
      } else if (c.valid) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (col + i < p.D) p.g_deter0[c.b * p.D + col + i] = acc[i];
      }
    }
  }
}

// ---- [dL/ds_t | dL/da_t] = g_x W_e.  dL/ds_t is carried (shared memory) or, at t = 0, the start-state gradient;
// dL/da_t goes through the tanh squash (actor.py:49-50: a = tanh(mu / scale) bound) into the g_mu panel + stash.
// Then - all warps - the Gaussian-sample backward of step t-1, which needs the carried dL/ds.
struct EmbPre { float act[8]; };
__device__ __forceinline__ void emb_prefetch(const BCtx& c, const JobDesc& jb, EmbPre& pf) {
  const BwdParams& p = *c.p;
  // the (single) sub-group of this thread that holds action columns, if any
  const int hw = jb.n / 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) pf.act[i] = 0.f;
  for (int sub = c.q; sub * 8 < hw; sub += kQ) {
    const int col = c.hf * hw + sub * 8;
    if (col >= p.Sp && c.valid) {
      const int a0 = col - p.Sp;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (a0 + i < p.A) pf.act[i] = p.actions[(c.b * p.H + c.t) * p.A + a0 + i];
    }
  }
}
__device__ __forceinline__ void epib_emb(const BCtx& c, const JobDesc& jb, const EmbPre& pf) {
  const BwdParams& p = *c.p;
  const int hw = jb.n / 2;
  // the sample-backward inputs of step t-1 (five 8-column rows, L2-resident since the GRU job's prefetch) are requested
  // first, so that their round trip overlaps the accumulator read-out below
  SamplePre sp;
#pragma unroll
  for (int i = 0; i < 8; ++i) { sp.gs[i] = 0.f; sp.gmu[i] = 0.f; sp.gsd[i] = 0.f; sp.e[i] = 0.f; sp.sd[i] = 1.f; }
  if (c.t > 0) sample_load(c, c.t, sp);
  for (int sub = c.q; sub * 8 < hw; sub += kQ) {
    const int col = c.hf * hw + sub * 8;
    float acc[8];
    tmem_ld8(c.tmem_lane + jb.tmem_col + sub * 8, acc);
    tmem_ld_wait();
    if (col < p.Sp) {
      if (c.t > 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(c.smem_base + p.off_carry + ((col + i) * kRowsB + c.row) * 4), "f"(acc[i])
                       : "memory");
      } else if (c.valid) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (col + i < p.S) {
            float g = acc[i];
            if (p.g_stoch) g += p.g_stoch[(c.b * (p.H + 1)) * p.S + col + i];
            p.g_stoch0[c.b * p.S + col + i] = g;
          }
      }
    } else {
      const int a0 = col - p.Sp;
      float gm[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int ai = a0 + i;
        gm[i] = 0.f;
        if (ai < p.A && c.valid) {
          const float bd = cb_eparams[p.bound_off + ai];
          const float th = (bd != 0.f) ? pf.act[i] / bd : 0.f;
          gm[i] = acc[i] * bd * (1.f - th * th) * p.inv_action_scale;
        }
      }
      st_chunk_sg_cs(c.smem_base + jb.dst_off + (a0 / 8) * kPSB + c.row * 16, c.gblk + p.sl.gmu + (a0 / 8) * kStashChunk, gm);
    }
  }
  if (c.t > 0) {
    // the sample-backward inputs are loaded HERE, not before the accumulator wait: holding their 40 registers across the
    // embed epilogue spilled them to local memory; the rows were pulled into L2 by prefetch_step_rows during the GRU job
    epi_sync_all();
    sample_bwd(c, true, sp);
  }
}

// ---- LayerNorm(eps 1e-5, affine) + ELU backward of actor block `layer` (actor.py:23-36), from the recorded x-hat, rstd:
//   u = gamma xh + beta, g_u = g_y ELU'(u), g_xh = g_u gamma, g_pre = rstd (g_xh - mean(g_xh) - xh mean(g_xh xh)).
// g_u and g_pre go to the gradient stash (column sums / weight gradients after the scan); g_pre also becomes the A panel
// of the next dgrad GEMM unless this is block 0 (the actor input is detached, rssm.py:135).
// A thread owns up to two 32-column sub-groups; their x-hat (8 x 16 B, packed bf16) is fetched before the wait and
// kept in registers across both passes.
struct LnPre { uint4 xh[2][4]; float rstd; };
__device__ __forceinline__ bool ln_sub(const BwdParams& p, int li, int hf, int& f0, int& tcol) {
  const int w0 = min(256, p.Ha), nsub0 = w0 / 64;     // 32-column sub-groups per lane half of MMA group 0
  if (li < nsub0) { f0 = hf * (w0 / 2) + li * 32; tcol = li * 32; return true; }
  const int l2 = li - nsub0, w1 = p.Ha - 256;
  if (w1 <= 0 || l2 >= w1 / 64) return false;
  f0 = 256 + hf * (w1 / 2) + l2 * 32; tcol = 128 + l2 * 32;
  return true;
}
__device__ __forceinline__ void ln_prefetch(const BCtx& c, const JobDesc& jb, LnPre& pf) {
  const uint8_t* xh_f = c.fblk + jb.st0;
#pragma unroll
  for (int o = 0; o < 2; ++o) {
    int f0, tcol;
    if (ln_sub(*c.p, c.q + o * kQ, c.hf, f0, tcol)) {
#pragma unroll
      for (int k = 0; k < 4; ++k) pf.xh[o][k] = ldg16(xh_f + (f0 / 8 + k) * kStashChunk);
    }
  }
  pf.rstd = *reinterpret_cast<const float*>(c.fblk - c.row * 16 + c.p->sl.rstd + (jb.layer * kRowsB + c.row) * 4);
}
__device__ __forceinline__ void epib_ln(const BCtx& c, const JobDesc& jb, const LnPre& pf) {
  const BwdParams& p = *c.p;
  const uint32_t gsc = jb.p_off, bsc = gsc + p.Ha, gam = bsc + p.Ha;   // gamma*log2e, beta*log2e, gamma
  uint8_t* gu_g = c.gblk + p.sl.gu[jb.layer];
  uint8_t* gp_g = c.gblk + p.sl.gp[jb.layer];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int o = 0; o < 2; ++o) {
    int f0, tcol;
    if (!ln_sub(p, c.q + o * kQ, c.hf, f0, tcol)) continue;
#pragma unroll
    for (int hk = 0; hk < 2; ++hk) {
      float v[16];
      tmem_ld16(c.tmem_lane + jb.tmem_col + tcol + hk * 16, v);
      tmem_ld_wait();
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int k = hk * 2 + kk;
        float xh[8], gu[8];
        unpack8(pf.xh[o][k], xh);
#pragma unroll
        for (int i = 0; i < 8; i += 4) {
          const float4 g4 = ldcb4(gsc + f0 + k * 8 + i), b4 = ldcb4(bsc + f0 + k * 8 + i), w4 = ldcb4(gam + f0 + k * 8 + i);
          const float gs[4] = {g4.x, g4.y, g4.z, g4.w}, bs[4] = {b4.x, b4.y, b4.z, b4.w}, gw[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float us = fmaf(gs[e], xh[i + e], bs[e]);                          // u * log2(e)
            const float g_u = v[kk * 8 + i + e] * ex2_fast(fminf(us, 0.f));          // ELU'(u) = 1 (u > 0) | exp(u)
            const float g_x = g_u * gw[e];
            gu[i + e] = g_u;
            s1 += g_x;
            s2 = fmaf(g_x, xh[i + e], s2);
          }
        }
        st_chunk_g_cs(gu_g + (f0 / 8 + k) * kStashChunk, gu);
      }
    }
  }
  // the 8 threads of a row (2 lane halves x kQ warps) exchange their partial sums
  const uint32_t scratch = c.smem_base + p.off_red;
  asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(scratch + (c.sidx * kRowsB + c.row) * 8), "f"(s1), "f"(s2) : "memory");
  rowpair_sync_b(c.quarter);
  s1 = 0.f; s2 = 0.f;
#pragma unroll
  for (int k = 0; k < 2 * kQ; ++k) {
    float a, b;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(a), "=f"(b) : "r"(scratch + (k * kRowsB + c.row) * 8) : "memory");
    s1 += a; s2 += b;
  }
  const float inv = 1.f / (float)p.Ha;
  const float m1 = s1 * inv, m2 = s2 * inv;
  const float rstd = pf.rstd;
#pragma unroll
  for (int o = 0; o < 2; ++o) {
    int f0, tcol;
    if (!ln_sub(p, c.q + o * kQ, c.hf, f0, tcol)) continue;
#pragma unroll
    for (int hk = 0; hk < 2; ++hk) {
      float v[16];
      tmem_ld16(c.tmem_lane + jb.tmem_col + tcol + hk * 16, v);
      tmem_ld_wait();
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int k = hk * 2 + kk;
        float xh[8], gp[8];
        unpack8(pf.xh[o][k], xh);
#pragma unroll
        for (int i = 0; i < 8; i += 4) {
          const float4 g4 = ldcb4(gsc + f0 + k * 8 + i), b4 = ldcb4(bsc + f0 + k * 8 + i), w4 = ldcb4(gam + f0 + k * 8 + i);
          const float gs[4] = {g4.x, g4.y, g4.z, g4.w}, bs[4] = {b4.x, b4.y, b4.z, b4.w}, gw[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float us = fmaf(gs[e], xh[i + e], bs[e]);
            const float g_x = v[kk * 8 + i + e] * ex2_fast(fminf(us, 0.f)) * gw[e];
            gp[i + e] = rstd * (g_x - m1 - xh[i + e] * m2);
          }
        }
        if (jb.aux & 1u) st_chunk_sg_cs(c.smem_base + jb.dst_off + (f0 / 8 + k) * kPSB + c.row * 16, gp_g + (f0 / 8 + k) * kStashChunk, gp);
        else st_chunk_g_cs(gp_g + (f0 / 8 + k) * kStashChunk, gp);
      }
    }
  }
}

// --------------------------------------------------------------------------
// the persistent BPTT kernel
// --------------------------------------------------------------------------
template <bool kTrace>
__global__ void __launch_bounds__(kThreadsB, 1) tc_bptt_kernel(const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t kSlot = p.slot_bytes;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t ring = smem_base + p.ring_off;
  const uint32_t bars = smem_base + p.bar_off;
  // barrier map (8 B each): full[16] @0 | empty[16] @128 | acc_full[2] @384 | acc_empty[2] @400 | tmem ptr @416 | ready @424
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 128u + 8u * s; };
  auto accf_bar = [&](uint32_t j) { return bars + 384u + 8u * (j & 1u); };
  auto acce_bar = [&](uint32_t j) { return bars + 400u + 8u * (j & 1u); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + p.bar_off + 416);
  volatile uint32_t* ready_ctr = reinterpret_cast<volatile uint32_t*>(smem + p.bar_off + 424);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int unit = (int)(blockIdx.x >> 1), nunits = (int)(gridDim.x >> 1);
  const int ngroups = p.ntiles;

  // per-stage MMA records (as in the rollout kernel): x = A descriptor low word, y = instruction descriptor,
  // z = B LBO field << 16 | nk16 << 8 | flags, w = tmem column | dep << 16
  const uint4* stab = reinterpret_cast<const uint4*>(smem + p.bar_off + kBarBytesB);
  for (int i = threadIdx.x; i < p.nstages; i += kThreadsB) {
    const uint4 raw = cb_stages[i];
    const StageDesc st = *reinterpret_cast<const StageDesc*>(&raw);
    uint4 rec;
    rec.x = (uint32_t)make_smem_desc(smem_base + (uint32_t)st.a_off16 * 16u, kPSB, 128u);
    rec.y = make_idesc_bf16(128, st.n);
    rec.z = ((uint32_t)(st.n / 2) << 16) | ((uint32_t)st.nk16 << 8) | st.flags;
    rec.w = (uint32_t)st.tmem_col | ((uint32_t)st.dep << 16);
    reinterpret_cast<uint4*>(smem + p.bar_off + kBarBytesB)[i] = rec;
  }
  if (threadIdx.x == 0) {
    *ready_ctr = 0u;
    for (int s = 0; s < p.nslots; ++s) {
      mbar_init(full_bar(s), rank == 0 ? 2 : 1);   // own half (arrive + tx bytes) and the peer's forwarded arrival
      mbar_init(empty_bar(s), 1);
    }
    for (uint32_t j = 0; j < 2; ++j) {
      mbar_init(accf_bar(j), 1);
      mbar_init(acce_bar(j), (kEpiThreads / 32) * 2);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) { tmem_alloc2(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512); tmem_relinquish2(); }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kProducerWarp) {
    if (lane == 0) {
      uint32_t slot = 0, phase = 0;
      // the fields the epilogues read (x-hat of every block; x, gates, prior hidden, rstd) are two contiguous ranges
      const uint32_t pf0 = p.sl.xh[0], pf0_bytes = (uint32_t)p.La * (uint32_t)(p.Ha / 8) * kStashChunk;
      const uint32_t pf1 = p.sl.xe, pf1_bytes = p.sl.blk_f - p.sl.xe;
      auto pull_ln = [&](size_t blk) {
        const uint8_t* b0 = p.stash_f + blk * p.sl.blk_f;
        for (uint32_t o = 0; o < pf0_bytes; o += 32768u) bulk_prefetch_l2(b0 + pf0 + o, min(32768u, pf0_bytes - o));
      };
      auto pull_dyn = [&](size_t blk) {
        const uint8_t* b0 = p.stash_f + blk * p.sl.blk_f;
        for (uint32_t o = 0; o < pf1_bytes; o += 32768u) bulk_prefetch_l2(b0 + pf1 + o, min(32768u, pf1_bytes - o));
      };
      for (int grp = unit; grp < ngroups; grp += nunits)
        for (int ti = 0; ti < p.H; ++ti) {
          const size_t blk0 = ((size_t)grp * 2 + rank) * (size_t)p.H;
          const size_t blk = blk0 + p.H - 1 - ti;
          if (ti == 0) pull_dyn(blk);                        // tile start: no lead time for the first step
          int jidx = -1;
          for (int s = 0; s < p.nstages; ++s) {
            const uint4 raw = cb_stages[s];
            if ((raw.y >> 24) & ST_FIRST_OF_JOB) {
              // operands are pulled into L2 a few jobs (tens of microseconds) ahead of the epilogue that reads them - a
              // whole step ahead they would be evicted again by the ~100 MB every step streams through the L2:
              // x-hat of this step's LayerNorm jobs while the GRU job runs, the dynamics fields of the NEXT step
              // while the second LayerNorm job runs
              ++jidx;
              if (jidx == 1) pull_ln(blk);
              if (jidx == p.njobs - 2 && ti + 1 < p.H) pull_dyn(blk - 1);
            }
            const uint32_t bytes = ((raw.y & 0xFFFFu) * 16u) / 2u;
            mbar_wait(empty_bar(slot), phase ^ 1u);
            mbar_arrive_expect_tx(full_bar(slot), bytes);
            bulk_g2s(ring + slot * kSlot, p.stage_data + (size_t)raw.x * 16u + (size_t)rank * bytes, bytes, full_bar(slot));
            if (++slot == (uint32_t)p.nslots) { slot = 0; phase ^= 1u; }
          }
        }
    }
  } else if (warp == kMmaWarp) {
    if (lane == 0 && rank != 0) {
      // peer CTA: forward "my half of the stage has landed" to the leader's full barrier
      uint32_t slot = 0, phase = 0;
      for (int grp = unit; grp < ngroups; grp += nunits) {
        const int total = p.H * p.nstages;
        for (int i = 0; i < total; ++i) {
          mbar_wait(full_bar(slot), phase);
          mbar_arrive_remote(mapa(full_bar(slot), 0));
          if (++slot == (uint32_t)p.nslots) { slot = 0; phase ^= 1u; }
        }
      }
    } else if (lane == 0) {
      uint32_t slot = 0, job = 0, g = 0, ready_seen = 0;
      const uint32_t bslot0 = (ring >> 4) & 0x3FFFu, bslot_step = kSlot >> 4;
      uint32_t bslot = bslot0;
      for (int grp = unit; grp < ngroups; grp += nunits) {
        mbar_arrive(accf_bar(job));                    // INIT job (sample backward of the last step) has no MMA
        mbar_arrive_remote(mapa(accf_bar(job), 1));
        ++job;
        const bool tr = kTrace && p.trace && blockIdx.x == 0 && grp == unit;
        for (int ti = 0; ti < p.H; ++ti) {
          uint4 rec_next = stab[0];
          int jidx = -1;
          for (int s = 0; s < p.nstages; ++s) {
            const uint4 rec = rec_next;
            if (s + 1 < p.nstages) rec_next = stab[s + 1];
            const uint32_t flags = rec.z & 0xFFu;
            if (flags & ST_FIRST_OF_JOB) {
              ++jidx;
              const uint32_t dep = rec.w >> 16;
              if (dep) {
                const uint32_t dj = job - dep;
                mbar_wait(acce_bar(dj), (dj >> 1) & 1u);
              }
              if (kTrace && tr && ti < 4) p.trace[((ti * p.njobs) + jidx) * 4 + 0] = clock64();
            }
            if (g >= ready_seen) {
              do { ready_seen = *ready_ctr; } while (ready_seen <= g);
            }
            ++g;
            tc_fence_after();
            const uint32_t tm = tmem_base + (rec.w & 0xFFFFu);
            const uint32_t idesc = rec.y;
            uint32_t alo = rec.x;
            uint32_t blo = bslot | (rec.z & 0xFFFF0000u);
            const uint32_t bstep = (rec.z >> 16) * 2u;
            constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);
            constexpr uint32_t kAStep = (2u * kPSB) >> 4;
            uint32_t acc = (flags & ST_FIRST_ACC) ? 0u : 1u;
            const uint32_t nk = (rec.z >> 8) & 0xFFu;
            uint32_t i = 0;
            for (; i + 4 <= nk; i += 4) {
              mma_bf16_lohi2(tm, alo, kDescHi, blo, kDescHi, idesc, acc);
              mma_bf16_lohi2(tm, alo + kAStep, kDescHi, blo + bstep, kDescHi, idesc, 1u);
              mma_bf16_lohi2(tm, alo + 2 * kAStep, kDescHi, blo + 2 * bstep, kDescHi, idesc, 1u);
              mma_bf16_lohi2(tm, alo + 3 * kAStep, kDescHi, blo + 3 * bstep, kDescHi, idesc, 1u);
              alo += 4 * kAStep;
              blo += 4 * bstep;
              acc = 1u;
            }
            for (; i < nk; ++i) {
              mma_bf16_lohi2(tm, alo, kDescHi, blo, kDescHi, idesc, acc);
              alo += kAStep;
              blo += bstep;
              acc = 1u;
            }
            tc_commit2(empty_bar(slot), 3);
            if (flags & ST_LAST_OF_JOB) {
              tc_commit2(accf_bar(job), 3);
              if (kTrace && tr && ti < 4) p.trace[((ti * p.njobs) + jidx) * 4 + 1] = clock64();
              ++job;
            }
            bslot += bslot_step;
            if (++slot == (uint32_t)p.nslots) { slot = 0; bslot = bslot0; }
          }
        }
      }
    }
  } else if (warp == kReadyWarpB) {
    if (lane == 0 && rank == 0) {
      uint32_t slot = 0, phase = 0, n = 0;
      for (int grp = unit; grp < ngroups; grp += nunits) {
        const int total = p.H * p.nstages;
        for (int i = 0; i < total; ++i) {
          mbar_wait(full_bar(slot), phase);
          *ready_ctr = ++n;
          if (++slot == (uint32_t)p.nslots) { slot = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    BCtx c;
    c.p = &p;
    c.smem_base = smem_base;
    c.quarter = warp & 3;
    c.hf = c.quarter >> 1;
    c.q = warp >> 2;
    c.sidx = c.hf * kQ + c.q;
    c.row = (c.quarter & 1) * 32 + lane;
    c.tmem_lane = tmem_base + ((uint32_t)(c.quarter * 32) << 16);
    uint32_t job = 0;
#define WMRL_JOB_DONE_B(job)                                                            \
  do {                                                                                  \
    __syncwarp();                                                                       \
    if (lane == 0) {                                                                    \
      if (rank != 0) mbar_arrive_remote(mapa(acce_bar(job), 0));                        \
      else mbar_arrive(acce_bar(job));                                                  \
    }                                                                                   \
  } while (0)
    for (int grp = unit; grp < ngroups; grp += nunits) {
      c.prow = (long long)grp * kTileM + (long long)rank * kRowsB + c.row;
      c.b = c.prow;
      c.valid = c.b < p.B;
      const size_t blk0 = ((size_t)grp * 2 + rank) * (size_t)p.H;
      // INIT: the Gaussian-sample backward of the last step (no carried dL/ds yet)
      mbar_wait_backoff_quiet(accf_bar(job), (job >> 1) & 1u);
      c.t = p.H - 1;
      c.first = true;
      {
        SamplePre sp;
        sample_load(c, p.H, sp);
        sample_bwd(c, false, sp);
      }
      carry_init(c);
      fence_proxy_async_smem();
      WMRL_JOB_DONE_B(job);
      ++job;
      for (int ti = 0; ti < p.H; ++ti) {
        c.t = p.H - 1 - ti;
        c.first = ti == 0;
        c.fblk = p.stash_f + (blk0 + c.t) * p.sl.blk_f + c.row * 16;
        c.gblk = p.stash_g + (blk0 + c.t) * p.sl.blk_g + c.row * 16;
        const bool tr = kTrace && p.trace && blockIdx.x == 0 && grp == unit && threadIdx.x == 0 && ti < 4;
        for (int jb_i = 0; jb_i < p.njobs; ++jb_i) {
          const JobDesc jb = cb_jobs[jb_i];
          if (jb.type == JB_LNB) {
            LnPre pf;
            ln_prefetch(c, jb, pf);
            mbar_wait_backoff_quiet(accf_bar(job), (job >> 1) & 1u);
            tc_fence_after();
            if (kTrace && tr) p.trace[((ti * p.njobs) + jb_i) * 4 + 2] = clock64();
            epib_ln(c, jb, pf);
          } else if (jb.type == JB_ELUB) {
            EluPre pf;
            elu_prefetch(c, jb, pf);
            mbar_wait_backoff_quiet(accf_bar(job), (job >> 1) & 1u);
            tc_fence_after();
            if (kTrace && tr) p.trace[((ti * p.njobs) + jb_i) * 4 + 2] = clock64();
            epib_elu(c, jb, pf);
          } else if (jb.type == JB_GRUB) {
            mbar_wait_backoff_quiet(accf_bar(job), (job >> 1) & 1u);
            tc_fence_after();
            if (kTrace && tr) p.trace[((ti * p.njobs) + jb_i) * 4 + 2] = clock64();
            epib_gru(c, jb);
          } else if (jb.type == JB_CARRY) {
            // (its scratch rows were written by this thread in the GRU epilogue of this step: program order)
            CarryPre pf;
            carry_prefetch(c, pf);
            mbar_wait_backoff_quiet(accf_bar(job), (job >> 1) & 1u);
            tc_fence_after();
            if (kTrace && tr) p.trace[((ti * p.njobs) + jb_i) * 4 + 2] = clock64();
            epib_carry(c, jb, pf);
          } else {
            EmbPre pf;
            emb_prefetch(c, jb, pf);
            mbar_wait_backoff_quiet(accf_bar(job), (job >> 1) & 1u);
            tc_fence_after();
            if (kTrace && tr) p.trace[((ti * p.njobs) + jb_i) * 4 + 2] = clock64();
            epib_emb(c, jb, pf);
          }
          tc_fence_before();
          fence_proxy_async_smem();
          if (kTrace && tr) {
            p.trace[((ti * p.njobs) + jb_i) * 4 + 3] = clock64();
          }
          WMRL_JOB_DONE_B(job);
          ++job;
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == kMmaWarp) tmem_dealloc2(tmem_base, 512);
}

// --------------------------------------------------------------------------
// weight gradients: dW[n][map(m)] += sum over 64-row blocks of  A_blk[m][rows] . B_blk[n][rows]
//   A = in-feature field (actor block input: feat / y[l-1]), 128 in-features per CTA tile (TMEM lanes)
//   B = out-feature field (g_pre[l] / g_mu), up to 256 out-features per CTA tile (TMEM columns)
// Both fields are stash panels [feature/8][64 rows][8]: for a contraction over ROWS that is exactly the canonical
// no-swizzle MN-major operand layout (8 features contiguous, 8-row groups 128 B apart, feature chunks 1 KB apart).
// --------------------------------------------------------------------------
struct WgParams {
  const uint8_t* a_base; size_t a_blk;
  const uint8_t* b_base; size_t b_blk;
  int nblocks, m_cols, n_cols, n_valid;
  int seg0_len, seg0_valid, seg1_valid;   // in-feature panel column -> weight column (feat = [h | pad | s | pad])
  float* dW; int ldw;
  int mtiles, ntiles, splits;
  uint32_t lbo, sbo;
};
constexpr int kWgStages = 4, kWgThreads = 192;
constexpr uint32_t kWgA = 16384, kWgB = 32768, kWgStage = kWgA + kWgB;
constexpr uint32_t kWgSmem = kWgStages * kWgStage + 1024;

__device__ __forceinline__ uint32_t make_idesc_bf16_mn(uint32_t M, uint32_t N) {
  return make_idesc_bf16(M, N) | (1u << 15) | (1u << 16);   // A and B MN-major
}

__global__ void __launch_bounds__(kWgThreads, 1) tc_wgrad_kernel(const WgParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bars = smem_base + kWgStages * kWgStage;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 64u + 8u * s; };
  const uint32_t acc_bar = bars + 128u;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kWgStages * kWgStage + 136);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, split = blockIdx.y;
  const int mt = tile / p.ntiles, nt = tile % p.ntiles;
  const int m0 = mt * 128, n0 = nt * 256;
  const int mc = min(128, p.m_cols - m0), nc = min(256, p.n_cols - n0);
  const uint32_t a_bytes = (uint32_t)(mc / 8) * kStashChunk, b_bytes = (uint32_t)(nc / 8) * kStashChunk;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(acc_bar, 1);
    fence_barrier_init();
  }
  if (warp == 5) { tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nmine = (p.nblocks > split) ? (p.nblocks - split + p.splits - 1) / p.splits : 0;
  if (warp == 4) {
    if (lane == 0) {
      uint32_t slot = 0, phase = 0;
      for (int i = 0; i < nmine; ++i) {
        const size_t blk = (size_t)split + (size_t)i * p.splits;
        mbar_wait(empty_bar(slot), phase ^ 1u);
        mbar_arrive_expect_tx(full_bar(slot), a_bytes + b_bytes);
        bulk_g2s(smem_base + slot * kWgStage, p.a_base + blk * p.a_blk + (size_t)(m0 / 8) * kStashChunk, a_bytes, full_bar(slot));
        bulk_g2s(smem_base + slot * kWgStage + kWgA, p.b_base + blk * p.b_blk + (size_t)(n0 / 8) * kStashChunk, b_bytes,
                 full_bar(slot));
        if (++slot == kWgStages) { slot = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16_mn(128, (uint32_t)nc);
      uint32_t slot = 0, phase = 0;
      for (int i = 0; i < nmine; ++i) {
        mbar_wait(full_bar(slot), phase);
        tc_fence_after();
        const uint32_t a0 = smem_base + slot * kWgStage, b0 = a0 + kWgA;
#pragma unroll
        for (int k = 0; k < kRowsB / 16; ++k) {   // 16 rows (K) per MMA: two 8-row groups = 256 B further on
          const uint64_t da = make_smem_desc(a0 + k * 256, p.lbo, p.sbo);
          const uint64_t db = make_smem_desc(b0 + k * 256, p.lbo, p.sbo);
          mma_bf16_ss(tmem_base, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        tc_commit(empty_bar(slot));
        if (++slot == kWgStages) { slot = 0; phase ^= 1u; }
      }
      tc_commit(acc_bar);
    }
  } else {
    // epilogue: thread = TMEM lane = in-feature m; consecutive lanes -> consecutive dW columns (coalesced atomics)
    mbar_wait_backoff(acc_bar, 0);
    tc_fence_after();
    if (nmine > 0) {
      const int m = m0 + warp * 32 + lane;
      int wc = -1;
      if (m < p.m_cols) {
        if (m < p.seg0_len) wc = (m < p.seg0_valid) ? m : -1;
        else wc = (m - p.seg0_len < p.seg1_valid) ? p.seg0_valid + (m - p.seg0_len) : -1;
      }
      const uint32_t tl = tmem_base + ((uint32_t)(warp * 32) << 16);
      for (int j0 = 0; j0 < nc; j0 += 16) {
        float v[16];
        tmem_ld16(tl + j0, v);
        tmem_ld_wait();
        if (wc >= 0) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int n = n0 + j0 + j;
            if (n < p.n_valid) atomicAdd(p.dW + (size_t)n * p.ldw + wc, v[j]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, 256);
}

// --------------------------------------------------------------------------
// column sums of stash fields: out[c] += sum over blocks and rows of X[c] (* Y[c])
// --------------------------------------------------------------------------
struct CsParams {
  const uint8_t* x_base; size_t x_blk;
  const uint8_t* y_base; size_t y_blk;   // null: plain column sum
  // fused form (one pass over the three panels of an actor block instead of three passes reading four):
  // out += sum X (ln_b), out2 += sum X Y (ln_g), out3 += sum Z (lin_b); X = g_u, Y = x-hat, Z = g_pre
  const uint8_t* z_base; size_t z_blk;
  int nblocks, valid;
  float* out; float* out2; float* out3;
};
constexpr int kCsWarps = 8;
__device__ __forceinline__ void cs_reduce_store(float (&acc)[8], float* out, int chunk, int valid, float (*part)[8]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], s);
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < 8; ++i) part[warp][i] = acc[i];
  __syncthreads();
  if (threadIdx.x < 8) {
    float s = 0.f;
    for (int w2 = 0; w2 < kCsWarps; ++w2) s += part[w2][threadIdx.x];
    const int col = chunk * 8 + threadIdx.x;
    if (col < valid) atomicAdd(out + col, s);
  }
}
__global__ void __launch_bounds__(kCsWarps * 32) colsum_kernel(const CsParams p) {
  const int chunk = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[8], acc2[8], acc3[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { acc[i] = 0.f; acc2[i] = 0.f; acc3[i] = 0.f; }
  const size_t off = (size_t)chunk * kStashChunk + (size_t)lane * 16;
  const bool fused = p.z_base != nullptr;
  for (size_t blk = (size_t)blockIdx.y * kCsWarps + warp; blk < (size_t)p.nblocks; blk += (size_t)gridDim.y * kCsWarps) {
    float x0[8], x1[8];
    ld_chunk_g(p.x_base + blk * p.x_blk + off, x0);
    ld_chunk_g(p.x_base + blk * p.x_blk + off + 512, x1);
    if (p.y_base) {
      float y0[8], y1[8];
      ld_chunk_g(p.y_base + blk * p.y_blk + off, y0);
      ld_chunk_g(p.y_base + blk * p.y_blk + off + 512, y1);
      if (fused) {
        float z0[8], z1[8];
        ld_chunk_g(p.z_base + blk * p.z_blk + off, z0);
        ld_chunk_g(p.z_base + blk * p.z_blk + off + 512, z1);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc[i] += x0[i] + x1[i];
          acc2[i] += x0[i] * y0[i] + x1[i] * y1[i];
          acc3[i] += z0[i] + z1[i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += x0[i] * y0[i] + x1[i] * y1[i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += x0[i] + x1[i];
    }
  }
  __shared__ float part[kCsWarps][8];
  cs_reduce_store(acc, p.out, chunk, p.valid, part);
  if (fused) {
    cs_reduce_store(acc2, p.out2, chunk, p.valid, part);
    cs_reduce_store(acc3, p.out3, chunk, p.valid, part);
  }
}

// --------------------------------------------------------------------------
// host entry
// --------------------------------------------------------------------------
struct ConstStateB {
  const void* src = nullptr;
  unsigned long long gen = 0;
  cudaEvent_t last_launch = nullptr;
  cudaStream_t last_stream = nullptr;
};
static ConstStateB g_const_b[64];

static int wgrad_mode_swap() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("WMRL_WGRAD_SWAP"); v = (e && e[0] == '1') ? 1 : 0; }
  return v;
}

int tc_imagine_bwd(const Dims& d, const void* packed, const float* noise, const float* deter_seq, const float* dist_b,
                   const float* actions, const void* saved, const float* g_deter_seq, const float* g_stoch_seq,
                   const float* g_dist_a, const float* g_dist_b, const wmrl_actor_grads* gaw, float* g_deter0,
                   float* g_stoch0, void* ws, cudaStream_t st) {
  const size_t img = tc_bwd_image_offset(d);
  WMRL_REQUIRE(img != 0, "tc_imagine_bwd: shape not supported by the bf16 BPTT kernel");
  Plan pl = tc_make_bwd_plan(d, nullptr, nullptr);
  WMRL_REQUIRE(pl.ok, "tc_imagine_bwd: plan");
  const uint8_t* base = (const uint8_t*)packed + img;
  BwdParams p{};
  p.stage_data = base + pl.off_data;
  p.stash_f = (const uint8_t*)saved;
  const size_t nblk = n_blocks(d);
  p.stash_g = (uint8_t*)ws;
  p.gh_scratch = (float*)((uint8_t*)ws + (nblk * pl.sl.blk_g + 255) / 256 * 256);
  p.deter_seq = deter_seq; p.std_seq = dist_b; p.actions = actions; p.noise = noise;
  p.g_deter = g_deter_seq; p.g_stoch = g_stoch_seq; p.g_mean = g_dist_a; p.g_std = g_dist_b;
  p.g_deter0 = g_deter0; p.g_stoch0 = g_stoch0;
  p.nstages = (int)pl.stages.size(); p.njobs = (int)pl.jobs.size();
  p.B = d.B; p.H = d.T; p.D = d.D; p.S = d.S; p.A = d.A; p.Dp = pl.Dp; p.Sp = pl.Sp; p.Ha = d.Ha; p.Hdp = pl.Hdp; p.La = d.La;
  p.inv_action_scale = 1.f / d.action_scale;
  p.off_go = pl.boff[0]; p.off_gmu = pl.boff[1]; p.off_gx = pl.boff[2]; p.off_big = pl.boff[3]; p.off_carry = pl.boff[4];
  p.off_red = pl.boff[5]; p.bound_off = pl.boff[6];
  p.ring_off = pl.ring_off; p.bar_off = pl.bar_off; p.slot_bytes = (uint32_t)(pl.stage_bytes / 2);
  p.nslots = pl.nslots; p.ntiles = (d.B + kTileM - 1) / kTileM;
  p.sl = pl.sl;
  // timeline (tools/trace_bptt.py): the last 64 KB of the gradient-stash workspace are not used by the kernels
  p.trace = (d.flags & WMRL_FLAG_TRACE) ? (long long*)(p.gh_scratch + (size_t)p.ntiles * kTileM * pl.Dp) : nullptr;
  int dev = 0;
  WMRL_CUDA_OK(cudaGetDevice(&dev));
  WMRL_REQUIRE(dev >= 0 && dev < 64, "device ordinal out of range");
  static int s_sms[64];
  static unsigned s_smem[64];
  static bool s_wg[64];
  {
    std::lock_guard<std::mutex> lk(g_tc_mutex);
    if (s_sms[dev] == 0) WMRL_CUDA_OK(cudaDeviceGetAttribute(&s_sms[dev], cudaDevAttrMultiProcessorCount, dev));
    if (s_smem[dev] < pl.smem_bytes) {
      WMRL_CUDA_OK(cudaFuncSetAttribute(tc_bptt_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes));
      WMRL_CUDA_OK(cudaFuncSetAttribute(tc_bptt_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes));
      s_smem[dev] = pl.smem_bytes;
    }
    if (!s_wg[dev]) {
      WMRL_CUDA_OK(cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWgSmem));
      s_wg[dev] = true;
    }
    ConstStateB& cs = g_const_b[dev];
    if (cs.src != packed || cs.gen != tc_pack_generation(packed)) {
      if (cs.last_launch && cs.last_stream != st) WMRL_CUDA_OK(cudaStreamWaitEvent(st, cs.last_launch, 0));
      WMRL_CUDA_OK(cudaMemcpyToSymbolAsync(cb_stages, base + pl.off_stage_tab, pl.stages.size() * sizeof(StageDesc), 0,
                                           cudaMemcpyDeviceToDevice, st));
      WMRL_CUDA_OK(cudaMemcpyToSymbolAsync(cb_jobs, base + pl.off_job_tab, pl.jobs.size() * sizeof(JobDesc), 0,
                                           cudaMemcpyDeviceToDevice, st));
      WMRL_CUDA_OK(cudaMemcpyToSymbolAsync(cb_eparams, base + pl.off_eparams, pl.eparams_floats * sizeof(float), 0,
                                           cudaMemcpyDeviceToDevice, st));
      cs.src = packed;
      cs.gen = tc_pack_generation(packed);
    }
    const int pairs = std::min(p.ntiles, s_sms[dev] / 2);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(kThreadsB);
    cfg.dynamicSmemBytes = pl.smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (p.trace) WMRL_CUDA_OK(cudaLaunchKernelEx(&cfg, tc_bptt_kernel<true>, p));
    else WMRL_CUDA_OK(cudaLaunchKernelEx(&cfg, tc_bptt_kernel<false>, p));
    WMRL_CUDA_OK(cudaGetLastError());
    if (!cs.last_launch) WMRL_CUDA_OK(cudaEventCreateWithFlags(&cs.last_launch, cudaEventDisableTiming));
    WMRL_CUDA_OK(cudaEventRecord(cs.last_launch, st));
    cs.last_stream = st;
  }
  // ---- weight gradients over the stashes
  const int sms = s_sms[dev];
  const uint8_t* sf = (const uint8_t*)saved;
  const uint8_t* sg = (const uint8_t*)ws;
  const StashLayout& sl = pl.sl;
  const bool swap = wgrad_mode_swap() != 0;
  auto wgrad = [&](const uint8_t* a, int m_cols, int seg0_len, int seg0_valid, int seg1_valid, const uint8_t* b, int n_cols,
                   int n_valid, float* dW, int ldw) -> int {
    if (!dW) return 0;
    WgParams q{};
    q.a_base = a; q.a_blk = sl.blk_f; q.b_base = b; q.b_blk = sl.blk_g;
    q.nblocks = (int)nblk; q.m_cols = m_cols; q.n_cols = n_cols; q.n_valid = n_valid;
    q.seg0_len = seg0_len; q.seg0_valid = seg0_valid; q.seg1_valid = seg1_valid;
    q.dW = dW; q.ldw = ldw;
    q.mtiles = (m_cols + 127) / 128; q.ntiles = (n_cols + 255) / 256;
    const int tiles = q.mtiles * q.ntiles;
    q.splits = std::max(1, std::min((int)nblk, sms / tiles));
    q.lbo = swap ? 1024u : 128u;    // MN-major, no swizzle: 8-row (K) groups 128 B apart ...
    q.sbo = swap ? 128u : 1024u;    // ... 8-feature (MN) chunks 1 KB apart
    tc_wgrad_kernel<<<dim3(tiles, q.splits), kWgThreads, kWgSmem, st>>>(q);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
  };
  auto colsum = [&](const uint8_t* x, size_t xb, const uint8_t* y, size_t yb, int cols, int valid, float* out) -> int {
    if (!out) return 0;
    CsParams q{};
    q.x_base = x; q.x_blk = xb; q.y_base = y; q.y_blk = yb; q.nblocks = (int)nblk; q.valid = valid; q.out = out;
    const int chunks = cols / 8;
    const int ysplit = std::max(1, std::min(((int)nblk + kCsWarps - 1) / kCsWarps, (8 * sms + chunks - 1) / chunks));
    colsum_kernel<<<dim3(chunks, ysplit), kCsWarps * 32, 0, st>>>(q);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
  };
  for (int l = 0; l < d.La; ++l) {
    int rc;
    if (l == 0) rc = wgrad(sf + sl.feat, pl.Dp + pl.Sp, pl.Dp, d.D, d.S, sg + sl.gp[0], d.Ha, d.Ha, gaw->lin_w[0], d.F);
    else rc = wgrad(sf + sl.y[l - 1], d.Ha, d.Ha, d.Ha, 0, sg + sl.gp[l], d.Ha, d.Ha, gaw->lin_w[l], d.Ha);
    WMRL_REQUIRE(rc == 0, "tc_imagine_bwd: weight-gradient launch failed");
    if (gaw->lin_b[l] && gaw->ln_b[l] && gaw->ln_g[l]) {   // one pass over g_u, x-hat, g_pre
      CsParams q{};
      q.x_base = sg + sl.gu[l]; q.x_blk = sl.blk_g; q.y_base = sf + sl.xh[l]; q.y_blk = sl.blk_f;
      q.z_base = sg + sl.gp[l]; q.z_blk = sl.blk_g; q.nblocks = (int)nblk; q.valid = d.Ha;
      q.out = gaw->ln_b[l]; q.out2 = gaw->ln_g[l]; q.out3 = gaw->lin_b[l];
      const int chunks = d.Ha / 8;
      const int ysplit = std::max(1, std::min(((int)nblk + kCsWarps - 1) / kCsWarps, (8 * sms + chunks - 1) / chunks));
      colsum_kernel<<<dim3(chunks, ysplit), kCsWarps * 32, 0, st>>>(q);
      WMRL_REQUIRE(cudaGetLastError() == cudaSuccess, "colsum launch failed");
    } else {
      WMRL_REQUIRE(colsum(sg + sl.gp[l], sl.blk_g, nullptr, 0, d.Ha, d.Ha, gaw->lin_b[l]) == 0, "colsum launch failed");
      WMRL_REQUIRE(colsum(sg + sl.gu[l], sl.blk_g, nullptr, 0, d.Ha, d.Ha, gaw->ln_b[l]) == 0, "colsum launch failed");
      WMRL_REQUIRE(colsum(sg + sl.gu[l], sl.blk_g, sf + sl.xh[l], sl.blk_f, d.Ha, d.Ha, gaw->ln_g[l]) == 0, "colsum launch failed");
    }
  }
  WMRL_REQUIRE(wgrad(sf + sl.y[d.La - 1], d.Ha, d.Ha, d.Ha, 0, sg + sl.gmu, 16, d.A, gaw->mu_w, d.Ha) == 0,
               "tc_imagine_bwd: weight-gradient launch failed");
  WMRL_REQUIRE(colsum(sg + sl.gmu, sl.blk_g, nullptr, 0, 16, d.A, gaw->mu_b) == 0, "colsum launch failed");
  return 0;
}

}  // namespace wmrl
