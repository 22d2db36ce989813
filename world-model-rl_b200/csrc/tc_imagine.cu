// bf16 tensor-core path of RSSMBase.imagine_rollout (rssm.py:123-140), continuous latent:
// ONE persistent kernel walks all H steps of a batch tile.  Four kernel variants share the plan
// builder, weight packer and epilogue math (WMRL_TC_MODE):
//   3 (default) tc_imagine_fold_kernel  CTA pairs, tcgen05.mma.cta_group::2 M=128 (64 rows per CTA),
//                                       accumulators folded over the TMEM lane halves
//   1           tc_imagine_kernel<1>    single CTAs, 128-row tiles
//   2           tc_imagine_kernel<2>    CTA pairs at M=256
//   4           tc_imagine_pp_kernel    folded pairs ping-ponging two tiles
//
//   * the tile's latent state and every intermediate activation stay in shared memory as bf16 MMA
//     A-operand panels ([k/8][rows][8] = tcgen05 K-major no-swizzle core-matrix layout); the fp32
//     master copy of h lives in the output tensor itself (one row slice per thread);
//   * the 9 dependent GEMMs of a step (actor x3, mu, embed, GRU in two column chunks, prior x2) are
//     tcgen05.mma tiles with fp32 accumulators in TMEM; weights are pre-packed once into the exact
//     per-stage operand panels and streamed L2 -> smem by the TMA engine (cp.async.bulk + mbarrier
//     tx-count) through a ring, in the fixed order the MMA lane consumes them;
//   * warp roles: 16 epilogue warps (8 threads per batch row in mode 3), then - at the highest warp
//     ids, which the issue arbiter favours - the TMA producer, the MMA issuer (one lane) and, in
//     mode 3, a helper lane that turns the ring's `full` mbarriers into a shared-memory counter so
//     that the MMA lane never pays for a try_wait;
//   * bias, LayerNorm+ELU (packed f32x2 arithmetic), GRU gates, tanh squash, softplus and the
//     reparameterised sample are fused into the TMEM->register epilogues, which write the next GEMM's
//     A panel (bf16, smem) and the fp32 sequence outputs;
//   * mode 3 hoists MMAs that do not depend on the epilogue in flight: the h part of the next step's
//     first actor layer runs under the prior-head epilogues, GRU chunk 0's h side under the last actor
//     layer's epilogue (ST_PRE stages / static TMEM halves, see make_plan_mode).
//
// The work is described by two host-built tables (same for every step): STAGES (what the producer
// loads and the MMA lane issues) and JOBS (what the epilogue warps do once a job's accumulator is
// complete).
#include <stdlib.h>

#include <algorithm>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "tc_plan.cuh"
#include "tc_ptx.cuh"

namespace wmrl {

// Kernel variants of the tcgen05 path (env WMRL_TC_MODE, read once):
//   1 = single CTAs, 128-row tiles (cta_group::1, M=128)
//   2 = CTA pairs, 2 x 128 rows (cta_group::2, M=256; every weight stage split across the pair)
//   3 = CTA pairs, 2 x 64 rows ("folded": cta_group::2, M=128; each CTA's 64 rows use all 128 TMEM lanes,
//       lanes 0-63 hold accumulator columns [0,N/2), lanes 64-127 columns [N/2,N); half-size panels leave room
//       for a 16-slot weight ring)
//   4 = mode 3 + two-tile ping-pong: every CTA holds TWO 64-row tiles (2 x 96 KB of panels, one 256-column
//       TMEM half each) and alternates tiles job by job, so one tile's epilogue overlaps the other's MMAs
static int tc_mode() {
  static int mode = 0;
  if (mode == 0) {
    const char* e = getenv("WMRL_TC_MODE");
    mode = (e && e[0] >= '1' && e[0] <= '4') ? (e[0] - '0') : kDefaultTcMode;
  }
  return mode;
}

// With 227 KB of shared memory carved out the L1D is ~1 KB, so any global load
// in the inner loops is a full L2 round trip.  The per-step tables and the
// epilogue parameters (biases, LayerNorm affine, bound) are therefore served
// from the constant cache: every access is warp-uniform.
constexpr int kMaxStages = 512, kMaxJobs = 32, kMaxEparams = 12032;
constexpr int kMaxStagesSmem = 144;      // stage descriptors mirrored in shared memory (16 B each)
constexpr int kMaxSlots = 16, kBarBytes = 512;   // barrier map: full[16] | empty[16] | peer_full[16] | acc | tmem ptr
__constant__ uint4 c_stages[kMaxStages];
__constant__ JobDesc c_jobs[kMaxJobs];
__constant__ __align__(16) float c_eparams[kMaxEparams];


// Build the per-step tables.  Weight pointers may be null (size / launch queries).
static Plan make_plan_mode(const Dims& d, const wmrl_rssm_params* w, const wmrl_actor_params* aw, int mode) {
  Plan pl;
  pl.cta = mode >= 2 ? 2 : 1;
  pl.fold = mode >= 3;
  pl.pp = mode == 4;
  pl.ps = pl.fold ? 1024u : 2048u;
  // folded pairs have ~128 KB of ring per CTA: 64 KB stages (32 KB per CTA, 4 slots) cut the number of
  // stages - and with it the per-stage handshakes of the pair protocol - by ~3.5x
  pl.stage_bytes = pl.pp ? 32768 : (pl.fold ? 65536 : kSlotBytes);
  if (d.variant != WMRL_CONTINUOUS) return pl;
  if (d.Ha % 32 != 0 || d.Ha > 512 || d.A > 16 || d.S > 8 * kQ * kP2Groups || d.La < 1) return pl;
  if (pl.fold && (d.Ha % 64 != 0 || d.A > 8)) return pl;
  pl.Dp = ceil16(d.D); pl.Sp = ceil16(d.S); pl.Ap = 16; pl.Ha = d.Ha; pl.Hdp = ceil16(d.Hd);
  if (pl.Hdp > 512 || pl.Dp > 4096) return pl;
  pl.sl = make_stash_layout(pl.Dp, pl.Sp, d.Ha, pl.Hdp, d.La);
  const uint32_t PS = pl.ps;
  pl.h2c = std::max(pl.Dp, pl.Hdp);
  pl.XW = std::max(d.Ha, pl.h2c + pl.Dp);
  pl.off_h = 0;
  pl.off_s = pl.off_h + pl.Dp / 8 * PS;
  pl.off_a = pl.off_s + pl.Sp / 8 * PS;
  pl.off_x = pl.off_a + pl.Ap / 8 * PS;
  pl.off_h2 = pl.off_x + pl.h2c / 8 * PS;
  pl.panels_bytes = pl.off_x + pl.XW / 8 * PS;
  const int ntl = pl.pp ? 2 : 1;   // tiles resident per CTA
  pl.ring_off = ntl * pl.panels_bytes;
  const int avail = kMaxSmem - ntl * (int)pl.panels_bytes - kBarBytes - kMaxStagesSmem * 16;
  const int slot_bytes = pl.stage_bytes / pl.cta;   // a CTA of a pair stores half of every stage
  pl.nslots = avail / slot_bytes;
  if (pl.nslots < 2) return pl;
  if (pl.nslots > kMaxSlots) pl.nslots = kMaxSlots;
  pl.bar_off = pl.ring_off + pl.nslots * slot_bytes;
  pl.smem_bytes = pl.bar_off + kBarBytes + kMaxStagesSmem * 16;   // barriers + the stage table copy
  if (pl.panels_bytes / 16 > 65535) return pl;
  // TMEM columns taken by an accumulator group of N columns
  auto tc = [&](int n) { return pl.fold ? n / 2 : n; };

  const float* nul = nullptr;
  const int Ke = d.S_in + d.A;
  // Folded pairs: the h part of the first actor layer (13 of its 15 K steps at C4) does not depend on the
  // prior head, so its MMAs for step t+1 are issued right after the prior2 MMAs of step t and run under
  // prior2's epilogue; only the s part stays on the critical path.  prior2 accumulates at TMEM column 256.
  const bool early_h = pl.fold && !pl.pp;
  // ---- actor blocks
  for (int l = 0; l < d.La; ++l) {
    JobDesc jb{};
    jb.type = JOB_ACTOR_LN; jb.n = (uint16_t)d.Ha; jb.tmem_col = 0; jb.dst_off = pl.off_x;
    jb.st0 = pl.sl.y[l]; jb.st1 = pl.sl.xh[l]; jb.layer = (uint32_t)l;
    const float* lw = aw ? aw->lin_w[l] : nul;
    const uint32_t pb = add_vec(pl, aw ? aw->lin_b[l] : nul, nul, 0, d.Ha, d.Ha);
    // LayerNorm affine pre-scaled by log2(e): the epilogue evaluates ELU in the exp2 domain
    add_vec(pl, aw ? aw->ln_g[l] : nul, nul, 0, d.Ha, d.Ha, 1.4426950408889634f);
    add_vec(pl, aw ? aw->ln_b[l] : nul, nul, 0, d.Ha, d.Ha, 1.4426950408889634f);
    jb.p_off = pb;
    // aux 1: the Linear bias is folded into the GEMM (bias row x the constant-one column 15 of P_a).  Only layer 0:
    // its [P_s | P_a] panels are contiguous, so the bias rides in the s-part stage; for the other layers it would
    // cost an extra weight stage per N group (measured: slower than the adds it removes).
    jb.aux = (early_h && l == 0) ? 1u : 0u;
    pl.jobs.push_back(jb);
    const int nchunks = (d.Ha + 255) / 256;
    for (int nc = 0; nc < nchunks; ++nc) {
      const int rows = std::min(256, d.Ha - nc * 256);
      RowSpec rs{nc * 256, rows, rows, 0, 0, 0};
      std::vector<Seg> segs;
      if (l == 0) {
        // A = [P_h | P_s]: contiguous panels, source columns [0,D) and [D,D+S)
        if (!early_h) {
          segs.push_back({pl.off_h, pl.Dp, lw, d.F, 0, d.D});
          segs.push_back({pl.off_s, pl.Sp, lw, d.F, d.D, d.S_in});
        } else {   // [P_s | P_a] as one K range: s columns, zeros, and the bias row at the ones column (P_a column 15)
          segs.push_back({pl.off_s, pl.Sp + pl.Ap, lw, d.F, d.D, d.S_in, aw ? aw->lin_b[l] : nul, pl.Sp + 15});
        }
      } else {
        segs.push_back({pl.off_x, d.Ha, lw, d.Ha, 0, d.Ha});
      }
      add_acc_group(pl, rs, segs, nc * tc(256), false, nc == 0, 1, nc == nchunks - 1, l == 0 && early_h);
    }
  }
  // ---- GRU chunk geometry (see the GRU block below)
  const int gru_groups = pl.Dp / 16;
  const int gru_nchunks = (pl.Dp + kGruChunk - 1) / kGruChunk;
  auto gru_geom = [&](int c, int& col0, int& cw) {
    col0 = 0;
    for (int i = 0; i <= c; ++i) {
      cw = (gru_groups / gru_nchunks + (i < gru_groups % gru_nchunks ? 1 : 0)) * 16;
      // early_h, two chunks: chunk 0 as wide as its TMEM half allows (128 state columns = 256 TMEM columns) - more of
      // the h-side work is hoisted and the second chunk's epilogue, which nothing overlaps, gets shorter
      if (early_h && gru_nchunks == 2) cw = (i == 0) ? std::min(kGruChunk, pl.Dp) : pl.Dp - std::min(kGruChunk, pl.Dp);
      if (i < c) col0 += cw;
    }
  };
  // static TMEM halves in early_h mode: even chunks at column 256, odd chunks at 0
  auto gru_base = [&](int c) { return early_h ? ((c % 2 == 0) ? 256 : 0) : 0; };
  auto gru_rz = [&](int col0, int cw) {
    const int hw = cw / 2;
    RowSpec rz;
    rz.add(col0, d.D - col0, hw);
    rz.add(d.D + col0, d.D - col0, hw);
    rz.add(col0 + hw, d.D - col0 - hw, hw);
    rz.add(d.D + col0 + hw, d.D - col0 - hw, hw);
    return rz;
  };
  if (early_h) {
    // h side of GRU chunk 0 (W_hh h for r|z and for h_n: half of the chunk's MMAs) only needs h(t): issued right
    // after the last actor layer's MMAs, it runs under that layer's LayerNorm epilogue.
    int col0, cw;
    gru_geom(0, col0, cw);
    const int valid = std::max(0, std::min(cw, d.D - col0));
    const float* whh = w ? w->w_hh : nul;
    add_acc_group(pl, gru_rz(col0, cw), {{pl.off_h, pl.Dp, whh, d.D, 0, d.D}}, gru_base(0), false, false, 0, false);
    RowSpec rn(2 * d.D + col0, valid, cw);
    add_acc_group(pl, rn, {{pl.off_h, pl.Dp, whh, d.D, 0, d.D}}, gru_base(0) + cw + cw / 2, false, false, 0, false);
  }
  // ---- mu head
  {
    JobDesc jb{};
    jb.type = JOB_MU; jb.n = 16; jb.tmem_col = 0; jb.dst_off = pl.off_a; jb.aux = early_h ? 1u : 0u;
    jb.st0 = jb.st1 = kNoStash;
    jb.p_off = add_vec(pl, aw ? aw->mu_b : nul, nul, 0, d.A, 16);
    add_vec(pl, aw ? aw->bound : nul, nul, 0, d.A, 16);
    pl.jobs.push_back(jb);
    RowSpec rs{0, d.A, 16, 0, 0, 0};
    add_acc_group(pl, rs, {{pl.off_x, d.Ha, aw ? aw->mu_w : nul, d.Ha, 0, d.Ha}}, 0, false, true, 1, true);
  }
  // ---- state-action embed: A = [P_s | P_a]
  {
    JobDesc jb{};
    jb.type = JOB_DENSE_ELU; jb.n = (uint16_t)pl.Dp; jb.tmem_col = 0; jb.dst_off = pl.off_x;
    jb.st0 = pl.sl.xe; jb.st1 = kNoStash;
    jb.p_off = add_vec(pl, w ? w->embed_b : nul, nul, 0, d.D, pl.Dp);
    pl.jobs.push_back(jb);
    const float* ew = w ? w->embed_w : nul;
    // N = Dp may exceed 256 -> chunks of 256 accumulator columns
    const int nchunks = (pl.Dp + 255) / 256;
    if (pl.Dp > 512) return pl;
    for (int nc = 0; nc < nchunks; ++nc) {
      const int padr = std::min(256, pl.Dp - nc * 256);
      const int valid = std::max(0, std::min(padr, d.D - nc * 256));
      RowSpec rs{nc * 256, valid, padr, 0, 0, 0};
      add_acc_group(pl, rs, {{pl.off_s, pl.Sp, ew, Ke, 0, d.S_in}, {pl.off_a, pl.Ap, ew, Ke, d.S_in, d.A}},
                    nc * tc(256), false, nc == 0, 1, nc == nchunks - 1);
    }
  }
  // ---- GRU, chunks of 64 state columns: TMEM [r | z | i_n | h_n], double buffered
  {
    const uint32_t pb = add_vec(pl, w ? w->b_ih : nul, w ? w->b_hh : nul, 0, d.D, pl.Dp);       // b_r
    add_vec(pl, w ? w->b_ih : nul, w ? w->b_hh : nul, d.D, d.D, pl.Dp);                         // b_z
    add_vec(pl, w ? w->b_ih : nul, nul, 2 * d.D, d.D, pl.Dp);                                    // b_in
    add_vec(pl, w ? w->b_hh : nul, nul, 2 * d.D, d.D, pl.Dp);                                    // b_hn
    // Chunks of up to 128 state columns (4 accumulators per column = up to 512 TMEM columns), balanced
    // in units of 16.  Wide chunks halve the number of MMAs (each costs >= ~61 cycles whatever its N)
    // and of weight stages; the chunks run serially (MMA -> epilogue -> next chunk).
    const int nchunks = gru_nchunks;
    for (int c = 0; c < nchunks; ++c) {
      int col0, cw;
      gru_geom(c, col0, cw);
      const int valid = std::max(0, std::min(cw, d.D - col0));
      JobDesc jb{};
      jb.type = JOB_GRU; jb.dbuf = (pl.fold && !pl.pp && !early_h) ? 1 : 0; jb.n = (uint16_t)cw;
      jb.tmem_col = (uint16_t)gru_base(c); jb.col0 = (uint16_t)col0;
      jb.dst_off = pl.off_h2; jb.p_off = pb;
      jb.st0 = pl.sl.gr; jb.st1 = kNoStash;
      pl.jobs.push_back(jb);
      const float* wih = w ? w->w_ih : nul;
      const float* whh = w ? w->w_hh : nul;
      if (pl.fold) {
        // folded: a lane half owns state columns [col0 + hf*cw/2, +cw/2) and needs r, z, i_n, h_n of exactly
        // those -> B rows of the r|z group are ordered [r lo, z lo, r hi, z hi]; TMEM (per lane half):
        // r @0, z @cw/2, i_n @cw, h_n @cw+cw/2 - 2*cw columns, double buffered in the two 256-column halves
        const int hw = cw / 2;
        const RowSpec rz = gru_rz(col0, cw);
        // ping-pong mode gives each tile one 256-column TMEM half: chunks are single-buffered there; early_h mode
        // uses static halves (even chunks at column 256, odd at 0) so that chunk 0's h side can be issued early
        const bool two_halves = !pl.pp;
        const bool db = two_halves && !early_h;
        const int base = gru_base(c);
        const int dep = two_halves ? ((c < 2) ? (c + 1) : 2) : 1;   // chunk c waits for chunk c-2's epilogue (same half)
        RowSpec rn(2 * d.D + col0, valid, cw);
        if (early_h && c == 0) {
          // the h side was issued after the actor layers; here only the x side, on top of it
          add_acc_group(pl, rz, {{pl.off_x, pl.Dp, wih, d.D, 0, d.D}}, base, false, true, dep, false, true);
          add_acc_group(pl, rn, {{pl.off_x, pl.Dp, wih, d.D, 0, d.D}}, base + cw, false, false, 0, true);
        } else {
          add_acc_group(pl, rz, {{pl.off_x, pl.Dp, wih, d.D, 0, d.D}, {pl.off_h, pl.Dp, whh, d.D, 0, d.D}}, base, db,
                        true, dep, false);
          add_acc_group(pl, rn, {{pl.off_x, pl.Dp, wih, d.D, 0, d.D}}, base + cw, db, false, 0, false);
          add_acc_group(pl, rn, {{pl.off_h, pl.Dp, whh, d.D, 0, d.D}}, base + cw + hw, db, false, 0, true);
        }
      } else {
        RowSpec rz{col0, valid, cw, d.D + col0, valid, cw};
        add_acc_group(pl, rz, {{pl.off_x, pl.Dp, wih, d.D, 0, d.D}, {pl.off_h, pl.Dp, whh, d.D, 0, d.D}}, 0, false,
                      true, 1, false);
        RowSpec rn{2 * d.D + col0, valid, cw, 0, 0, 0};
        add_acc_group(pl, rn, {{pl.off_x, pl.Dp, wih, d.D, 0, d.D}}, 2 * cw, false, false, 0, false);
        add_acc_group(pl, rn, {{pl.off_h, pl.Dp, whh, d.D, 0, d.D}}, 3 * cw, false, false, 0, true);
      }
    }
  }
  // ---- prior head
  {
    JobDesc jb{};
    jb.type = JOB_DENSE_ELU; jb.n = (uint16_t)pl.Hdp; jb.tmem_col = early_h ? 256 : 0; jb.dst_off = pl.off_x;
    jb.st0 = pl.sl.hid; jb.st1 = kNoStash;
    jb.p_off = add_vec(pl, w ? w->prior0_b : nul, nul, 0, d.Hd, pl.Hdp);
    pl.jobs.push_back(jb);
    const int nchunks = (pl.Hdp + 255) / 256;
    for (int nc = 0; nc < nchunks; ++nc) {
      const int padr = std::min(256, pl.Hdp - nc * 256);
      const int valid = std::max(0, std::min(padr, d.Hd - nc * 256));
      RowSpec rs{nc * 256, valid, padr, 0, 0, 0};
      add_acc_group(pl, rs, {{pl.off_h2, pl.Dp, w ? w->prior0_w : nul, d.D, 0, d.D}}, (early_h ? 256 : 0) + nc * tc(256), false,
                    nc == 0, 1, nc == nchunks - 1);
    }
  }
  // next step's first actor layer, h part (ST_PRE).  It reads h' where the GRU epilogues put it (X[h2..]); P_h is being
  // rewritten by prior2's epilogue at that time.  The tile initialisation stores h0 there as well.  The first N chunk
  // runs under prior1's epilogue, the rest under prior2's.
  auto early_actor0_h = [&](int nc_begin, int nc_end) {
    const float* lw = aw ? aw->lin_w[0] : nul;
    for (int nc = nc_begin; nc < nc_end; ++nc) {
      const int rows = std::min(256, d.Ha - nc * 256);
      RowSpec rs{nc * 256, rows, rows, 0, 0, 0};
      add_acc_group(pl, rs, {{pl.off_h2, pl.Dp, lw, d.F, 0, d.D}}, nc * tc(256), false, false, 0, false, false, ST_PRE);
    }
  };
  const int a0_chunks = (d.Ha + 255) / 256;
  if (early_h) early_actor0_h(0, 1);
  {
    JobDesc jb{};
    jb.type = JOB_PRIOR2_CONT; jb.n = (uint16_t)(2 * pl.Sp); jb.tmem_col = early_h ? 256 : 0; jb.dst_off = pl.off_s;
    jb.st0 = jb.st1 = kNoStash;
    jb.p_off = add_vec(pl, w ? w->prior2_b : nul, nul, 0, d.S, pl.Sp);
    add_vec(pl, w ? w->prior2_b : nul, nul, d.S, d.S, pl.Sp);
    pl.jobs.push_back(jb);
    RowSpec rs;
    if (pl.fold) {   // per lane half: [mean | raw] of the latent columns it owns
      const int hw = pl.Sp / 2;
      rs.add(0, d.S, hw);
      rs.add(d.S, d.S, hw);
      rs.add(hw, d.S - hw, hw);
      rs.add(d.S + hw, d.S - hw, hw);
    } else {         // mean rows -> cols [0,Sp), raw rows -> [Sp,2Sp)
      rs.add(0, d.S, pl.Sp);
      rs.add(d.S, d.S, pl.Sp);
    }
    add_acc_group(pl, rs, {{pl.off_x, pl.Hdp, w ? w->prior2_w : nul, d.Hd, 0, d.Hd}}, early_h ? 256 : 0, false, true, 1,
                  true);
  }
  if (early_h) early_actor0_h(1, a0_chunks);
  for (const StageDesc& st : pl.stages)
    if ((int)st.bytes16 * 16 > pl.stage_bytes) return pl;
  if ((int)pl.stages.size() > kMaxStagesSmem || (int)pl.stages.size() > kMaxStages || (int)pl.jobs.size() > kMaxJobs ||
      (int)pl.eparams_floats > kMaxEparams)
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

// preferred variant first; shapes it does not cover fall back to the single-CTA kernel
static Plan make_plan(const Dims& d, const wmrl_rssm_params* w, const wmrl_actor_params* aw) {
  Plan pl = make_plan_mode(d, w, aw, tc_mode());
  if (!pl.ok && tc_mode() != 1) pl = make_plan_mode(d, w, aw, 1);
  return pl;
}

bool tc_supported(const Dims& d) { return make_plan(d, nullptr, nullptr).ok; }
size_t tc_packed_bytes(const Dims& d) {
  Plan pl = make_plan(d, nullptr, nullptr);
  if (!pl.ok) return 0;
  size_t n = pl.total_bytes;
  if (pl.fold && !pl.pp) {
    Plan bp = tc_make_bwd_plan(d, nullptr, nullptr);
    if (bp.ok) n += bp.total_bytes;
  }
  return n;
}
// byte offset of the BPTT image inside the packed buffer (0: this shape has no bf16 training path)
size_t tc_bwd_image_offset(const Dims& d) {
  Plan pl = make_plan(d, nullptr, nullptr);
  if (!pl.ok || !pl.fold || pl.pp) return 0;
  return tc_make_bwd_plan(d, nullptr, nullptr).ok ? pl.total_bytes : 0;
}
size_t tc_workspace_bytes(const Dims&) { return 1024 * sizeof(long long); }

// --------------------------------------------------------------------------
// weight packing kernels
// --------------------------------------------------------------------------
__global__ void pack_stage_kernel(const PackOp* __restrict__ ops, uint8_t* __restrict__ data) {
  const PackOp op = ops[blockIdx.x];
  int N = 0;
  for (int q = 0; q < op.nseg; ++q) N += op.pad[q];
  const int total = op.nk8 * N;
  const int halfN = N / op.split;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int kc = i / N, n = i % N;
    const int di = (n / halfN) * (op.nk8 * halfN) + kc * halfN + (n % halfN);
    int row = -1, m = n;
    for (int q = 0; q < op.nseg; ++q) {
      if (m < op.pad[q]) { if (m < op.valid[q]) row = op.row0[q] + m; break; }
      m -= op.pad[q];
    }
    __align__(16) __nv_bfloat16 out[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = kc * 8 + e;
      float v = 0.f;
      int ks = (k < op.kvalid) ? op.k0 + k : -1;     // source K index, -1: padding
      if (op.nks > 0) {
        ks = -1;
        int m2 = op.kbegin + k;
        for (int q = 0; q < op.nks; ++q) {
          if (m2 < op.kkpad[q]) { if (m2 < op.kkvalid[q]) ks = op.kk0[q] + m2; break; }
          m2 -= op.kkpad[q];
        }
      }
      if (row >= 0 && ks >= 0)
        v = op.trans ? op.src[(size_t)ks * op.ld + row] : op.src[(size_t)row * op.ld + ks];
      if (op.bias && row >= 0 && k == op.bias_k) v = op.bias[row];
      out[e] = __float2bfloat16_rn(v);
    }
    *reinterpret_cast<uint4*>(data + op.dst_off + (size_t)di * 16) = *reinterpret_cast<const uint4*>(out);
  }
}
__global__ void pack_vec_kernel(const VecOp* __restrict__ ops, float* __restrict__ dst) {
  const VecOp op = ops[blockIdx.x];
  for (int i = threadIdx.x; i < op.pad; i += blockDim.x) {
    float v = 0.f;
    if (i < op.valid) {
      v = op.src1[op.off + i];
      if (op.src2) v += op.src2[op.off + i];
    }
    dst[op.dst + i] = v * op.scale;
  }
}

// which packed image currently sits in constant memory (single host thread per device, see wmrl.h)
// Which packed image currently sits in constant memory, PER DEVICE, plus the event of the last launch that reads
// it: several RSSM / actor pairs (different images) may be driven from different streams or host threads.  The
// mutex serialises {constant upload, launch, event record}; an upload of a different image first makes its stream
// wait for the previous launch, so a running persistent kernel never has its tables overwritten.
std::mutex g_tc_mutex;
unsigned long long g_pack_counter = 0;
std::unordered_map<const void*, unsigned long long> g_pack_gen;
unsigned long long tc_pack_generation(const void* packed) {   // caller holds g_tc_mutex
  auto it = g_pack_gen.find(packed);
  return it == g_pack_gen.end() ? 0ull : it->second;
}
struct ConstState {
  const void* src = nullptr;
  unsigned long long gen = 0;
  cudaEvent_t last_launch = nullptr;
  cudaStream_t last_stream = nullptr;
};
static ConstState g_const[64];

// tables + packed stages + epilogue parameters of one plan -> its image at `base`
int tc_pack_plan(const Plan& pl, uint8_t* base, cudaStream_t st);

int tc_pack_weights(const Dims& d, const wmrl_rssm_params* w, const wmrl_actor_params* aw,
                    void* packed, cudaStream_t st) {
  Plan pl = make_plan(d, w, aw);
  WMRL_REQUIRE(pl.ok, "pack_weights: unsupported shape");
  {
    std::lock_guard<std::mutex> lk(g_tc_mutex);
    if (g_pack_gen.size() > 4096) g_pack_gen.clear();   // stale entries only cost a re-upload of the tables
    g_pack_gen[packed] = ++g_pack_counter;
  }
  if (int r = tc_pack_plan(pl, (uint8_t*)packed, st)) return r;
  // training: the BPTT kernel's image (transposed weight stages) follows the rollout kernel's
  if (pl.fold && !pl.pp) {
    Plan bp = tc_make_bwd_plan(d, w, aw);
    if (bp.ok) return tc_pack_plan(bp, (uint8_t*)packed + pl.total_bytes, st);
  }
  return 0;
}

int tc_pack_plan(const Plan& pl, uint8_t* base, cudaStream_t st) {
  // tables travel host -> device inside the packed image (synchronous for pageable memory)
  WMRL_CUDA_OK(cudaMemcpyAsync(base + pl.off_stage_tab, pl.stages.data(), pl.stages.size() * sizeof(StageDesc),
                               cudaMemcpyHostToDevice, st));
  WMRL_CUDA_OK(cudaMemcpyAsync(base + pl.off_job_tab, pl.jobs.data(), pl.jobs.size() * sizeof(JobDesc),
                               cudaMemcpyHostToDevice, st));
  uint8_t* ops_dev = base + pl.off_packops;
  WMRL_CUDA_OK(cudaMemcpyAsync(ops_dev, pl.pack_ops.data(), pl.pack_ops.size() * sizeof(PackOp),
                               cudaMemcpyHostToDevice, st));
  uint8_t* vops_dev = ops_dev + pl.pack_ops.size() * sizeof(PackOp);
  WMRL_CUDA_OK(cudaMemcpyAsync(vops_dev, pl.vec_ops.data(), pl.vec_ops.size() * sizeof(VecOp),
                               cudaMemcpyHostToDevice, st));
  pack_stage_kernel<<<(unsigned)pl.pack_ops.size(), 256, 0, st>>>((const PackOp*)ops_dev, base + pl.off_data);
  WMRL_CUDA_OK(cudaGetLastError());
  pack_vec_kernel<<<(unsigned)pl.vec_ops.size(), 128, 0, st>>>((const VecOp*)vops_dev,
                                                                (float*)(base + pl.off_eparams));
  WMRL_CUDA_OK(cudaGetLastError());
  return 0;
}

// --------------------------------------------------------------------------
// the fused kernel
// --------------------------------------------------------------------------
using namespace ptx;

// fine-grained timeline probes (thread 0 of CTA 0, step 1 only; active when tracing)
#define WMRL_PROBE(c, k)                                                                      \
  do {                                                                                        \
    if (kTraceBuild) {                                                                        \
      if ((c).p->trace && blockIdx.x == 0 && threadIdx.x == 0 && (c).t == 1 && (c).b == 0)    \
        (c).p->trace[600 + (k)] = clock64();                                                  \
    }                                                                                         \
  } while (0)

// Epilogue organisation: kQ warps per TMEM lane quarter (a warp may only read
// lanes 32*(warp%4)..+31).  The kQ warps of a quarter own the same 32 batch
// rows and split every job's accumulator COLUMNS between them, so each SM
// sub-partition has kQ resident epilogue warps to hide TMEM / MUFU / constant
// cache latency.  LayerNorm row statistics are combined across the kQ warps
// through a scratch area (the P_a panel, idle during the actor blocks) and a
// named barrier per quarter.
struct EpiCtx {
  const TcParams* p;
  uint32_t smem_base, tmem_lane;   // tmem base with this warp's lane quarter folded in
  int row;                         // row within the tile == TMEM lane
  int q, quarter;                  // column-split index within the quarter, lane quarter
  long long b;                     // global row
  bool valid;
  int t;
};

__device__ __forceinline__ float4 ldc4(uint32_t off) {   // off: float offset, multiple of 4
  return reinterpret_cast<const float4*>(c_eparams)[off >> 2];
}
__device__ __forceinline__ void quarter_sync(int quarter) {
  asm volatile("bar.sync %0, %1;" ::"r"(quarter + 1), "r"(kQ * 32) : "memory");
}

// tile start: deter0 / stoch0 -> index 0 of the outputs and the bf16 panels
__device__ __forceinline__ void epi_init(const EpiCtx& c) {
  const TcParams& p = *c.p;
  const long long T1 = p.H + 1;
  for (int k8 = c.q; k8 < p.Dp / 8; k8 += kQ) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int col = k8 * 8 + e;
      v[e] = (c.valid && col < p.D) ? p.deter0[c.b * p.D + col] : 0.f;
      if (c.valid && col < p.D) p.deter_seq[(c.b * T1) * p.D + col] = v[e];
    }
    st_chunk(c.smem_base + p.off_h + k8 * 2048 + c.row * 16, v);
  }
  for (int k8 = c.q; k8 < p.Sp / 8; k8 += kQ) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int col = k8 * 8 + e;
      v[e] = (c.valid && col < p.S) ? p.stoch0[c.b * p.S + col] : 0.f;
      if (c.valid && col < p.S) p.stoch_seq[(c.b * T1) * p.S + col] = v[e];
    }
    st_chunk(c.smem_base + p.off_s + k8 * 2048 + c.row * 16, v);
  }
}

// Linear -> LayerNorm(eps 1e-5) -> ELU (actor.py:23-36).  This warp owns every kQ-th 32-column group.
template <bool kTraceBuild>
__device__ __forceinline__ void epi_actor_ln(const EpiCtx& c, const JobDesc& jb) {
  const TcParams& p = *c.p;
  const uint32_t bias = jb.p_off, gam = bias + p.Ha, bet = gam + p.Ha;
  const int ngrp = p.Ha / 32;          // 32-column groups, dealt round-robin to the kQ warps
  float sum = 0.f, sq = 0.f;
  WMRL_PROBE(c, 8);
  for (int g = c.q; g < ngrp; g += kQ) {
    float v[32];
    tmem_ld32(c.tmem_lane + jb.tmem_col + g * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 b4 = ldc4(bias + g * 32 + i);
      const float x0 = v[i] + b4.x, x1 = v[i + 1] + b4.y, x2 = v[i + 2] + b4.z, x3 = v[i + 3] + b4.w;
      sum += (x0 + x1) + (x2 + x3);
      sq = fmaf(x0, x0, sq); sq = fmaf(x1, x1, sq); sq = fmaf(x2, x2, sq); sq = fmaf(x3, x3, sq);
    }
  }
  // exchange partials with the other kQ-1 warps of this quarter (scratch = idle P_a panel)
  WMRL_PROBE(c, 9);
  const uint32_t scratch = c.smem_base + p.off_a;
  asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(scratch + (c.q * kTileM + c.row) * 8), "f"(sum), "f"(sq)
               : "memory");
  quarter_sync(c.quarter);
  WMRL_PROBE(c, 10);
  sum = 0.f; sq = 0.f;
#pragma unroll
  for (int k = 0; k < kQ; ++k) {
    float a, b;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(a), "=f"(b) : "r"(scratch + (k * kTileM + c.row) * 8) : "memory");
    sum += a; sq += b;
  }
  const float inv = 1.f / (float)p.Ha;
  const float mean = sum * inv;
  const float var = fmaxf(sq * inv - mean * mean, 0.f);
  const float rstd = rsqrtf(var + 1e-5f);
  const float nmr = -mean * rstd;
  for (int g = c.q; g < ngrp; g += kQ) {
    float v[32];
    tmem_ld32(c.tmem_lane + jb.tmem_col + g * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 b4 = ldc4(bias + g * 32 + i);
      const float4 g4 = ldc4(gam + g * 32 + i);
      const float4 e4 = ldc4(bet + g * 32 + i);
      v[i] = elu_scaled(fmaf(fmaf(v[i] + b4.x, rstd, nmr), g4.x, e4.x));
      v[i + 1] = elu_scaled(fmaf(fmaf(v[i + 1] + b4.y, rstd, nmr), g4.y, e4.y));
      v[i + 2] = elu_scaled(fmaf(fmaf(v[i + 2] + b4.z, rstd, nmr), g4.z, e4.z));
      v[i + 3] = elu_scaled(fmaf(fmaf(v[i + 3] + b4.w, rstd, nmr), g4.w, e4.w));
    }
    const uint32_t dst = c.smem_base + jb.dst_off + (g * 4) * 2048 + c.row * 16;
#pragma unroll
    for (int k = 0; k < 4; ++k) st_chunk(dst + k * 2048, v + k * 8);
  }
  WMRL_PROBE(c, 11);
}

// bias + ELU -> bf16 panel (embed rssm.py:65, prior hidden rssm.py:33-35); 16-column groups round-robin
__device__ __forceinline__ void epi_dense_elu(const EpiCtx& c, const JobDesc& jb) {
  const uint32_t bias = jb.p_off;
  for (int g = c.q; g < jb.n / 16; g += kQ) {
    float v[16];
    tmem_ld16(c.tmem_lane + jb.tmem_col + g * 16, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      const float4 b4 = ldc4(bias + g * 16 + i);
      v[i] = elu_fast(v[i] + b4.x); v[i + 1] = elu_fast(v[i + 1] + b4.y);
      v[i + 2] = elu_fast(v[i + 2] + b4.z); v[i + 3] = elu_fast(v[i + 3] + b4.w);
    }
    const uint32_t dst = c.smem_base + jb.dst_off + (g * 2) * 2048 + c.row * 16;
    st_chunk(dst, v);
    st_chunk(dst + 2048, v + 8);
  }
}

// a = tanh(mu / action_scale) * bound  (actor.py:49-50) -> actions[B,H,A] + P_a panel
__device__ __forceinline__ void epi_mu(const EpiCtx& c, const JobDesc& jb) {
  const TcParams& p = *c.p;
  if (c.q != 0) return;
  const uint32_t bias = jb.p_off, bound = bias + 16;
  float v[16];
  tmem_ld16(c.tmem_lane + jb.tmem_col, v);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 16; i += 4) {
    const float4 b4 = ldc4(bias + i), d4 = ldc4(bound + i);
    v[i] = tanh_fast((v[i] + b4.x) * p.inv_action_scale) * d4.x;
    v[i + 1] = tanh_fast((v[i + 1] + b4.y) * p.inv_action_scale) * d4.y;
    v[i + 2] = tanh_fast((v[i + 2] + b4.z) * p.inv_action_scale) * d4.z;
    v[i + 3] = tanh_fast((v[i + 3] + b4.w) * p.inv_action_scale) * d4.w;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    if (i >= p.A) v[i] = 0.f;
    else if (c.valid) p.actions[(c.b * p.H + c.t) * p.A + i] = v[i];
  }
  const uint32_t dst = c.smem_base + jb.dst_off + c.row * 16;
  st_chunk(dst, v);
  st_chunk(dst + 2048, v + 8);
}

// GRUCell gates (rssm.py:66) for the 16-column groups of one chunk owned by this warp.
// h_prev is the fp32 master copy in deter_seq[:, t] (written by the warp that owns the same
// column group one step earlier: same thread).  It is fetched BEFORE the accumulator barrier
// so the L2 round trip overlaps the MMAs.
constexpr int kGruGroups = (kGruChunk / 16 + kQ - 1) / kQ;   // 16-column groups per warp per chunk
struct GruPrefetch { float hp[16]; };   // h_prev of this warp's FIRST group; later groups are fetched in flight
__device__ __forceinline__ void gru_load_h(const EpiCtx& c, const JobDesc& jb, int g, float (&hp)[16]) {
  const TcParams& p = *c.p;
  const long long T1 = p.H + 1;
  const float* hprev = p.deter_seq + (c.b * T1 + c.t) * p.D;
  const int col = jb.col0 + g * 16;
  const bool in = c.valid && g * 16 < jb.n;
  if (in && col + 16 <= p.D && (p.D & 3) == 0) {
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      const float4 h4 = *reinterpret_cast<const float4*>(hprev + col + i);
      hp[i] = h4.x; hp[i + 1] = h4.y; hp[i + 2] = h4.z; hp[i + 3] = h4.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) hp[i] = (in && col + i < p.D) ? hprev[col + i] : 0.f;
  }
}
__device__ __forceinline__ void gru_prefetch(const EpiCtx& c, const JobDesc& jb, GruPrefetch& pf) {
  gru_load_h(c, jb, c.q, pf.hp);
}
__device__ __forceinline__ void epi_gru(const EpiCtx& c, const JobDesc& jb, uint32_t tmem, const GruPrefetch& pf) {
  const TcParams& p = *c.p;
  const int cw = jb.n;
  const uint32_t b_r = jb.p_off, b_z = b_r + p.Dp, b_in = b_z + p.Dp, b_hn = b_in + p.Dp;
  const long long T1 = p.H + 1;
  float* hnext = p.deter_seq + (c.b * T1 + c.t + 1) * p.D;
  float hp[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) hp[i] = pf.hp[i];
#pragma unroll
  for (int k = 0; k < kGruGroups; ++k) {
    const int g = c.q + k * kQ;
    if (g * 16 < cw) {
      float hn_[16];
      if (k + 1 < kGruGroups) gru_load_h(c, jb, g + kQ, hn_);     // next group's h_prev, in flight during the math
      // two 8-column halves: 4 x 8 accumulator registers live at a time
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int col = jb.col0 + g * 16 + hf * 8;
        float r[8], z[8], gi[8], gh[8];
        tmem_ld8(tmem + g * 16 + hf * 8, r);
        tmem_ld8(tmem + cw + g * 16 + hf * 8, z);
        tmem_ld8(tmem + 2 * cw + g * 16 + hf * 8, gi);
        tmem_ld8(tmem + 3 * cw + g * 16 + hf * 8, gh);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; i += 4) {
          const float4 br = ldc4(b_r + col + i), bz = ldc4(b_z + col + i);
          const float4 bi = ldc4(b_in + col + i), bh = ldc4(b_hn + col + i);
          const float brv[4] = {br.x, br.y, br.z, br.w}, bzv[4] = {bz.x, bz.y, bz.z, bz.w};
          const float biv[4] = {bi.x, bi.y, bi.z, bi.w}, bhv[4] = {bh.x, bh.y, bh.z, bh.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float rr = sigmoid_fast(r[i + e] + brv[e]);
            const float zz = sigmoid_fast(z[i + e] + bzv[e]);
            const float nn = tanh_fast(gi[i + e] + biv[e] + rr * (gh[i + e] + bhv[e]));
            r[i + e] = fmaf(zz, hp[hf * 8 + i + e] - nn, nn);   // (1-z)*n + z*h
          }
        }
        if (c.valid) {
          if (col + 8 <= p.D && (p.D & 3) == 0) {
            *reinterpret_cast<float4*>(hnext + col) = make_float4(r[0], r[1], r[2], r[3]);
            *reinterpret_cast<float4*>(hnext + col + 4) = make_float4(r[4], r[5], r[6], r[7]);
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (col + i < p.D) hnext[col + i] = r[i];
          }
        }
        st_chunk(c.smem_base + jb.dst_off + (col / 8) * 2048 + c.row * 16, r);
      }
      if (k + 1 < kGruGroups) {
#pragma unroll
        for (int i = 0; i < 16; ++i) hp[i] = hn_[i];
      }
    }
  }
}

// Gaussian head (rssm.py:179-181) + end-of-step hand-over of h' into the P_h panel.
// 8-column groups dealt round-robin to the kQ warps; eps is fetched before the accumulator wait.
struct EpsPrefetch { float e[kP2Groups][8]; };
__device__ __forceinline__ void eps_prefetch(const EpiCtx& c, EpsPrefetch& pf) {
  const TcParams& p = *c.p;
  const float* eps = p.noise + ((long long)c.t * p.B + c.b) * p.S;
#pragma unroll
  for (int k = 0; k < kP2Groups; ++k) {
    const int col = (c.q + k * kQ) * 8;
    if (c.valid && col + 8 <= p.S && (p.S & 1) == 0) {
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        const float2 e2 = *reinterpret_cast<const float2*>(eps + col + i);
        pf.e[k][i] = e2.x; pf.e[k][i + 1] = e2.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) pf.e[k][i] = (c.valid && col + i < p.S) ? eps[col + i] : 0.f;
    }
  }
}
template <bool kTraceBuild>
__device__ __forceinline__ void epi_prior2_cont(const EpiCtx& c, const JobDesc& jb, const EpsPrefetch& pf) {
  const TcParams& p = *c.p;
  WMRL_PROBE(c, 0);
  const uint32_t b_m = jb.p_off, b_s = b_m + p.Sp;
  const long long T1 = p.H + 1;
  const long long o = (c.b * T1 + c.t + 1) * p.S;
#pragma unroll
  for (int k = 0; k < kP2Groups; ++k) {
    const int g = c.q + k * kQ;
    if (g >= p.Sp / 8) break;
    const int col = g * 8;
    float m[8], s[8], sd[8], st[8];
    tmem_ld8(c.tmem_lane + jb.tmem_col + col, m);
    tmem_ld8(c.tmem_lane + jb.tmem_col + p.Sp + col, s);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 8; i += 4) {
      const float4 bm = ldc4(b_m + col + i), bs = ldc4(b_s + col + i);
      const float bmv[4] = {bm.x, bm.y, bm.z, bm.w}, bsv[4] = {bs.x, bs.y, bs.z, bs.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float mean = m[i + e] + bmv[e];
        const float raw = s[i + e] + bsv[e];
        // softplus(raw) = max(raw,0) + log1p(exp(-|raw|))  (no threshold branch needed)
        const float sp = fmaxf(raw, 0.f) + __logf(1.f + ex2_fast(-fabsf(raw) * 1.4426950408889634f));
        const float ee = pf.e[k][i + e];
        m[i + e] = mean;
        sd[i + e] = sp + 0.1f;
        st[i + e] = (col + i + e < p.S) ? fmaf(sd[i + e], ee, mean) : 0.f;
      }
    }
    if (p.stage_out) {
      // fp32 staging [field][row][Sp] in X[0..h2) (idle: prior2's MMAs have completed); 16-byte chunks
      // XOR-swizzled by (row & 7) so both these row-strided writes and the row-contiguous reads below
      // are bank-conflict free
      const uint32_t rowb = c.smem_base + p.off_x + (uint32_t)c.row * (uint32_t)p.Sp * 4u;
      const uint32_t fstride = (uint32_t)kTileM * (uint32_t)p.Sp * 4u;
      const uint32_t c0 = (uint32_t)((2 * g) ^ (c.row & 7)) * 16u, c1 = (uint32_t)((2 * g + 1) ^ (c.row & 7)) * 16u;
      asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(rowb + c0), "f"(m[0]), "f"(m[1]), "f"(m[2]), "f"(m[3]) : "memory");
      asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(rowb + c1), "f"(m[4]), "f"(m[5]), "f"(m[6]), "f"(m[7]) : "memory");
      asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(rowb + fstride + c0), "f"(sd[0]), "f"(sd[1]), "f"(sd[2]), "f"(sd[3]) : "memory");
      asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(rowb + fstride + c1), "f"(sd[4]), "f"(sd[5]), "f"(sd[6]), "f"(sd[7]) : "memory");
      asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(rowb + 2 * fstride + c0), "f"(st[0]), "f"(st[1]), "f"(st[2]), "f"(st[3]) : "memory");
      asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(rowb + 2 * fstride + c1), "f"(st[4]), "f"(st[5]), "f"(st[6]), "f"(st[7]) : "memory");
    } else if (c.valid) {
      if (col + 8 <= p.S && (p.S & 1) == 0) {
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          *reinterpret_cast<float2*>(p.mean_seq + o + col + i) = make_float2(m[i], m[i + 1]);
          *reinterpret_cast<float2*>(p.std_seq + o + col + i) = make_float2(sd[i], sd[i + 1]);
          *reinterpret_cast<float2*>(p.stoch_seq + o + col + i) = make_float2(st[i], st[i + 1]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (col + i < p.S) {
            p.mean_seq[o + col + i] = m[i];
            p.std_seq[o + col + i] = sd[i];
            p.stoch_seq[o + col + i] = st[i];
          }
      }
    }
    st_chunk(c.smem_base + jb.dst_off + g * 2048 + c.row * 16, st);
  }
  WMRL_PROBE(c, 1);
  // h' (X[h2..]) -> P_h: every MMA that read the old P_h has completed (this job's
  // accumulator barrier is committed after them)
  for (int k8 = c.q; k8 < p.Dp / 8; k8 += kQ) {
    uint32_t a, b, cc, d;
    const uint32_t src = c.smem_base + p.off_h2 + k8 * 2048 + c.row * 16;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(cc), "=r"(d) : "r"(src) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(c.smem_base + p.off_h + k8 * 2048 + c.row * 16),
                 "r"(a), "r"(b), "r"(cc), "r"(d)
                 : "memory");
  }
  WMRL_PROBE(c, 2);
  if (p.stage_out) {
    // all epilogue warps have staged their columns -> each warp streams whole rows out:
    // one row-field (S contiguous floats) per pass, lanes along the row = coalesced
    asm volatile("bar.sync 5, %0;" ::"r"(kEpiThreads) : "memory");
    WMRL_PROBE(c, 3);
    const int warp = c.q * 4 + c.quarter;
    const int lane = c.row & 31;
    const long long tile_b0 = c.b - c.row;
    const uint32_t fstride = (uint32_t)kTileM * (uint32_t)p.Sp * 4u;
    const uint32_t xbase = c.smem_base + p.off_x;
    for (int row = warp; row < kTileM; row += kEpiThreads / 32) {
      const long long bb = tile_b0 + row;
      if (bb >= p.B) break;
      const long long off = (bb * T1 + c.t + 1) * p.S;
      const uint32_t rb = xbase + (uint32_t)row * (uint32_t)p.Sp * 4u;
      for (int col = lane; col < p.S; col += 32) {
        const uint32_t a0 = rb + (uint32_t)(((col >> 2) ^ (row & 7)) * 16 + (col & 3) * 4);
        float v0, v1, v2;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v0) : "r"(a0));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v1) : "r"(a0 + fstride));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v2) : "r"(a0 + 2 * fstride));
        p.mean_seq[off + col] = v0;
        p.std_seq[off + col] = v1;
        p.stoch_seq[off + col] = v2;
      }
    }
  }
  WMRL_PROBE(c, 4);
}

template <int kCta, bool kTraceBuild>
__global__ void __launch_bounds__(kThreads, 1) tc_imagine_kernel(const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr uint32_t kSlot = kSlotBytes / kCta;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t ring = smem_base + p.ring_off;
  const uint32_t bars = smem_base + p.bar_off;
  // barrier map (8 B each): full[16] @0 | empty[16] @128 | peer_full[16] @256 | acc_full[2] @384 |
  // acc_empty[2] @400 | tmem ptr @416; the stage records follow at +kBarBytes
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 128u + 8u * s; };
  auto peer_full_bar = [&](int s) { return bars + 256u + 8u * s; };
  auto accf_bar = [&](uint32_t j) { return bars + 384u + 8u * (j & 1u); };
  auto acce_bar = [&](uint32_t j) { return bars + 400u + 8u * (j & 1u); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + p.bar_off + 416);
  // warp index broadcast through a shuffle: the compiler then treats it (and every epilogue
  // parameter address derived from it) as warp-uniform -> uniform-datapath constant loads
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  // CTA pair: rank 0 is the leader (issues every MMA for both SMs)
  const uint32_t rank = (kCta == 2) ? cluster_ctarank() : 0u;
  const int unit = (kCta == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int nunits = (kCta == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int ngroups = (p.ntiles + kCta - 1) / kCta;     // a group = kCta consecutive 128-row tiles

  // The per-stage descriptors are on the critical path of the single-lane producer / MMA loops; the
  // constant cache is being streamed through by the epilogue warps, so mirror the table in smem.
  // Each 16-byte record is everything the MMA lane needs, precomputed here in parallel:
  //   x = A descriptor low word (address | LBO), y = instruction descriptor,
  //   z = B LBO field (rows per CTA) << 16 | nk16 << 8 | flags, w = tmem column | dep << 16
  const uint4* stab = reinterpret_cast<const uint4*>(smem + p.bar_off + kBarBytes);
  for (int i = threadIdx.x; i < p.nstages; i += kThreads) {
    const uint4 raw = c_stages[i];
    const StageDesc st = *reinterpret_cast<const StageDesc*>(&raw);
    uint4 rec;
    rec.x = (uint32_t)make_smem_desc(smem_base + (uint32_t)st.a_off16 * 16u, 2048u, 128u);
    rec.y = make_idesc_bf16(kTileM * kCta, st.n);
    rec.z = ((uint32_t)(st.n / kCta) << 16) | ((uint32_t)st.nk16 << 8) | st.flags;
    rec.w = (uint32_t)st.tmem_col | ((uint32_t)st.dep << 16);
    reinterpret_cast<uint4*>(smem + p.bar_off + kBarBytes)[i] = rec;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.nslots; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
      mbar_init(peer_full_bar(s), 1);
    }
    for (uint32_t j = 0; j < 2; ++j) {
      mbar_init(accf_bar(j), 1);
      mbar_init(acce_bar(j), (kEpiThreads / 32) * kCta);   // one arrival per epilogue warp of the group
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    if (kCta == 2) { tmem_alloc2(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512); tmem_relinquish2(); }
    else { tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512); tmem_relinquish(); }
  }
  tc_fence_before();
  if (kCta == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kProducerWarp) {
    // ===================== TMA producer: stream this CTA's share of every weight stage
    if (lane == 0) {
      uint32_t slot = 0, phase = 0;
      for (int grp = unit; grp < ngroups; grp += nunits)
        for (int t = 0; t < p.H; ++t) {
          const bool trp = kTraceBuild && p.trace && blockIdx.x == 0 && grp == unit && t == 1;
          long long pw = 0, pi = 0;
          for (int s = 0; s < p.nstages; ++s) {
            const uint4 raw = c_stages[s];
            const uint32_t bytes = ((raw.y & 0xFFFFu) * 16u) / kCta;
            const long long c0 = trp ? clock64() : 0;
            mbar_wait(empty_bar(slot), phase ^ 1u);
            const long long c1 = trp ? clock64() : 0;
            mbar_arrive_expect_tx(full_bar(slot), bytes);
            bulk_g2s(ring + slot * kSlot, p.stage_data + (size_t)raw.x * 16u + (size_t)rank * bytes, bytes,
                     full_bar(slot));
            if (trp && s >= 16 && s < 48) { pw += c1 - c0; pi += clock64() - c1; }
            if (++slot == (uint32_t)p.nslots) { slot = 0; phase ^= 1u; }
          }
          if (trp) { p.trace[710] = pw; p.trace[711] = pi; }
        }
    }
  } else if (warp == kMmaWarp) {
    if (lane == 0 && rank != 0) {
      // ===================== peer CTA: forward "my half of the stage has landed" to the leader
      uint32_t slot = 0, phase = 0;
      for (int grp = unit; grp < ngroups; grp += nunits)
        for (int t = 0; t < p.H; ++t)
          for (int s = 0; s < p.nstages; ++s) {
            mbar_wait(full_bar(slot), phase);
            mbar_arrive_remote(mapa(peer_full_bar(slot), 0));
            if (++slot == (uint32_t)p.nslots) { slot = 0; phase ^= 1u; }
          }
    } else if (lane == 0) {
      // ===================== MMA issuer (leader)
      uint32_t slot = 0, phase = 0, job = 0;
      for (int grp = unit; grp < ngroups; grp += nunits) {
        mbar_arrive(accf_bar(job));   // INIT job has no MMA
        if (kCta == 2) mbar_arrive_remote(mapa(accf_bar(job), 1));
        ++job;
        const bool tr = kTraceBuild && p.trace && blockIdx.x == 0 && grp == unit;
        for (int t = 0; t < p.H; ++t) {
          int jidx = -1;
          long long acc_dep = 0, acc_full = 0, acc_peer = 0, a1_full = 0, a1_peer = 0, a1_issue = 0;
          long long g_dep = 0, g_full = 0, g_issue = 0, g_n = 0;
          for (int s = 0; s < p.nstages; ++s) {
            const uint4 rec = stab[s];
            const uint32_t flags = rec.z & 0xFFu;
            if (flags & ST_FIRST_OF_JOB) ++jidx;
            const bool trw = tr && t == 1;
            long long w0 = trw ? clock64() : 0;
            if (flags & ST_FIRST_OF_JOB) {
              const uint32_t dep = rec.w >> 16;
              if (dep) {
                const uint32_t dj = job - dep;
                mbar_wait(acce_bar(dj), (dj >> 1) & 1u);
              }
            }
            long long w1 = trw ? clock64() : 0;
            mbar_wait(full_bar(slot), phase);
            long long w2 = trw ? clock64() : 0;
            if (kCta == 2) mbar_wait(peer_full_bar(slot), phase);
            tc_fence_after();
            long long w3 = 0;
            if (trw) {
              w3 = clock64();
              acc_dep += w1 - w0; acc_full += w2 - w1; acc_peer += w3 - w2;
              if (s >= 16 && s < 48) { a1_full += w2 - w1; a1_peer += w3 - w2; }
              if (flags & ST_DBUF) { g_dep += w1 - w0; g_full += w2 - w1; g_n += 1; }
            }
            if (tr && (flags & ST_FIRST_OF_JOB) && t < 4) p.trace[((t * p.njobs) + jidx) * 4 + 0] = clock64();
            const uint32_t tm = tmem_base + (rec.w & 0xFFFFu) + ((flags & ST_DBUF) ? (job & 1u) * 256u : 0u);
            const uint32_t idesc = rec.y;
            const uint32_t nh = rec.z >> 16;                         // B rows held by each CTA
            uint32_t alo = rec.x;
            uint32_t blo = (((ring + slot * kSlot) >> 4) & 0x3FFFu) | (nh << 16);   // address | LBO = nh*16 B
            const uint32_t bstep = nh * 2u;                          // (nh*32 bytes) >> 4
            constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);   // SBO = 128 B, descriptor version 1
            uint32_t acc = (flags & ST_FIRST_ACC) ? 0u : 1u;
            const uint32_t nk = (rec.z >> 8) & 0xFFu;
#pragma unroll 2
            for (uint32_t i = 0; i < nk; ++i) {
              if (kCta == 2) mma_bf16_lohi2(tm, alo, kDescHi, blo, kDescHi, idesc, acc);
              else mma_bf16_lohi(tm, alo, kDescHi, blo, kDescHi, idesc, acc);
              alo += 256u;                                           // 4096 bytes >> 4
              blo += bstep;
              acc = 1u;
            }
            if (kCta == 2) tc_commit2(empty_bar(slot), 3); else tc_commit(empty_bar(slot));
            if (flags & ST_LAST_OF_JOB) {
              if (kCta == 2) tc_commit2(accf_bar(job), 3); else tc_commit(accf_bar(job));
              if (tr && t < 4) p.trace[((t * p.njobs) + jidx) * 4 + 1] = clock64();
              ++job;
            }
            if (trw && s >= 16 && s < 48) a1_issue += clock64() - w3;
            if (trw && (flags & ST_DBUF)) g_issue += clock64() - w3;
            if (++slot == (uint32_t)p.nslots) { slot = 0; phase ^= 1u; }
          }
          if (tr && t == 1) {
            p.trace[704] = a1_full; p.trace[705] = a1_peer; p.trace[706] = a1_issue;
            p.trace[712] = g_dep; p.trace[713] = g_full; p.trace[714] = g_issue; p.trace[715] = g_n;
            p.trace[700] = acc_dep; p.trace[701] = acc_full; p.trace[702] = acc_peer; p.trace[703] = p.nstages;
          }
        }
      }
    }
  } else {
    // ===================== epilogue warps: one batch row (TMEM lane) per thread, columns split kQ ways
    EpiCtx c;
    c.p = &p;
    c.smem_base = smem_base;
    c.quarter = warp & 3;
    c.q = warp >> 2;
    c.row = c.quarter * 32 + lane;
    c.tmem_lane = tmem_base + ((uint32_t)(c.quarter * 32) << 16);
    // job_done: every thread has fenced its smem writes / TMEM reads; one arrival per warp on the leader
#define WMRL_JOB_DONE(job)                                                              \
  do {                                                                                  \
    __syncwarp();                                                                       \
    if (lane == 0) {                                                                    \
      if (kCta == 2 && rank != 0) mbar_arrive_remote(mapa(acce_bar(job), 0));           \
      else mbar_arrive(acce_bar(job));                                                  \
    }                                                                                   \
  } while (0)
#define WMRL_WAIT_ACC(job) mbar_wait_backoff(accf_bar(job), ((job) >> 1) & 1u)
    uint32_t job = 0;
    for (int grp = unit; grp < ngroups; grp += nunits) {
      const int tile = grp * kCta + (int)rank;
      c.b = (long long)tile * kTileM + c.row;
      c.valid = c.b < p.B;
      c.t = 0;
      WMRL_WAIT_ACC(job);
      epi_init(c);
      fence_proxy_async_smem();
      WMRL_JOB_DONE(job);
      ++job;
      const bool tr = kTraceBuild && p.trace && blockIdx.x == 0 && grp == unit && threadIdx.x == 0;
      for (int t = 0; t < p.H; ++t) {
        c.t = t;
        for (int jb_i = 0; jb_i < p.njobs; ++jb_i) {
          const JobDesc jb = c_jobs[jb_i];
          long long t_wake = 0;
          if (jb.type == JOB_GRU) {
            GruPrefetch pf;
            gru_prefetch(c, jb, pf);
            WMRL_WAIT_ACC(job);
            tc_fence_after();
            if (tr) t_wake = clock64();
            epi_gru(c, jb, c.tmem_lane + (jb.dbuf ? (job & 1u) * 256u : 0u), pf);
          } else if (jb.type == JOB_PRIOR2_CONT) {
            EpsPrefetch pf;
            eps_prefetch(c, pf);
            WMRL_WAIT_ACC(job);
            tc_fence_after();
            if (tr) t_wake = clock64();
            epi_prior2_cont<kTraceBuild>(c, jb, pf);
          } else {
            WMRL_WAIT_ACC(job);
            tc_fence_after();
            if (tr) t_wake = clock64();
            if (jb.type == JOB_ACTOR_LN) epi_actor_ln<kTraceBuild>(c, jb);
            else if (jb.type == JOB_DENSE_ELU) epi_dense_elu(c, jb);
            else if (jb.type == JOB_MU) epi_mu(c, jb);
          }
          WMRL_PROBE(c, 16 + jb_i * 2);
          tc_fence_before();
          fence_proxy_async_smem();
          WMRL_PROBE(c, 17 + jb_i * 2);
          if (tr && t < 4) {
            p.trace[((t * p.njobs) + jb_i) * 4 + 2] = t_wake;
            p.trace[((t * p.njobs) + jb_i) * 4 + 3] = clock64();
          }
          WMRL_JOB_DONE(job);
          ++job;
        }
      }
    }
  }
  tc_fence_before();
  if (kCta == 2) cluster_sync_all(); else __syncthreads();
  if (warp == kMmaWarp) {
    if (kCta == 2) tmem_dealloc2(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

// ==========================================================================
// Folded CTA-pair variant (mode 3): cta_group::2 with M=128, 64 batch rows per CTA.
//
// The D tile of such an MMA is folded over the TMEM lanes: lane r (0..63) holds accumulator columns
// [0,N/2) of row r, lane 64+r holds columns [N/2,N) at the same TMEM column offsets (measured,
// tools/micro/tmem_layout.cu).  So every batch row is served by TWO lane halves x kQ warps = 8 threads,
// each accumulator group takes only N/2 TMEM columns, the A panels are [k/8][64 rows][8] (1 KB per
// chunk, half the bytes), and the freed shared memory holds a 16-slot weight ring.
// ==========================================================================
constexpr uint32_t kPSF = 1024;   // bytes per 8-column chunk of a 64-row panel
constexpr int kRowsF = 64;

struct EpiCtxF {
  const TcParams* p;
  uint32_t smem_base, tmem_lane;
  int row;        // row within this CTA's 64-row half tile
  int hf;         // lane half: 0 -> accumulator columns [0,N/2), 1 -> [N/2,N)
  int q;          // column-split index among the kQ warps of this lane quarter
  int sidx;       // hf*kQ + q: index among the 8 threads that share a row
  int quarter;
  long long b;    // global row
  bool valid;
  int t;
  uint8_t* sblk;       // training: this (tile, CTA, step)'s forward stash block, already offset by row*16
  uint8_t* sblk_next;  // the block of step t+1 (its `feat` field is written during step t); null in the last step
};
__device__ __forceinline__ void rowpair_sync(int quarter) {
  // the 2*kQ warps (two lane quarters) that share the same 32 batch rows
  asm volatile("bar.sync %0, %1;" ::"r"((quarter & 1) + 1), "r"(2 * kQ * 32) : "memory");
}

template <bool kSave>
__device__ __forceinline__ void epif_init(const EpiCtxF& c) {
  const TcParams& p = *c.p;
  const long long T1 = p.H + 1;
  for (int k8 = c.sidx; k8 < p.Dp / 8; k8 += 2 * kQ) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int col = k8 * 8 + e;
      v[e] = (c.valid && col < p.D) ? p.deter0[c.b * p.D + col] : 0.f;
      if (c.valid && col < p.D) p.deter_seq[(c.b * T1) * p.D + col] = v[e];
    }
    st_chunk(c.smem_base + p.off_h + k8 * kPSF + c.row * 16, v);
    st_chunk(c.smem_base + p.off_h2 + k8 * kPSF + c.row * 16, v);   // read by the early first-actor-layer MMAs (ST_PRE)
    if (kSave) st_chunk_g(c.sblk + p.sl.feat + k8 * kStashChunk, v);
  }
  for (int k8 = c.sidx; k8 < p.Sp / 8; k8 += 2 * kQ) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int col = k8 * 8 + e;
      v[e] = (c.valid && col < p.S) ? p.stoch0[c.b * p.S + col] : 0.f;
      if (c.valid && col < p.S) p.stoch_seq[(c.b * T1) * p.S + col] = v[e];
    }
    st_chunk(c.smem_base + p.off_s + k8 * kPSF + c.row * 16, v);
    if (kSave) st_chunk_g(c.sblk + p.sl.feat + (p.Dp / 8 + k8) * kStashChunk, v);
  }
  // action panel: zeros (it is read as a K segment before the first mu epilogue writes it) and the constant one
  // in column 15 that carries the folded biases
  if (c.sidx == 0) {
    const float z8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float o8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 1.f};
    st_chunk(c.smem_base + p.off_a + c.row * 16, z8);
    st_chunk(c.smem_base + p.off_a + kPSF + c.row * 16, o8);
  }
}

// Linear -> LayerNorm -> ELU.  MMA group mg (<= 256 features) sits in TMEM columns [mg*128, +Ng/2); this lane
// half owns features mg*256 + hf*Ng/2 + [0, Ng/2).  32-column sub-groups are dealt round-robin to the kQ warps.
// kBias = false: the Linear bias is already in the accumulators (folded into the GEMM, JobDesc.aux & 1)
template <bool kBias, bool kSave>
__device__ __forceinline__ void epif_actor_ln(const EpiCtxF& c, const JobDesc& jb) {
  const TcParams& p = *c.p;
  const uint32_t bias = jb.p_off, gam = bias + p.Ha, bet = gam + p.Ha;
  const int nmg = (p.Ha + 255) / 256;
  float sum = 0.f, sq = 0.f, sum1 = 0.f, sq1 = 0.f;   // two partial sums each: packed f32x2 arithmetic
  auto pass1 = [&](const float (&v)[32], int f0) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float x0 = v[i], x1 = v[i + 1], x2 = v[i + 2], x3 = v[i + 3];
      if (kBias) {
        const float4 b4 = ldc4(bias + f0 + i);
        add2(x0, x1, x0, x1, b4.x, b4.y);
        add2(x2, x3, x2, x3, b4.z, b4.w);
      }
      add2(sum, sum1, sum, sum1, x0, x1);
      add2(sum, sum1, sum, sum1, x2, x3);
      fma2(sq, sq1, x0, x1, x0, x1, sq, sq1);
      fma2(sq, sq1, x2, x3, x2, x3, sq, sq1);
    }
  };
  int li = 0;
  if (p.Ha == 512) {
    // the usual width: this thread's two sub-groups (one per MMA group); both TMEM loads in flight before the wait
    float v0[32], v1[32];
    tmem_ld32(c.tmem_lane + jb.tmem_col + c.q * 32, v0);
    tmem_ld32(c.tmem_lane + jb.tmem_col + 128 + c.q * 32, v1);
    tmem_ld_wait();
    pass1(v0, c.hf * 128 + c.q * 32);
    pass1(v1, 256 + c.hf * 128 + c.q * 32);
  } else {
    for (int mg = 0; mg < nmg; ++mg) {
      const int hw = min(256, p.Ha - mg * 256) / 2;
      for (int sub = 0; sub < hw / 32; ++sub, ++li) {
        if ((li % kQ) != c.q) continue;
        float v[32];
        tmem_ld32(c.tmem_lane + jb.tmem_col + mg * 128 + sub * 32, v);
        tmem_ld_wait();
        pass1(v, mg * 256 + c.hf * hw + sub * 32);
      }
    }
  }
  // 8 partials per row (2 lane halves x kQ warps) exchanged through the head of the X panel, which is idle
  // between this layer's MMAs (done) and its own pass-2 writes (after the second barrier)
  sum += sum1;
  sq += sq1;
  const uint32_t scratch = c.smem_base + p.off_x;
  asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(scratch + (c.sidx * kRowsF + c.row) * 8), "f"(sum), "f"(sq)
               : "memory");
  rowpair_sync(c.quarter);
  sum = 0.f; sq = 0.f;
#pragma unroll
  for (int k = 0; k < 2 * kQ; ++k) {
    float a, b;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(a), "=f"(b) : "r"(scratch + (k * kRowsF + c.row) * 8) : "memory");
    sum += a; sq += b;
  }
  // every epilogue thread of the CTA must have read its partials before anybody's pass 2 overwrites the
  // scratch (it aliases the head of X, shared by both 32-row groups)
  asm volatile("bar.sync 6, %0;" ::"r"(kEpiThreads) : "memory");
  const float inv = 1.f / (float)p.Ha;
  const float mean = sum * inv;
  const float var = fmaxf(sq * inv - mean * mean, 0.f);
  const float rstd = rsqrtf(var + 1e-5f);
  const float nmr = -mean * rstd;
  if (kSave && c.sidx == 0)   // one thread per row: rstd[layer][row] (sblk carries row*16; the fp32 table is row*4)
    *reinterpret_cast<float*>(c.sblk - c.row * 16 + p.sl.rstd + (jb.layer * kRowsF + c.row) * 4) = rstd;
  li = 0;
  for (int mg = 0; mg < nmg; ++mg) {
    const int hw = min(256, p.Ha - mg * 256) / 2;
    for (int sub = 0; sub < hw / 32; ++sub, ++li) {
      if ((li % kQ) != c.q) continue;
      const int f0 = mg * 256 + c.hf * hw + sub * 32;
      float v[32];
      float xh[8];
      tmem_ld32(c.tmem_lane + jb.tmem_col + mg * 128 + sub * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 g4 = ldc4(gam + f0 + i);
        const float4 e4 = ldc4(bet + f0 + i);
        float t0 = v[i], t1 = v[i + 1], t2 = v[i + 2], t3 = v[i + 3];
        if (kBias) {
          const float4 b4 = ldc4(bias + f0 + i);
          add2(t0, t1, t0, t1, b4.x, b4.y);
          add2(t2, t3, t2, t3, b4.z, b4.w);
        }
        fma2(t0, t1, t0, t1, rstd, rstd, nmr, nmr);
        fma2(t2, t3, t2, t3, rstd, rstd, nmr, nmr);
        if (kSave) {   // x-hat, 8 columns at a time -> stash panel
          xh[(i & 4) + 0] = t0; xh[(i & 4) + 1] = t1; xh[(i & 4) + 2] = t2; xh[(i & 4) + 3] = t3;
          if (i & 4) st_chunk_g(c.sblk + jb.st1 + (f0 / 8 + (i >> 3)) * kStashChunk, xh);
        }
        fma2(t0, t1, t0, t1, g4.x, g4.y, e4.x, e4.y);
        fma2(t2, t3, t2, t3, g4.z, g4.w, e4.z, e4.w);
        elu_scaled2(v[i], v[i + 1], t0, t1);
        elu_scaled2(v[i + 2], v[i + 3], t2, t3);
      }
      const uint32_t dst = c.smem_base + jb.dst_off + (f0 / 8) * kPSF + c.row * 16;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (kSave) st_chunk_sg(dst + k * kPSF, c.sblk + jb.st0 + (f0 / 8 + k) * kStashChunk, v + k * 8);
        else st_chunk(dst + k * kPSF, v + k * 8);
      }
    }
  }
}

// bias + ELU -> bf16 panel; 8-column sub-groups of this lane half's features, dealt to the kQ warps.  A thread has
// only a few sub-groups, so the job is TMEM-latency-bound: its loads are issued four at a time before one wait.
template <bool kSave>
__device__ __forceinline__ void epif_dense_elu(const EpiCtxF& c, const JobDesc& jb) {
  const uint32_t bias = jb.p_off;
  const int nmg = (jb.n + 255) / 256;
  for (int mg = 0; mg < nmg; ++mg) {
    const int hw = min(256, (int)jb.n - mg * 256) / 2;
    const int nsub = hw / 8;
    // this thread's sub-groups: sub = first, first + kQ, ... (li = sub + mg * (sub-groups of earlier groups) keeps the
    // round-robin of the unbatched version: groups before this one have 128/8 = 16 sub-groups, a multiple of kQ)
    for (int sub0 = c.q; sub0 < nsub; sub0 += 4 * kQ) {
      float v[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int sub = sub0 + u * kQ;
        if (sub < nsub) tmem_ld8(c.tmem_lane + jb.tmem_col + mg * 128 + sub * 8, v[u]);
      }
      tmem_ld_wait();
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int sub = sub0 + u * kQ;
        if (sub < nsub) {
          const int f0 = mg * 256 + c.hf * hw + sub * 8;
          const float4 b0 = ldc4(bias + f0), b1 = ldc4(bias + f0 + 4);
          constexpr float kL2e = 1.4426950408889634f;
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
          for (int e = 0; e < 8; e += 2) {
            float t0, t1;
            add2(t0, t1, v[u][e], v[u][e + 1], bb[e], bb[e + 1]);
            fma2(t0, t1, t0, t1, kL2e, kL2e, 0.f, 0.f);
            elu_scaled2(v[u][e], v[u][e + 1], t0, t1);
          }
          if (kSave) st_chunk_sg(c.smem_base + jb.dst_off + (f0 / 8) * kPSF + c.row * 16,
                                 c.sblk + jb.st0 + (f0 / 8) * kStashChunk, v[u]);
          else st_chunk(c.smem_base + jb.dst_off + (f0 / 8) * kPSF + c.row * 16, v[u]);
        }
      }
    }
  }
}

// action head (N=16): lane half 0 holds actions 0..7, half 1 actions 8..15 (zero padding when A <= 8)
__device__ __forceinline__ void epif_mu(const EpiCtxF& c, const JobDesc& jb) {
  const TcParams& p = *c.p;
  if (c.q != 0) return;
  const uint32_t bias = jb.p_off + c.hf * 8, bound = jb.p_off + 16 + c.hf * 8;
  float v[8];
  tmem_ld8(c.tmem_lane + jb.tmem_col, v);
  tmem_ld_wait();
  const float4 b0 = ldc4(bias), b1 = ldc4(bias + 4), d0 = ldc4(bound), d1 = ldc4(bound + 4);
  const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  const float dv[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int a = c.hf * 8 + i;
    // padding columns are zero, except column 15: the constant one the folded biases multiply (JobDesc.aux & 1)
    v[i] = (a < p.A) ? tanh_fast((v[i] + bv[i]) * p.inv_action_scale) * dv[i] : ((a == 15 && (jb.aux & 1u)) ? 1.f : 0.f);
    if (c.valid && a < p.A) p.actions[(c.b * p.H + c.t) * p.A + a] = v[i];
  }
  st_chunk(c.smem_base + jb.dst_off + c.hf * kPSF + c.row * 16, v);
}

// GRU gates: this lane half owns state columns col0 + hf*cw/2 + [0, cw/2); per lane half the chunk's TMEM
// is [r | z | i_n | h_n], cw/2 columns each.  8-column sub-groups dealt to the kQ warps.
__device__ __forceinline__ void gruf_load_h(const EpiCtxF& c, const JobDesc& jb, int sub, float (&hp)[8]) {
  const TcParams& p = *c.p;
  const long long T1 = p.H + 1;
  const float* hprev = p.deter_seq + (c.b * T1 + c.t) * p.D;
  const int hw = jb.n / 2;
  const int col = jb.col0 + c.hf * hw + sub * 8;
  const bool in = c.valid && sub * 8 < hw;
  if (in && col + 8 <= p.D && (p.D & 3) == 0) {
    const float4 a = *reinterpret_cast<const float4*>(hprev + col);
    const float4 b = *reinterpret_cast<const float4*>(hprev + col + 4);
    hp[0] = a.x; hp[1] = a.y; hp[2] = a.z; hp[3] = a.w; hp[4] = b.x; hp[5] = b.y; hp[6] = b.z; hp[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) hp[i] = (in && col + i < p.D) ? hprev[col + i] : 0.f;
  }
}
constexpr int kGruSubsF = (kGruChunk / 2 / 8 + kQ - 1) / kQ;   // 8-column sub-groups per warp per chunk
template <bool kSave>
__device__ __forceinline__ void epif_gru(const EpiCtxF& c, const JobDesc& jb, uint32_t tmem, const float (&hp0)[8]) {
  const TcParams& p = *c.p;
  const int cw = jb.n, hw = cw / 2;
  const uint32_t b_r = jb.p_off, b_z = b_r + p.Dp, b_in = b_z + p.Dp, b_hn = b_in + p.Dp;
  const long long T1 = p.H + 1;
  float* hnext = p.deter_seq + (c.b * T1 + c.t + 1) * p.D;
  float hp[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) hp[i] = hp0[i];
#pragma unroll
  for (int k = 0; k < kGruSubsF; ++k) {
    const int sub = c.q + k * kQ;
    if (sub * 8 < hw) {
      float hn_[8];
      if (k + 1 < kGruSubsF) gruf_load_h(c, jb, sub + kQ, hn_);
      const int col = jb.col0 + c.hf * hw + sub * 8;
      float r[8], z[8], gi[8], gh[8];
      float rs[8];
      tmem_ld8(tmem + sub * 8, r);
      tmem_ld8(tmem + hw + sub * 8, z);
      tmem_ld8(tmem + cw + sub * 8, gi);
      tmem_ld8(tmem + cw + hw + sub * 8, gh);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 8; i += 4) {
        const float4 br = ldc4(b_r + col + i), bz = ldc4(b_z + col + i);
        const float4 bi = ldc4(b_in + col + i), bh = ldc4(b_hn + col + i);
        const float brv[4] = {br.x, br.y, br.z, br.w}, bzv[4] = {bz.x, bz.y, bz.z, bz.w};
        const float biv[4] = {bi.x, bi.y, bi.z, bi.w}, bhv[4] = {bh.x, bh.y, bh.z, bh.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float rr = sigmoid_fast(r[i + e] + brv[e]);
          const float zz = sigmoid_fast(z[i + e] + bzv[e]);
          const float hn = gh[i + e] + bhv[e];
          const float nn = tanh_fast(gi[i + e] + biv[e] + rr * hn);
          r[i + e] = fmaf(zz, hp[i + e] - nn, nn);
          if (kSave) { rs[i + e] = rr; z[i + e] = zz; gi[i + e] = nn; gh[i + e] = hn; }
        }
      }
      if (c.valid) {
        if (col + 8 <= p.D && (p.D & 3) == 0) {
          *reinterpret_cast<float4*>(hnext + col) = make_float4(r[0], r[1], r[2], r[3]);
          *reinterpret_cast<float4*>(hnext + col + 4) = make_float4(r[4], r[5], r[6], r[7]);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (col + i < p.D) hnext[col + i] = r[i];
        }
      }
      if (kSave) {
        uint8_t* g = c.sblk + (col / 8) * kStashChunk;
        st_chunk_g(g + p.sl.gr, rs);
        st_chunk_g(g + p.sl.gz, z);
        st_chunk_g(g + p.sl.gn, gi);
        st_chunk_g(g + p.sl.ghn, gh);
        if (c.sblk_next) st_chunk_g(c.sblk_next + p.sl.feat + (col / 8) * kStashChunk, r);   // h_{t+1}: next step's actor input
      }
      st_chunk(c.smem_base + jb.dst_off + (col / 8) * kPSF + c.row * 16, r);
      if (k + 1 < kGruSubsF) {
#pragma unroll
        for (int i = 0; i < 8; ++i) hp[i] = hn_[i];
      }
    }
  }
}

// Gaussian head: this lane half owns latent columns hf*Sp/2 + [0, Sp/2); TMEM per lane half [mean | raw]
__device__ __forceinline__ void epsf_prefetch(const EpiCtxF& c, float (&e)[8]) {
  const TcParams& p = *c.p;
  const float* eps = p.noise + ((long long)c.t * p.B + c.b) * p.S;
  const int col = c.hf * (p.Sp / 2) + c.q * 8;
#pragma unroll
  for (int i = 0; i < 8; ++i) e[i] = (c.valid && c.q * 8 < p.Sp / 2 && col + i < p.S) ? eps[col + i] : 0.f;
}
__device__ __forceinline__ void epif_prior2_copyout(const EpiCtxF& c);
template <bool kCopyOut, bool kSave>
__device__ __forceinline__ void epif_prior2(const EpiCtxF& c, const JobDesc& jb, const float (&e0)[8]) {
  const TcParams& p = *c.p;
  const uint32_t b_m = jb.p_off, b_s = b_m + p.Sp;
  const long long T1 = p.H + 1;
  const long long o = (c.b * T1 + c.t + 1) * p.S;
  const float* eps = p.noise + ((long long)c.t * p.B + c.b) * p.S;
  const int hw = p.Sp / 2;
  for (int sub = c.q; sub * 8 < hw; sub += kQ) {
    const int col = c.hf * hw + sub * 8;
    float m[8], s[8], sd[8], st[8];
    tmem_ld8(c.tmem_lane + jb.tmem_col + sub * 8, m);
    tmem_ld8(c.tmem_lane + jb.tmem_col + hw + sub * 8, s);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 8; i += 4) {
      const float4 bm = ldc4(b_m + col + i), bs = ldc4(b_s + col + i);
      const float bmv[4] = {bm.x, bm.y, bm.z, bm.w}, bsv[4] = {bs.x, bs.y, bs.z, bs.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float mean = m[i + e] + bmv[e];
        const float raw = s[i + e] + bsv[e];
        const float sp = fmaxf(raw, 0.f) + __logf(1.f + ex2_fast(-fabsf(raw) * 1.4426950408889634f));
        const float ee = (sub == c.q) ? e0[i + e] : ((c.valid && col + i + e < p.S) ? eps[col + i + e] : 0.f);
        m[i + e] = mean;
        sd[i + e] = sp + 0.1f;
        st[i + e] = (col + i + e < p.S) ? fmaf(sd[i + e], ee, mean) : 0.f;
      }
    }
    if (p.stage_out) {
      const uint32_t rowb = c.smem_base + p.off_x + (uint32_t)c.row * (uint32_t)p.Sp * 4u;
      const uint32_t fstride = (uint32_t)kRowsF * (uint32_t)p.Sp * 4u;
      const int g = col / 8;
      const uint32_t c0 = (uint32_t)((2 * g) ^ (c.row & 7)) * 16u, c1 = (uint32_t)((2 * g + 1) ^ (c.row & 7)) * 16u;
      asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(rowb + c0), "f"(m[0]), "f"(m[1]), "f"(m[2]), "f"(m[3]) : "memory");
      asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(rowb + c1), "f"(m[4]), "f"(m[5]), "f"(m[6]), "f"(m[7]) : "memory");
      asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(rowb + fstride + c0), "f"(sd[0]), "f"(sd[1]), "f"(sd[2]), "f"(sd[3]) : "memory");
      asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(rowb + fstride + c1), "f"(sd[4]), "f"(sd[5]), "f"(sd[6]), "f"(sd[7]) : "memory");
      asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(rowb + 2 * fstride + c0), "f"(st[0]), "f"(st[1]), "f"(st[2]), "f"(st[3]) : "memory");
      asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(rowb + 2 * fstride + c1), "f"(st[4]), "f"(st[5]), "f"(st[6]), "f"(st[7]) : "memory");
    } else if (c.valid) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (col + i < p.S) {
          p.mean_seq[o + col + i] = m[i];
          p.std_seq[o + col + i] = sd[i];
          p.stoch_seq[o + col + i] = st[i];
        }
    }
    st_chunk(c.smem_base + jb.dst_off + (col / 8) * kPSF + c.row * 16, st);
    if (kSave && c.sblk_next) st_chunk_g(c.sblk_next + p.sl.feat + (p.Dp / 8 + col / 8) * kStashChunk, st);   // s_{t+1}
  }
  // h' (X[h2..]) -> P_h for the next step's GRU.  Independent of this job's accumulators: done by the warps that have
  // no latent columns here (q >= active) while the others do the head's math; with every warp active, by all.
  {
    const int active = min(kQ, (hw + 7) / 8);     // warps per lane quarter that own a sub-group of latent columns
    const bool all = active >= kQ;
    if (all || c.q >= active) {
      const int nhelp = all ? 2 * kQ : 2 * (kQ - active);
      const int idx = all ? c.sidx : c.hf * (kQ - active) + (c.q - active);
      for (int k8 = idx; k8 < p.Dp / 8; k8 += nhelp) {
        uint32_t a, b, cc, d;
        const uint32_t src = c.smem_base + p.off_h2 + k8 * kPSF + c.row * 16;
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(cc), "=r"(d) : "r"(src) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(c.smem_base + p.off_h + k8 * kPSF + c.row * 16),
                     "r"(a), "r"(b), "r"(cc), "r"(d)
                     : "memory");
      }
    }
  }
  if (kCopyOut) epif_prior2_copyout(c);
}

// Coalesced copy-out of the rows staged by epif_prior2 (mean / std / stoch of this step).  Mode 3 runs it AFTER the
// job has been signalled done: the next job only needs P_s, which is final before this.  The staging region (head of
// X) is next written by these same warps (LayerNorm scratch of the following epilogue); the closing barrier keeps a
// fast warp from getting there while a slow one is still reading.
__device__ __forceinline__ void epif_prior2_copyout(const EpiCtxF& c) {
  const TcParams& p = *c.p;
  const long long T1 = p.H + 1;
  if (p.stage_out) {
    asm volatile("bar.sync 5, %0;" ::"r"(kEpiThreads) : "memory");
    const int warp = c.q * 4 + c.quarter;
    const int lane = c.row & 31;
    const long long tile_b0 = c.b - c.row;
    const uint32_t fstride = (uint32_t)kRowsF * (uint32_t)p.Sp * 4u;
    const uint32_t xbase = c.smem_base + p.off_x;
    for (int row = warp; row < kRowsF; row += kEpiThreads / 32) {
      const long long bb = tile_b0 + row;
      if (bb >= p.B) break;
      const long long off = (bb * T1 + c.t + 1) * p.S;
      const uint32_t rb = xbase + (uint32_t)row * (uint32_t)p.Sp * 4u;
      for (int col = lane; col < p.S; col += 32) {
        const uint32_t a0 = rb + (uint32_t)(((col >> 2) ^ (row & 7)) * 16 + (col & 3) * 4);
        float v0, v1, v2;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v0) : "r"(a0));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v1) : "r"(a0 + fstride));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v2) : "r"(a0 + 2 * fstride));
        p.mean_seq[off + col] = v0;
        p.std_seq[off + col] = v1;
        p.stoch_seq[off + col] = v2;
      }
    }
    asm volatile("bar.sync 5, %0;" ::"r"(kEpiThreads) : "memory");
  }
}

// + one warp: its lane 0 (leader CTA) turns the ring's `full` mbarriers into a plain shared-memory counter.  A
// try_wait on an already-complete mbarrier costs the MMA-issuing lane ~130 cycles per stage that nothing can hide (the
// tcgen05 queue is ~1 deep: tools/micro/stage_overhead.cu, 515 -> 644 cycles per 8-MMA stage); a shared-memory load of
// the counter, re-read only when the cached value runs out, costs a fraction of that.
constexpr int kThreadsFold = kThreads + 32;
constexpr int kReadyWarp = kMmaWarp + 1;
template <bool kSave>
__global__ void __launch_bounds__(kThreadsFold, 1) tc_imagine_fold_kernel(const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t kSlot = p.slot_bytes;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t ring = smem_base + p.ring_off;
  const uint32_t bars = smem_base + p.bar_off;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 128u + 8u * s; };
  auto peer_full_bar = [&](int s) { return bars + 256u + 8u * s; };
  auto accf_bar = [&](uint32_t j) { return bars + 384u + 8u * (j & 1u); };
  auto acce_bar = [&](uint32_t j) { return bars + 400u + 8u * (j & 1u); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + p.bar_off + 416);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int unit = (int)(blockIdx.x >> 1), nunits = (int)(gridDim.x >> 1);
  const int ngroups = p.ntiles;   // one group = one 128-row tile = 2 x 64 rows

  // per-stage MMA records (see the non-folded kernel); A panels have 64 rows -> LBO = 1024 B
  const uint4* stab = reinterpret_cast<const uint4*>(smem + p.bar_off + kBarBytes);
  for (int i = threadIdx.x; i < p.nstages; i += kThreadsFold) {
    const uint4 raw = c_stages[i];
    const StageDesc st = *reinterpret_cast<const StageDesc*>(&raw);
    uint4 rec;
    rec.x = (uint32_t)make_smem_desc(smem_base + (uint32_t)st.a_off16 * 16u, kPSF, 128u);
    rec.y = make_idesc_bf16(128, st.n);
    rec.z = ((uint32_t)(st.n / 2) << 16) | ((uint32_t)st.nk16 << 8) | st.flags;
    rec.w = (uint32_t)st.tmem_col | ((uint32_t)st.dep << 16);
    reinterpret_cast<uint4*>(smem + p.bar_off + kBarBytes)[i] = rec;
  }
  volatile uint32_t* ready_ctr = reinterpret_cast<volatile uint32_t*>(smem + p.bar_off + 424);   // stages landed so far
  if (threadIdx.x == 0) {
    *ready_ctr = 0u;
    for (int s = 0; s < p.nslots; ++s) {
      // leader: a stage is "full" when its own half has landed (producer arrive + tx bytes) AND the peer
      // has forwarded that its half landed -> two arrivals on ONE barrier, one wait per stage
      mbar_init(full_bar(s), rank == 0 ? 2 : 1);
      mbar_init(empty_bar(s), 1);
      mbar_init(peer_full_bar(s), 1);
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
      // Stage sequence of a tile (all three role loops walk exactly this): the ST_PRE stages once, then per step
      // every stage - except that the last step skips the ST_PRE stages (they belong to a step that does not exist).
      uint32_t slot = 0, phase = 0;
      auto issue = [&](int s, const uint4& raw) {
        const uint32_t bytes = ((raw.y & 0xFFFFu) * 16u) / 2u;
        mbar_wait(empty_bar(slot), phase ^ 1u);
        mbar_arrive_expect_tx(full_bar(slot), bytes);
        bulk_g2s(ring + slot * kSlot, p.stage_data + (size_t)raw.x * 16u + (size_t)rank * bytes, bytes, full_bar(slot));
        if (++slot == (uint32_t)p.nslots) { slot = 0; phase ^= 1u; }
      };
      for (int grp = unit; grp < ngroups; grp += nunits) {
        for (int s = 0; s < p.nstages; ++s) {
          const uint4 raw = c_stages[s];
          if ((raw.y >> 24) & ST_PRE) issue(s, raw);
        }
        for (int t = 0; t < p.H; ++t)
          for (int s = 0; s < p.nstages; ++s) {
            const uint4 raw = c_stages[s];
            if (((raw.y >> 24) & ST_PRE) && t == p.H - 1) continue;
            issue(s, raw);
          }
      }
    }
  } else if (warp == kMmaWarp) {
    if (lane == 0 && rank != 0) {
      uint32_t slot = 0, phase = 0;
      int npre = 0;
      for (int s = 0; s < p.nstages; ++s) npre += ((c_stages[s].y >> 24) & ST_PRE) ? 1 : 0;
      for (int grp = unit; grp < ngroups; grp += nunits) {
        const int total = npre + p.H * p.nstages - npre;   // prologue + H steps - the last step's skipped stages
        for (int i = 0; i < total; ++i) {
          mbar_wait(full_bar(slot), phase);
          mbar_arrive_remote(mapa(full_bar(slot), 0));
          if (++slot == (uint32_t)p.nslots) { slot = 0; phase ^= 1u; }
        }
      }
    } else if (lane == 0) {
      uint32_t slot = 0, job = 0;
      uint32_t g = 0, ready_seen = 0;   // stage sequence number; last value read from the ready counter
      const uint32_t bslot0 = (ring >> 4) & 0x3FFFu, bslot_step = kSlot >> 4;
      uint32_t bslot = bslot0;          // descriptor address field of the current ring slot
      // issue one stage's MMAs and release its slot
#ifdef WMRL_TC_STAGE_TRACE   // per-stage timing of one job (tools/trace_tc.py); off in production: the lane is latency-critical
      long long* stage_trace = nullptr;   // when set: [2*i] = wait for the stage's weights, [2*i+1] = issue + commit
#endif
      auto run_stage = [&](const uint4& rec, uint32_t flags) {
#ifdef WMRL_TC_STAGE_TRACE
        const long long w0 = stage_trace ? clock64() : 0;
#endif
        if (g >= ready_seen) {
          do { ready_seen = *ready_ctr; } while (ready_seen <= g);   // published by the kReadyWarp lane below
        }
        ++g;
#ifdef WMRL_TC_STAGE_TRACE
        const long long w1 = stage_trace ? clock64() : 0;
#endif
        tc_fence_after();
        // (mode 3 places every accumulator statically: no ST_DBUF here)
        const uint32_t tm = tmem_base + (rec.w & 0xFFFFu);
        const uint32_t idesc = rec.y;
        uint32_t alo = rec.x;
        uint32_t blo = bslot | (rec.z & 0xFFFF0000u);       // slot address field | LBO field (N/2 rows x 16 B)
        const uint32_t bstep = (rec.z >> 16) * 2u;
        constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);
        uint32_t acc = (flags & ST_FIRST_ACC) ? 0u : 1u;
        const uint32_t nk = (rec.z >> 8) & 0xFFu;
        constexpr uint32_t kAStep = (2u * kPSF) >> 4;
        uint32_t i = 0;
        for (; i + 4 <= nk; i += 4) {       // 4 MMAs per trip: the lane must stay ahead of a 64-cycle MMA
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
#ifdef WMRL_TC_STAGE_TRACE
        if (stage_trace) { stage_trace[0] = w1 - w0; stage_trace[1] = clock64() - w1; stage_trace += 2; }
#endif
      };
      for (int grp = unit; grp < ngroups; grp += nunits) {
        mbar_arrive(accf_bar(job));
        mbar_arrive_remote(mapa(accf_bar(job), 1));
        ++job;
        const bool tr = p.trace && blockIdx.x == 0 && grp == unit;
        // prologue: the ST_PRE stages (first actor layer over the h panel) need the INIT job's panels
        {
          bool waited = false;
          for (int s = 0; s < p.nstages; ++s) {
            const uint4 rec = stab[s];
            const uint32_t flags = rec.z & 0xFFu;
            if (!(flags & ST_PRE)) continue;
            if (!waited) {
              const uint32_t dj = job - 1u;
              mbar_wait(acce_bar(dj), (dj >> 1) & 1u);
              waited = true;
            }
            run_stage(rec, flags);
            bslot += bslot_step;
            if (++slot == (uint32_t)p.nslots) { slot = 0; bslot = bslot0; }
          }
        }
        for (int t = 0; t < p.H; ++t) {
          int jidx = -1;
          uint4 rec_next = stab[0];
          for (int s = 0; s < p.nstages; ++s) {
            const uint4 rec = rec_next;
            if (s + 1 < p.nstages) rec_next = stab[s + 1];   // the next record's LDS latency hides under this stage
            const uint32_t flags = rec.z & 0xFFu;
            if ((flags & ST_PRE) && t == p.H - 1) continue;
            if (flags & ST_FIRST_OF_JOB) {
              ++jidx;
              const uint32_t dep = rec.w >> 16;
#ifdef WMRL_TC_STAGE_TRACE
              const long long d0 = tr ? clock64() : 0;
#endif
              if (dep) {
                const uint32_t dj = job - dep;
                mbar_wait(acce_bar(dj), (dj >> 1) & 1u);
              }
              if (tr && t < 4) p.trace[((t * p.njobs) + jidx) * 4 + 0] = clock64();
#ifdef WMRL_TC_STAGE_TRACE
              if (tr && t == 1) p.trace[760 + jidx] = clock64() - d0;          // dependency wait of each job
              #ifdef WMRL_TC_STAGE_TRACE_JOB
              stage_trace = (tr && t == 1 && jidx == WMRL_TC_STAGE_TRACE_JOB) ? p.trace + 730 : nullptr;
#else
              stage_trace = (tr && t == 1 && jidx == p.njobs - 3) ? p.trace + 730 : nullptr;   // gru1's stages
#endif
#endif
            }
            run_stage(rec, flags);
            if (flags & ST_LAST_OF_JOB) {
              tc_commit2(accf_bar(job), 3);
              if (tr && t < 4) p.trace[((t * p.njobs) + jidx) * 4 + 1] = clock64();
              ++job;
            }
            bslot += bslot_step;
            if (++slot == (uint32_t)p.nslots) { slot = 0; bslot = bslot0; }
          }
        }
      }
    }
  } else if (warp == kReadyWarp) {
    if (lane == 0 && rank == 0) {
      // every stage instance in ring order: wait until both halves have landed (own copy + the peer's forwarded
      // arrival), then publish the count
      uint32_t slot = 0, phase = 0, n = 0;
      for (int grp = unit; grp < ngroups; grp += nunits) {
        const int total = p.H * p.nstages;   // prologue + H steps - the stages the last step skips
        for (int i = 0; i < total; ++i) {
          mbar_wait(full_bar(slot), phase);
          *ready_ctr = ++n;
          if (++slot == (uint32_t)p.nslots) { slot = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    EpiCtxF c;
    c.p = &p;
    c.smem_base = smem_base;
    c.quarter = warp & 3;
    c.hf = c.quarter >> 1;
    c.q = warp >> 2;
    c.sidx = c.hf * kQ + c.q;
    c.row = (c.quarter & 1) * 32 + lane;
    c.tmem_lane = tmem_base + ((uint32_t)(c.quarter * 32) << 16);
    uint32_t job = 0;
#define WMRL_JOB_DONE_F(job)                                                            \
  do {                                                                                  \
    __syncwarp();                                                                       \
    if (lane == 0) {                                                                    \
      if (rank != 0) mbar_arrive_remote(mapa(acce_bar(job), 0));                        \
      else mbar_arrive(acce_bar(job));                                                  \
    }                                                                                   \
  } while (0)
    for (int grp = unit; grp < ngroups; grp += nunits) {
      c.b = (long long)grp * kTileM + (long long)rank * kRowsF + c.row;
      c.valid = c.b < p.B;
      c.t = 0;
      // training: stash block of (tile, CTA, step), see StashLayout; the pointer carries this thread's row offset
      uint8_t* const stash_tile = kSave ? p.stash + ((size_t)grp * 2 + rank) * (size_t)p.H * p.sl.blk_f + c.row * 16 : nullptr;
      c.sblk = stash_tile;
      c.sblk_next = nullptr;
      mbar_wait_backoff(accf_bar(job), (job >> 1) & 1u);
      epif_init<kSave>(c);
      fence_proxy_async_smem();
      WMRL_JOB_DONE_F(job);
      ++job;
      const bool tr = p.trace && blockIdx.x == 0 && grp == unit && threadIdx.x == 0;
      for (int t = 0; t < p.H; ++t) {
        c.t = t;
        if (kSave) {
          c.sblk = stash_tile + (size_t)t * p.sl.blk_f;
          c.sblk_next = (t + 1 < p.H) ? c.sblk + p.sl.blk_f : nullptr;
        }
        for (int jb_i = 0; jb_i < p.njobs; ++jb_i) {
          const JobDesc jb = c_jobs[jb_i];
          long long t_wake = 0;
          if (jb.type == JOB_GRU) {
            float hp[8];
            gruf_load_h(c, jb, c.q, hp);
            mbar_wait_backoff(accf_bar(job), (job >> 1) & 1u);
            tc_fence_after();
            if (tr) t_wake = clock64();
            epif_gru<kSave>(c, jb, c.tmem_lane + jb.tmem_col + (jb.dbuf ? (job & 1u) * 256u : 0u), hp);
          } else if (jb.type == JOB_PRIOR2_CONT) {
            float e0[8];
            epsf_prefetch(c, e0);
            mbar_wait_backoff(accf_bar(job), (job >> 1) & 1u);
            tc_fence_after();
            if (tr) t_wake = clock64();
            epif_prior2<false, kSave>(c, jb, e0);
          } else {
            mbar_wait_backoff(accf_bar(job), (job >> 1) & 1u);
            tc_fence_after();
            if (tr) t_wake = clock64();
            if (jb.type == JOB_ACTOR_LN) { if (jb.aux & 1u) epif_actor_ln<false, kSave>(c, jb); else epif_actor_ln<true, kSave>(c, jb); }
            else if (jb.type == JOB_DENSE_ELU) epif_dense_elu<kSave>(c, jb);
            else if (jb.type == JOB_MU) epif_mu(c, jb);
          }
          tc_fence_before();
          fence_proxy_async_smem();
          if (tr && t < 4) {
            p.trace[((t * p.njobs) + jb_i) * 4 + 2] = t_wake;
            p.trace[((t * p.njobs) + jb_i) * 4 + 3] = clock64();
          }
          WMRL_JOB_DONE_F(job);
          ++job;
          if (jb.type == JOB_PRIOR2_CONT) epif_prior2_copyout(c);   // off the critical path: after the job-done signal
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == kMmaWarp) tmem_dealloc2(tmem_base, 512);
}

// ==========================================================================
// Mode 4: folded CTA pairs with a two-tile ping-pong.  Same epilogues as mode 3; the three role loops walk
// the events (step, job, tile) in the order ... (j,A) (j,B) (j+1,A) (j+1,B) ...  While the epilogue warps work
// on (j,A) the MMA lane issues (j,B); (j+1,A) only needs epilogue (j,A).  Tile T owns the A panels at
// T*panel_bytes, TMEM columns [256*T, +256) and its own accumulator barriers.
// ==========================================================================
__global__ void __launch_bounds__(kThreads, 1) tc_imagine_pp_kernel(const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t kSlot = p.slot_bytes;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t ring = smem_base + p.ring_off;
  const uint32_t bars = smem_base + p.bar_off;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 128u + 8u * s; };
  auto accf_bar = [&](int tile, uint32_t j) { return bars + 384u + 16u * tile + 8u * (j & 1u); };
  auto acce_bar = [&](int tile, uint32_t j) { return bars + 416u + 16u * tile + 8u * (j & 1u); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + p.bar_off + 448);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int unit = (int)(blockIdx.x >> 1), nunits = (int)(gridDim.x >> 1);
  const int ngroups = p.ntiles;   // one group = one 128-row tile = 2 x 64 rows
  const int niter = (ngroups > unit) ? ((ngroups - unit + nunits - 1) / nunits + 1) / 2 : 0;   // two groups per trip

  const uint4* stab = reinterpret_cast<const uint4*>(smem + p.bar_off + kBarBytes);
  for (int i = threadIdx.x; i < p.nstages; i += kThreads) {
    const uint4 raw = c_stages[i];
    const StageDesc st = *reinterpret_cast<const StageDesc*>(&raw);
    uint4 rec;
    rec.x = (uint32_t)make_smem_desc(smem_base + (uint32_t)st.a_off16 * 16u, kPSF, 128u);
    rec.y = make_idesc_bf16(128, st.n);
    rec.z = ((uint32_t)(st.n / 2) << 16) | ((uint32_t)st.nk16 << 8) | st.flags;
    rec.w = (uint32_t)st.tmem_col | ((uint32_t)st.dep << 16);
    reinterpret_cast<uint4*>(smem + p.bar_off + kBarBytes)[i] = rec;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.nslots; ++s) {
      mbar_init(full_bar(s), rank == 0 ? 2 : 1);   // own half (arrive + tx) and the peer's forwarded arrival
      mbar_init(empty_bar(s), 1);
    }
    for (int tile = 0; tile < 2; ++tile)
      for (uint32_t j = 0; j < 2; ++j) {
        mbar_init(accf_bar(tile, j), 1);
        mbar_init(acce_bar(tile, j), (kEpiThreads / 32) * 2);
      }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) { tmem_alloc2(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512); tmem_relinquish2(); }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kProducerWarp) {
    // weight stages of a job are streamed twice in a row: once for tile A's MMAs, once for tile B's
    if (lane == 0) {
      uint32_t slot = 0, phase = 0;
      for (int it = 0; it < niter; ++it)
        for (int t = 0; t < p.H; ++t)
          for (int s0 = 0; s0 < p.nstages;) {
            int s_end = s0;
            for (int tile = 0; tile < 2; ++tile) {
              int s = s0;
              bool last;
              do {
                const uint4 raw = c_stages[s];
                const uint32_t bytes = ((raw.y & 0xFFFFu) * 16u) / 2u;
                last = ((raw.y >> 24) & ST_LAST_OF_JOB) != 0;
                mbar_wait(empty_bar(slot), phase ^ 1u);
                mbar_arrive_expect_tx(full_bar(slot), bytes);
                bulk_g2s(ring + slot * kSlot, p.stage_data + (size_t)raw.x * 16u + (size_t)rank * bytes, bytes,
                         full_bar(slot));
                if (++slot == (uint32_t)p.nslots) { slot = 0; phase ^= 1u; }
                ++s;
              } while (!last);
              s_end = s;
            }
            s0 = s_end;
          }
    }
  } else if (warp == kMmaWarp) {
    if (lane == 0 && rank != 0) {
      // peer: forward "my half landed", once per stage instance (2 x nstages per step)
      uint32_t slot = 0, phase = 0;
      for (int it = 0; it < niter; ++it)
        for (int t = 0; t < p.H; ++t)
          for (int s = 0; s < 2 * p.nstages; ++s) {
            mbar_wait(full_bar(slot), phase);
            mbar_arrive_remote(mapa(full_bar(slot), 0));
            if (++slot == (uint32_t)p.nslots) { slot = 0; phase ^= 1u; }
          }
    } else if (lane == 0) {
      uint32_t slot = 0, phase = 0, job = 0;
      const uint32_t tile_a16 = p.panel_bytes >> 4;
      for (int it = 0; it < niter; ++it) {
        for (int tile = 0; tile < 2; ++tile) {   // INIT jobs have no MMA
          mbar_arrive(accf_bar(tile, job));
          mbar_arrive_remote(mapa(accf_bar(tile, job), 1));
        }
        ++job;
        for (int t = 0; t < p.H; ++t)
          for (int s0 = 0; s0 < p.nstages;) {
            int s_end = s0;
            for (int tile = 0; tile < 2; ++tile) {
              int s = s0;
              uint32_t flags;
              do {
                const uint4 rec = stab[s];
                flags = rec.z & 0xFFu;
                if (flags & ST_FIRST_OF_JOB) {
                  const uint32_t dep = rec.w >> 16;
                  if (dep) {
                    const uint32_t dj = job - dep;
                    mbar_wait(acce_bar(tile, dj), (dj >> 1) & 1u);
                  }
                }
                mbar_wait(full_bar(slot), phase);
                tc_fence_after();
                const uint32_t tm = tmem_base + (uint32_t)tile * 256u + (rec.w & 0xFFFFu);
                const uint32_t idesc = rec.y;
                const uint32_t nh = rec.z >> 16;
                uint32_t alo = rec.x + (uint32_t)tile * tile_a16;
                uint32_t blo = (((ring + slot * kSlot) >> 4) & 0x3FFFu) | (nh << 16);
                const uint32_t bstep = nh * 2u;
                constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);
                constexpr uint32_t kAStep = (2u * kPSF) >> 4;
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
                if (flags & ST_LAST_OF_JOB) tc_commit2(accf_bar(tile, job), 3);
                if (++slot == (uint32_t)p.nslots) { slot = 0; phase ^= 1u; }
                ++s;
              } while (!(flags & ST_LAST_OF_JOB));
              s_end = s;
            }
            s0 = s_end;
            ++job;
          }
      }
    }
  } else {
    EpiCtxF c;
    c.p = &p;
    c.quarter = warp & 3;
    c.hf = c.quarter >> 1;
    c.q = warp >> 2;
    c.sidx = c.hf * kQ + c.q;
    c.row = (c.quarter & 1) * 32 + lane;
    uint32_t job = 0;
#define WMRL_JOB_DONE_P(tile, job)                                                      \
  do {                                                                                  \
    __syncwarp();                                                                       \
    if (lane == 0) {                                                                    \
      if (rank != 0) mbar_arrive_remote(mapa(acce_bar(tile, job), 0));                  \
      else mbar_arrive(acce_bar(tile, job));                                            \
    }                                                                                   \
  } while (0)
    // per-tile context (recomputed, never indexed arrays: local memory is an L2 round trip here)
#define WMRL_SET_TILE(tile)                                                                       \
  do {                                                                                            \
    const int grp_ = unit + (2 * it + (tile)) * nunits;                                           \
    c.smem_base = smem_base + (uint32_t)(tile) * p.panel_bytes;                                   \
    c.tmem_lane = tmem_base + (uint32_t)(tile) * 256u + ((uint32_t)(c.quarter * 32) << 16);       \
    c.b = (long long)grp_ * kTileM + (long long)rank * kRowsF + c.row;                            \
    c.valid = grp_ < ngroups && c.b < p.B;                                                        \
  } while (0)
    for (int it = 0; it < niter; ++it) {
      for (int tile = 0; tile < 2; ++tile) {
        WMRL_SET_TILE(tile);
        c.t = 0;
        mbar_wait_backoff(accf_bar(tile, job), (job >> 1) & 1u);
        epif_init<false>(c);
        fence_proxy_async_smem();
        WMRL_JOB_DONE_P(tile, job);
      }
      ++job;
      for (int t = 0; t < p.H; ++t) {
        c.t = t;
        for (int jb_i = 0; jb_i < p.njobs; ++jb_i) {
          const JobDesc jb = c_jobs[jb_i];
          for (int tile = 0; tile < 2; ++tile) {
            WMRL_SET_TILE(tile);
            if (jb.type == JOB_GRU) {
              float hp[8];
              gruf_load_h(c, jb, c.q, hp);
              mbar_wait_backoff(accf_bar(tile, job), (job >> 1) & 1u);
              tc_fence_after();
              epif_gru<false>(c, jb, c.tmem_lane, hp);
            } else if (jb.type == JOB_PRIOR2_CONT) {
              float e0[8];
              epsf_prefetch(c, e0);
              mbar_wait_backoff(accf_bar(tile, job), (job >> 1) & 1u);
              tc_fence_after();
              epif_prior2<true, false>(c, jb, e0);
            } else {
              mbar_wait_backoff(accf_bar(tile, job), (job >> 1) & 1u);
              tc_fence_after();
              if (jb.type == JOB_ACTOR_LN) { if (jb.aux & 1u) epif_actor_ln<false, false>(c, jb); else epif_actor_ln<true, false>(c, jb); }
              else if (jb.type == JOB_DENSE_ELU) epif_dense_elu<false>(c, jb);
              else if (jb.type == JOB_MU) epif_mu(c, jb);
            }
            tc_fence_before();
            fence_proxy_async_smem();
            WMRL_JOB_DONE_P(tile, job);
          }
          ++job;
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == kMmaWarp) tmem_dealloc2(tmem_base, 512);
}

int tc_imagine_fwd(const Dims& d, const wmrl_rssm_params* w, const wmrl_actor_params* aw,
                   const void* packed, const float* deter0, const float* stoch0,
                   const float* noise, float* deter_seq, float* stoch_seq, float* dist_a,
                   float* dist_b, float* actions, int32_t* idx, void* saved, void* ws, cudaStream_t st) {
  (void)w; (void)aw; (void)idx;
  Plan pl = make_plan(d, nullptr, nullptr);
  WMRL_REQUIRE(pl.ok, "tc_imagine_fwd: unsupported shape");
  WMRL_REQUIRE(!saved || (pl.fold && !pl.pp), "tc_imagine_fwd: activations are recorded by the folded-pair kernel only");
  const uint8_t* base = (const uint8_t*)packed;
  TcParams p{};
  p.stage_data = base + pl.off_data;
  p.stages = (const StageDesc*)(base + pl.off_stage_tab);
  p.jobs = (const JobDesc*)(base + pl.off_job_tab);
  p.eparams = (const float*)(base + pl.off_eparams);
  p.deter0 = deter0; p.stoch0 = stoch0; p.noise = noise;
  p.deter_seq = deter_seq; p.stoch_seq = stoch_seq; p.mean_seq = dist_a; p.std_seq = dist_b;
  p.actions = actions;
  p.nstages = (int)pl.stages.size(); p.njobs = (int)pl.jobs.size();
  p.B = d.B; p.H = d.T; p.D = d.D; p.S = d.S; p.A = d.A;
  p.Dp = pl.Dp; p.Sp = pl.Sp; p.Ap = pl.Ap; p.Ha = d.Ha; p.Hdp = pl.Hdp;
  p.inv_action_scale = 1.f / d.action_scale;
  p.off_h = pl.off_h; p.off_s = pl.off_s; p.off_a = pl.off_a; p.off_x = pl.off_x; p.off_h2 = pl.off_h2;
  p.ring_off = pl.ring_off; p.bar_off = pl.bar_off; p.nslots = pl.nslots;
  p.slot_bytes = (uint32_t)(pl.stage_bytes / pl.cta);
  p.panel_bytes = pl.panels_bytes;
  p.ntiles = (d.B + kTileM - 1) / kTileM;
  p.trace = (d.flags & 4) ? (long long*)ws : nullptr;
  p.stash = (uint8_t*)saved;
  p.sl = pl.sl;
  // swizzle needs >= 8 16-byte chunks per row (Sp >= 32) and the staging must fit below the h' columns
  {
    const uint32_t rows = pl.fold ? 64u : 128u;
    p.stage_out = (pl.Sp >= 32 && pl.Sp % 32 == 0 && 3u * rows * pl.Sp * 4u <= (uint32_t)pl.h2c * 2u * rows) ? 1 : 0;
  }
  // per-device launch state, queried / set once
  static int s_sms[64];
  static unsigned s_smem[64];
  int dev = 0;
  WMRL_CUDA_OK(cudaGetDevice(&dev));
  WMRL_REQUIRE(dev >= 0 && dev < 64, "device ordinal out of range");
  std::lock_guard<std::mutex> lk(g_tc_mutex);
  if (s_sms[dev] == 0) WMRL_CUDA_OK(cudaDeviceGetAttribute(&s_sms[dev], cudaDevAttrMultiProcessorCount, dev));
  if (s_smem[dev] < pl.smem_bytes) {
    WMRL_CUDA_OK(cudaFuncSetAttribute(tc_imagine_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)pl.smem_bytes));
    WMRL_CUDA_OK(cudaFuncSetAttribute(tc_imagine_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)pl.smem_bytes));
    WMRL_CUDA_OK(cudaFuncSetAttribute(tc_imagine_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)pl.smem_bytes));
    WMRL_CUDA_OK(cudaFuncSetAttribute(tc_imagine_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)pl.smem_bytes));
    WMRL_CUDA_OK(cudaFuncSetAttribute(tc_imagine_fold_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)pl.smem_bytes));
    WMRL_CUDA_OK(cudaFuncSetAttribute(tc_imagine_fold_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)pl.smem_bytes));
    WMRL_CUDA_OK(cudaFuncSetAttribute(tc_imagine_pp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)pl.smem_bytes));
    s_smem[dev] = pl.smem_bytes;
  }
  // tables + epilogue parameters -> constant memory (device-to-device from the packed image, stream-ordered
  // before the launch; skipped when this image is already resident on this device)
  ConstState& cs = g_const[dev];
  if (cs.src != packed || cs.gen != tc_pack_generation(packed)) {
    if (cs.last_launch && cs.last_stream != st) WMRL_CUDA_OK(cudaStreamWaitEvent(st, cs.last_launch, 0));
    WMRL_CUDA_OK(cudaMemcpyToSymbolAsync(c_stages, base + pl.off_stage_tab, pl.stages.size() * sizeof(StageDesc), 0,
                                         cudaMemcpyDeviceToDevice, st));
    WMRL_CUDA_OK(cudaMemcpyToSymbolAsync(c_jobs, base + pl.off_job_tab, pl.jobs.size() * sizeof(JobDesc), 0,
                                         cudaMemcpyDeviceToDevice, st));
    WMRL_CUDA_OK(cudaMemcpyToSymbolAsync(c_eparams, base + pl.off_eparams, pl.eparams_floats * sizeof(float), 0,
                                         cudaMemcpyDeviceToDevice, st));
    cs.src = packed;
    cs.gen = tc_pack_generation(packed);
  }
  const int sms = s_sms[dev];
  if (pl.cta == 2) {
    // CTA pairs: one cluster of 2 per pair of 128-row tiles (mode 2) / per 128-row tile (mode 3, folded)
    const int pairs = std::min(pl.pp ? (p.ntiles + 1) / 2 : (pl.fold ? p.ntiles : (p.ntiles + 1) / 2), sms / 2);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = pl.smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (pl.pp) WMRL_CUDA_OK(cudaLaunchKernelEx(&cfg, tc_imagine_pp_kernel, p));
    else if (pl.fold) {
      cfg.blockDim = dim3(kThreadsFold);
      if (p.stash) WMRL_CUDA_OK(cudaLaunchKernelEx(&cfg, tc_imagine_fold_kernel<true>, p));
      else WMRL_CUDA_OK(cudaLaunchKernelEx(&cfg, tc_imagine_fold_kernel<false>, p));
    }
    else if (p.trace) WMRL_CUDA_OK(cudaLaunchKernelEx(&cfg, tc_imagine_kernel<2, true>, p));
    else WMRL_CUDA_OK(cudaLaunchKernelEx(&cfg, tc_imagine_kernel<2, false>, p));
  } else {
    const int grid = std::min(p.ntiles, sms);
    if (p.trace) tc_imagine_kernel<1, true><<<grid, kThreads, pl.smem_bytes, st>>>(p);
    else tc_imagine_kernel<1, false><<<grid, kThreads, pl.smem_bytes, st>>>(p);
  }
  WMRL_CUDA_OK(cudaGetLastError());
  if (!cs.last_launch) WMRL_CUDA_OK(cudaEventCreateWithFlags(&cs.last_launch, cudaEventDisableTiming));
  WMRL_CUDA_OK(cudaEventRecord(cs.last_launch, st));
  cs.last_stream = st;
  return 0;
}

}  // namespace wmrl
