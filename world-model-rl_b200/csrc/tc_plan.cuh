// Host-side plan tables shared by the bf16 tcgen05 kernels: the rollout kernels (tc_imagine.cu) and the
// BPTT kernels (tc_bptt.cu).  STAGES describe what the producer lane loads and the MMA lane issues, JOBS
// what the epilogue warps do once a job's accumulator is complete; PackOps / VecOps describe how the fp32
// torch-layout weights are turned into the packed bf16 image the kernels stream.
#pragma once
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace wmrl {

// --------------------------------------------------------------------------
// tables
// --------------------------------------------------------------------------
constexpr int kTileM = 128;
constexpr int kSlotBytes = 16384;       // ring slot (one weight stage)
constexpr int kMaxSmem = 232448;        // 227 KB opt-in limit on sm_100
constexpr int kQ = 4;                   // epilogue warps per TMEM lane quarter
constexpr int kEpiThreads = 4 * kQ * 32;
constexpr int kThreads = kEpiThreads + 64;   // epilogue warps first; producer and MMA warps LAST: the issue
                                              // arbiter favours high warp ids and these two must never starve
constexpr int kProducerWarp = 4 * kQ, kMmaWarp = 4 * kQ + 1;
constexpr int kDefaultTcMode = 3;
constexpr int kGruChunk = 128;           // max state columns per GRU job
constexpr int kP2Groups = 2;            // Gaussian head: 8-column groups per epilogue warp (S <= 8*kQ*kP2Groups)

// ST_PRE: stage belongs to the NEXT step (issued early, once before the first step, skipped in the last)
enum : uint8_t { ST_FIRST_ACC = 1, ST_LAST_OF_JOB = 2, ST_FIRST_OF_JOB = 4, ST_DBUF = 8, ST_PRE = 16 };
struct __align__(16) StageDesc {
  uint32_t w_off16;   // byte offset / 16 of this stage in the packed stage area
  uint16_t bytes16;   // stage bytes / 16
  uint8_t nk16;       // number of K=16 MMAs in the stage
  uint8_t flags;      // ST_*
  uint16_t a_off16;   // smem byte offset / 16 of the A panel at the stage's first K step
  uint16_t n;         // MMA N
  uint16_t tmem_col;  // accumulator column (plus (job&1)*256 when ST_DBUF)
  uint8_t dep;        // first stage of a job: wait for the epilogue of job (j - dep); 0 = none
  uint8_t pad;
};
static_assert(sizeof(StageDesc) == 16, "StageDesc must be 16 bytes");

enum : uint8_t { JOB_ACTOR_LN = 1, JOB_MU, JOB_DENSE_ELU, JOB_GRU, JOB_PRIOR2_CONT };
struct __align__(16) JobDesc {
  uint8_t type;
  uint8_t dbuf;        // accumulator lives in the (job&1) half of TMEM
  uint16_t n;          // padded accumulator width handled by the epilogue
  uint16_t tmem_col;
  uint16_t col0;       // GRU: first state column of the chunk
  uint32_t dst_off;    // smem byte offset of the destination A panel (column 0 of it)
  uint32_t p_off;      // float offset of this job's epilogue parameters
  uint32_t aux;
  uint32_t st0;        // training: byte offset of the stash field this epilogue records (kNoStash: none)
  uint32_t st1;        // second stash field (LayerNorm jobs: x-hat; BPTT LayerNorm jobs: g_u)
  uint32_t layer;      // actor block index of a LayerNorm job
};
constexpr uint32_t kNoStash = 0xFFFFFFFFu;

// Activation stash of the bf16 training path.  One BLOCK per (128-row tile, CTA of the pair,
// step): every field is a bf16 panel [k/8][64 rows][8] - the layout of the shared-memory A panels, so a warp
// stores / loads 32 rows x 16 B = 512 contiguous bytes and the weight-gradient GEMM reads a field as an
// MN-major tcgen05 operand with no transpose.  Forward block (written by the rollout kernel):
//   feat  [h_t | s_t]                  (Dp + Sp columns)   input of actor layer 0 (its weight gradient)
//   y[l]  ELU(LayerNorm(.)) of block l (Ha)                input of the next layer's weight gradient
//   xh[l] LayerNorm x-hat of block l   (Ha)                LayerNorm / ELU backward
//   xe    embed output x (post ELU)    (Dp)
//   gr gz gn ghn  GRU r, z, n and W_hn h + b_hn (Dp each)
//   hid   prior hidden (post ELU)      (Hdp)
//   rstd  fp32 [La][64]
// Gradient block (written by the BPTT kernel, read by the weight-gradient / column-sum kernels):
//   gp[l] dL/d(pre-LayerNorm) of block l (Ha), gu[l] dL/d(LayerNorm output) (Ha), gmu dL/d(mu pre-tanh) (16)
struct StashLayout {
  uint32_t feat, y[WMRL_MAX_ACTOR_LAYERS], xh[WMRL_MAX_ACTOR_LAYERS], xe, gr, gz, gn, ghn, hid, rstd, blk_f;
  uint32_t gp[WMRL_MAX_ACTOR_LAYERS], gu[WMRL_MAX_ACTOR_LAYERS], gmu, blk_g;
};
constexpr uint32_t kStashChunk = 1024;   // bytes of one 8-column chunk of a 64-row field
static inline StashLayout make_stash_layout(int Dp, int Sp, int Ha, int Hdp, int La) {
  StashLayout s{};
  uint32_t o = 0;
  auto take = [&](int cols) { const uint32_t at = o; o += (uint32_t)(cols / 8) * kStashChunk; return at; };
  s.feat = take(Dp + Sp);
  for (int l = 0; l < La; ++l) s.y[l] = take(Ha);
  for (int l = 0; l < La; ++l) s.xh[l] = take(Ha);
  s.xe = take(Dp);
  s.gr = take(Dp); s.gz = take(Dp); s.gn = take(Dp); s.ghn = take(Dp);
  s.hid = take(Hdp);
  s.rstd = o; o += ((uint32_t)La * 64u * 4u + kStashChunk - 1) / kStashChunk * kStashChunk;
  s.blk_f = o;
  o = 0;
  for (int l = 0; l < La; ++l) s.gp[l] = take(Ha);
  for (int l = 0; l < La; ++l) s.gu[l] = take(Ha);
  s.gmu = take(16);
  s.blk_g = o;
  return s;
}
static_assert(sizeof(JobDesc) == 32, "JobDesc must be 32 bytes");

struct PackOp {
  const float* src;
  int ld;
  int trans;                // 1: element (row n, k) is src[(k0 + k) * ld + n] (a transposed weight: dgrad operands)
  int nseg;                 // the stage's N rows are the concatenation of up to 4 source row ranges
  int row0[4], valid[4], pad[4];
  int k0, kvalid, nk8;
  // composite K (nks > 0): the K index space is the concatenation of up to 4 source ranges [kk0, +kkvalid) padded to
  // kkpad each (e.g. the GRU gate blocks [n | r | z] of W_ih^T); this op covers K [kbegin, kbegin + 8 * nk8) of it
  int nks, kbegin;
  int kk0[4], kkvalid[4], kkpad[4];
  int split;   // 2: the stage is stored as two N-halves (one per CTA of a pair), each [k/8][N/2][8]
  unsigned long long dst_off;
  const float* bias;   // optional: K index bias_k of this op holds bias[row] (bias folded into the GEMM through a
  int bias_k;          // constant-one column of the A panel); -1: none
};
struct VecOp {
  const float* src1;
  const float* src2;
  int off, valid, pad;
  float scale;
  unsigned long long dst;   // float offset
};

struct TcParams {
  const uint8_t* stage_data;
  const StageDesc* stages;
  const JobDesc* jobs;
  const float* eparams;
  const float* deter0;
  const float* stoch0;
  const float* noise;
  float* deter_seq;
  float* stoch_seq;
  float* mean_seq;
  float* std_seq;
  float* actions;
  int nstages, njobs;
  int B, H, D, S, A, Dp, Sp, Ap, Ha, Hdp;
  float inv_action_scale;
  uint32_t off_h, off_s, off_a, off_x, off_h2;   // smem byte offsets of the A panels
  uint32_t ring_off, bar_off;
  int nslots, ntiles;
  uint32_t slot_bytes;   // ring slot size per CTA
  uint32_t panel_bytes;  // bytes of one tile's A panels (ping-pong: tile 1 sits at +panel_bytes)
  int stage_out;      // 1: the Gaussian head stages its fp32 rows in the idle X panel and stores coalesced
  uint8_t* stash;     // training (WMRL_FLAG_SAVE_FOR_BWD): forward stash blocks, see StashLayout
  StashLayout sl;
  long long* trace;   // optional timeline (WMRL_FLAG_TRACE): [step][job][4] SM clocks of CTA 0's first tile
};

static inline int ceil16(int x) { return (x + 15) / 16 * 16; }

struct Plan {
  int cta = 1;          // CTAs per group (1 or 2)
  bool fold = false;    // 64 rows per CTA, accumulator columns folded over the two TMEM lane halves
  bool pp = false;      // two tiles per CTA in ping-pong (folded mode only)
  uint32_t ps = 2048;   // bytes between the 8-column chunks of an A panel (rows per CTA * 16)
  int stage_bytes = kSlotBytes;   // max bytes of one weight stage (all CTAs of the group together)
  int Dp, Sp, Ap, Ha, Hdp, h2c, XW;
  uint32_t off_h, off_s, off_a, off_x, off_h2, panels_bytes, ring_off, bar_off, smem_bytes;
  int nslots;
  StashLayout sl{};     // training stash (folded-pair kernel only)
  uint32_t boff[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // BPTT plan: panel offsets (go, gmu, gx, big, carry, red) and bound's eparams offset
  std::vector<StageDesc> stages;
  std::vector<JobDesc> jobs;
  std::vector<PackOp> pack_ops;
  std::vector<VecOp> vec_ops;
  size_t eparams_floats = 0, data_bytes = 0;
  size_t off_stage_tab, off_job_tab, off_packops, off_eparams, off_data, total_bytes;
  bool ok = false;
};

struct RowSeg { int row0, valid, pad; };
struct RowSpec {
  int nseg = 0;
  RowSeg seg[4];
  RowSpec() {}
  RowSpec(int r0, int v, int p) { nseg = 1; seg[0] = {r0, v, p}; }
  RowSpec(int r0, int v0, int p0, int r1, int v1, int p1) {
    nseg = 0;
    add(r0, v0, p0);
    if (p1 > 0) add(r1, v1, p1);
  }
  void add(int r0, int v, int p) { seg[nseg++] = {r0, v < 0 ? 0 : (v > p ? p : v), p}; }
  int N() const { int n = 0; for (int i = 0; i < nseg; ++i) n += seg[i].pad; return n; }
};
struct Seg {
  uint32_t a_off; int kpad; const float* src; int ld; int k0; int kvalid; const float* bias = nullptr; int bias_k = -1;
  int trans = 0;
  int nks = 0;           // composite K: kpad must equal the sum of kkpad
  int kk0[4] = {0, 0, 0, 0}, kkvalid[4] = {0, 0, 0, 0}, kkpad[4] = {0, 0, 0, 0};
  void add_k(int k0_, int valid_, int pad_) { kk0[nks] = k0_; kkvalid[nks] = valid_; kkpad[nks] = pad_; ++nks; }
};

static void add_acc_group(Plan& pl, const RowSpec& rs, const std::vector<Seg>& segs, int tmem_col,
                          bool dbuf, bool first_of_job, int dep, bool last_of_job, bool onto_existing = false,
                          uint8_t extra_flags = 0) {
  const int N = rs.N();
  int per = pl.stage_bytes / (N * 32);
  if (per < 1) per = 1;
  if (per > 64) per = 64;
  bool first_acc = !onto_existing;   // onto_existing: the accumulator already holds an earlier group's partial sums
  bool first_stage = true;
  for (size_t si = 0; si < segs.size(); ++si) {
    const Seg& sg = segs[si];
    const int total = sg.kpad / 16;
    for (int done = 0; done < total;) {
      const int nk = std::min(per, total - done);
      StageDesc st{};
      st.w_off16 = (uint32_t)(pl.data_bytes / 16);
      st.bytes16 = (uint16_t)(nk * N * 32 / 16);
      st.nk16 = (uint8_t)nk;
      st.flags = extra_flags;
      if (first_acc) st.flags |= ST_FIRST_ACC;
      if (first_of_job && first_stage) { st.flags |= ST_FIRST_OF_JOB; st.dep = (uint8_t)dep; }
      if (dbuf) st.flags |= ST_DBUF;
      const bool last = (si + 1 == segs.size()) && (done + nk == total);
      if (last && last_of_job) st.flags |= ST_LAST_OF_JOB;
      st.a_off16 = (uint16_t)((sg.a_off + (uint32_t)done * 2u * pl.ps) / 16);
      st.n = (uint16_t)N;
      st.tmem_col = (uint16_t)tmem_col;
      pl.stages.push_back(st);
      PackOp op{};
      op.src = sg.src; op.ld = sg.ld; op.trans = sg.trans;
      op.nseg = rs.nseg;
      for (int q = 0; q < rs.nseg; ++q) { op.row0[q] = rs.seg[q].row0; op.valid[q] = rs.seg[q].valid; op.pad[q] = rs.seg[q].pad; }
      op.k0 = sg.k0 + done * 16;
      op.kvalid = std::max(0, sg.kvalid - done * 16);
      op.nk8 = nk * 2;
      op.nks = sg.nks; op.kbegin = done * 16;
      for (int q = 0; q < 4; ++q) { op.kk0[q] = sg.kk0[q]; op.kkvalid[q] = sg.kkvalid[q]; op.kkpad[q] = sg.kkpad[q]; }
      op.split = pl.cta;
      op.dst_off = pl.data_bytes;
      op.bias = nullptr; op.bias_k = -1;
      if (sg.bias_k >= done * 16 && sg.bias_k < (done + nk) * 16) { op.bias = sg.bias; op.bias_k = sg.bias_k - done * 16; }
      pl.pack_ops.push_back(op);
      pl.data_bytes += (size_t)nk * N * 32;
      done += nk;
      first_acc = false;
      first_stage = false;
    }
  }
}

static uint32_t add_vec(Plan& pl, const float* s1, const float* s2, int off, int valid, int pad,
                        float scale = 1.f) {
  VecOp v{s1, s2, off, valid, pad, scale, (unsigned long long)pl.eparams_floats};
  pl.vec_ops.push_back(v);
  const uint32_t at = (uint32_t)pl.eparams_floats;
  pl.eparams_floats += pad;
  return at;
}

// BPTT plan (tc_bptt.cu): transposed weight stages + backward jobs; ok == false when the shape has no bf16 training path
Plan tc_make_bwd_plan(const Dims& d, const wmrl_rssm_params* w, const wmrl_actor_params* aw);
size_t tc_bwd_image_offset(const Dims& d);

}  // namespace wmrl
