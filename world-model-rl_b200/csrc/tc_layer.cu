// bf16 tensor-core path of RSSMBase.imagine_rollout (rssm.py:123-140) for the shapes whose latent tile does not
// fit one SM - the discrete-categorical RSSM (deter 600 / 1024, 32x32 one-hot latent: DESIGN.md 4.7) and every
// continuous shape outside the persistent folded-pair kernel's limits: the LAYER-WISE tcgen05 chain.
//
//   * every activation of a step lives in HBM/L2 as a bf16 PANEL per 128-row batch tile,
//     [tile][k/8][128 rows][8] - exactly the tcgen05 K-major no-swizzle core-matrix layout of an MMA A operand,
//     so a K-slice of a tile is ONE contiguous bulk copy (TMA engine, no tensor map) and the epilogue that
//     produces it stores 512 contiguous bytes per warp.  The whole working set (weights 10-25 MB bf16,
//     activations <= 16 MB per layer at 8 192 rows) stays in the 126 MB L2 across the horizon;
//   * one kernel, tc_layer_kernel, runs every layer of the step: persistent CTAs walk (row tile, column tile)
//     work items; warp roles = TMA producer lane (A K-slices + pre-packed weight K-slices through a 4 x 48 KB
//     ring), one MMA-issuing lane (tcgen05.mma.cta_group::1.kind::f16, M = 128, N <= 256, fp32 accumulators in
//     TMEM, double-buffered across work items when a tile needs <= 256 columns) and 16 epilogue warps;
//   * the layer's elementwise tail is fused into the TMEM -> register epilogue: bias + LayerNorm + ELU
//     (actor.py:23-36), bias + ELU (rssm.py:33-35, 65), tanh-squash action (actor.py:49-50), the GRUCell gates
//     with the fp32 master copy of h in the output tensor (rssm.py:66; the accumulator tile is
//     [i_n | r | z | h_n] of 64 state columns: W_ih x fills [i_n | r | z], W_hh h accumulates onto [r | z] and
//     fills h_n), the Gaussian reparameterised sample (rssm.py:179-181) and the categorical straight-through
//     sample (rssm.py:222-227, state/discrete.py:20-37): per 32-class group logsumexp-normalise -> softmax ->
//     first-index argmax(p / q) entirely in one thread's registers (a thread owns a batch row = a TMEM lane),
//     fp32 logits + one-hot out, uint8-equivalent index out, one-hot bf16 panel for the next step's GEMMs.
//
// The step is a chain of launches on the caller's stream (no host sync): actor blocks, mu, embed, GRU, prior
// hidden, prior head + sample = La + 5 launches per step.
#include <mutex>
#include <vector>

#include "common.cuh"
#include "tc_plan.cuh"
#include "tc_ptx.cuh"

namespace wmrl {
using namespace ptx;

int tc_pack_plan(const Plan& pl, uint8_t* base, cudaStream_t st);

constexpr int kLEpiWarps = 16, kLEpiThreads = kLEpiWarps * 32, kLThreads = kLEpiThreads + 64;
constexpr int kLProducerWarp = kLEpiWarps, kLMmaWarp = kLEpiWarps + 1;
constexpr uint32_t kLSlotBytes = 40960;
constexpr int kLSlots = 4;
constexpr uint32_t kLStageBytes = 32768;    // per epilogue warp 2 KB (16 rows x 128 B): transposes for coalesced global I/O
constexpr uint32_t kLTileChunk = 2048;      // bytes of one 8-column chunk of a 128-row panel tile
constexpr uint32_t kLParamBytes = 8192;     // per work item epilogue parameters (bias / LayerNorm affine / gate biases)
constexpr uint32_t kLScratchBytes = 4096;   // LayerNorm partial statistics: [q][row] x (sum, sq)
constexpr uint32_t kLBarBytes = 256;
constexpr int kLMaxSplit = 4;               // LayerNorm layers of few row tiles: N split over a cluster of <= 4 CTAs
constexpr uint32_t kLXchgBytes = 2 * kLMaxSplit * 128 * 8;   // [item parity][source CTA][row] x (two row statistics)
constexpr uint32_t kLSmemBytes = kLSlots * kLSlotBytes + kLStageBytes + kLParamBytes + kLScratchBytes + kLXchgBytes + kLBarBytes;
constexpr int kGruTile = 64;                // state columns per GRU work item (4 x 64 = 256 TMEM columns)

static_assert((kLSlots & (kLSlots - 1)) == 0, "kLSlots must be a power of two");

enum : int { LE_LN_ELU = 1, LE_BIAS_ELU, LE_MU, LE_GRU, LE_SAMPLE_DISC, LE_SAMPLE_CONT,
              LB_ELU, LB_GRU, LB_CARRY, LB_EMB, LB_LN, LB_POST,     // LB_*: backward epilogues
              // the MLP heads on the chain (reward / done models, critic): LayerNorm + SiLU / ReLU blocks, the output
              // Linear, their backward forms and the input-gradient layer
              LE_LN_SILU, LE_LN_RELU, LE_OUT, LB_LN_SILU, LB_LN_RELU, LB_XGRAD, kLTypeLast = LB_XGRAD };
enum : int { kActElu = 0, kActSilu = 1, kActRelu = 2 };
__host__ __device__ constexpr bool l_is_bwd(int type) {
  return (type >= LB_ELU && type <= LB_POST) || type == LB_LN_SILU || type == LB_LN_RELU || type == LB_XGRAD;
}

struct LSeg {
  const uint8_t* a;        // activation panel, tile 0
  uint32_t a_tile_bytes;   // bytes between row tiles
  uint32_t w_off;          // byte offset of this segment's weights inside a column tile's weight block
  uint16_t k16;            // K = 16 steps
  uint16_t n;              // weight rows (MMA N, split into <= 256 pieces)
  uint16_t tmem_col;
  uint16_t split;          // != 0: in the first K step columns [0,split) accumulate and [split,n) start fresh
  uint16_t fresh;          // the first K step overwrites the accumulator
  uint16_t pad;
};

struct LayerArgs {
  LSeg seg[3];
  int nseg;
  const uint8_t* w;        // packed weights of the layer: column tile j at w + j * w_tile_bytes
  uint32_t w_tile_bytes;
  const float* ep;         // epilogue parameters: column tile j at ep + j * ep_floats
  uint32_t ep_floats;
  int kc;                  // K = 16 steps per ring stage
  int dbuf;                // accumulators double-buffered across work items (tile needs <= 256 TMEM columns)
  int type;                // LE_*
  int n;                   // accumulator columns of one column tile seen by the epilogue
  int NT, MT;              // column tiles, row tiles
  int B, H, t;
  int D, S, C, A, Sp, Dp, Sip, E;
  float inv_action_scale;
  int save;                // forward: record what the BPTT needs (LayerNorm x-hat + rstd, GRU r / z / n / h_n)
  uint8_t* aux[5];         // extra panels (tile 0) an epilogue reads or writes: recorded activations, g_u
  uint32_t aux_tile_bytes;
  float *f0, *f1, *f2, *f3; // extra fp32 tensors (rstd, carry panels, upstream / start-state / input gradients; see the epilogues)
  int csplit;              // > 1: LayerNorm layer whose NT = csplit column tiles run on the CTAs of one cluster and
                           // exchange their row statistics through distributed shared memory (few row tiles)
  int ln_width;            // full LayerNorm width (= n unless csplit > 1)
  int mode, nt_split;      // epilogue variants (LB_EMB: 1 = raw action gradients out; LB_POST: first column tile of the obs part)
  uint8_t* out;            // destination panel (tile 0)
  uint32_t out_tile_bytes;
  int out_cols;            // padded width of the destination panel: chunks beyond it are not stored
  float *deter_seq, *stoch_seq, *dist_a, *dist_b, *actions;
  int32_t* idx;
  const float* noise;
  long long* trace;        // optional (WMRL_FLAG_TRACE): SM clocks of CTA 0, 16 slots per launch (tools/trace_layer.py)
};

// tight spin for the single-lane producer / MMA loops (no printf call in the loop body: its argument set-up bloats the
// issuing lane's instruction stream); a protocol bug still traps
__device__ __forceinline__ void mbar_wait_backoff_free(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity))
    if (++spins > (1u << 26)) __trap();
}

// --------------------------------------------------------------------------
// epilogues.  Thread -> (row = TMEM lane, q = column split among the 4 warps of a lane quarter)
// --------------------------------------------------------------------------
struct LCtx {
  const LayerArgs* p;
  uint32_t prm;        // smem address of this work item's epilogue parameters
  uint32_t scratch;    // smem address of the LayerNorm scratch
  uint32_t xchg, xbar; // cluster exchange area / its mbarrier (csplit > 1)
  uint32_t xphase;     // parity of this CTA's LayerNorm item count
  uint32_t stage;      // smem address of this warp's 2 KB transpose staging
  uint32_t tmem_lane;  // accumulator base of this work item with the warp's lane quarter folded in
  int row, q, quarter, lane;
  int mt, nt;
  long long b;         // global row
  bool valid;
  uint8_t* out_row;    // destination panel of this row tile, offset by row * 16
  long long* probe;    // timeline probes of the first work item (thread 0 of CTA 0 when tracing), else null
};
#define LPROBE(c, k) do { if ((c).probe) (c).probe[k] = clock64(); } while (0)

__device__ __forceinline__ float4 lds4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void l_quarter_sync(int quarter) {
  asm volatile("bar.sync %0, %1;" ::"r"(quarter + 1), "r"(4 * 32) : "memory");
}
__device__ __forceinline__ void l_epi_sync() { asm volatile("bar.sync 5, %0;" ::"r"(kLEpiThreads) : "memory"); }

// Row statistics of a LayerNorm whose columns are split over the CTAs of a cluster: every CTA publishes its (a, b) of
// each row into all CTAs' exchange areas (DSMEM), signals their barriers, and sums what the others published.  Called by
// all epilogue threads after the per-quarter combine (a, b identical in the 4 warps of a quarter).
__device__ __forceinline__ void l_cluster_stats(const LCtx& c, float& a, float& b) {
  const int n = c.p->csplit;
  const uint32_t me = cluster_ctarank();
  const uint32_t slot = c.xchg + ((c.xphase * kLMaxSplit + me) * 128u + (uint32_t)c.row) * 8u;
  if (c.q == 0) {
    for (int r = 0; r < n; ++r) st_cluster_f32x2(mapa(slot, (uint32_t)r), a, b);
    fence_acq_rel_cluster();
    __syncwarp();
    if (c.lane == 0)
      for (int r = 0; r < n; ++r) mbar_arrive_remote_release(mapa(c.xbar, (uint32_t)r));
  }
  mbar_wait_cluster_quiet(c.xbar, c.xphase);
  a = 0.f; b = 0.f;
  for (int r = 0; r < n; ++r) {
    float x, y;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x), "=f"(y)
                 : "r"(c.xchg + ((c.xphase * kLMaxSplit + (uint32_t)r) * 128u + (uint32_t)c.row) * 8u) : "memory");
    a += x; b += y;
  }
}

// activation on zs = z * log2(e) (the LayerNorm affine is pre-scaled): ELU (actor.py:33), SiLU (models.py:12),
// ReLU (trainers/value/critic.py:19)
template <int kAct>
__device__ __forceinline__ float act_scaled(float zs) {
  if (kAct == kActElu) return elu_scaled(zs);
  if (kAct == kActRelu) return fmaxf(zs, 0.f) * 0.6931471805599453f;
  return __fdividef(zs * 0.6931471805599453f, 1.f + ex2_fast(-zs));   // z sigmoid(z)
}
// Linear -> LayerNorm(eps 1e-5) -> ELU | SiLU | ReLU (actor.py:23-36, models.py:3-37, critic.py:4-29); parameters
// [bias | gamma log2e | beta log2e]
template <int kAct>
__device__ __forceinline__ void lepi_ln_act(const LCtx& c) {
  const LayerArgs& p = *c.p;
  const int W = p.n, ngrp = W / 32;   // this column tile (the whole layer unless csplit > 1)
  const uint32_t bias = c.prm, gam = bias + W * 4, bet = gam + W * 4;
  float sum = 0.f, sq = 0.f;
  for (int g = c.q; g < ngrp; g += 4) {
    float v[32];
    tmem_ld32(c.tmem_lane + g * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 b4 = lds4(bias + (g * 32 + i) * 4);
      const float x0 = v[i] + b4.x, x1 = v[i + 1] + b4.y, x2 = v[i + 2] + b4.z, x3 = v[i + 3] + b4.w;
      sum += (x0 + x1) + (x2 + x3);
      sq = fmaf(x0, x0, sq); sq = fmaf(x1, x1, sq); sq = fmaf(x2, x2, sq); sq = fmaf(x3, x3, sq);
    }
  }
  asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(c.scratch + (c.q * 128 + c.row) * 8), "f"(sum), "f"(sq) : "memory");
  l_quarter_sync(c.quarter);
  sum = 0.f; sq = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float a, b;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(a), "=f"(b) : "r"(c.scratch + (k * 128 + c.row) * 8) : "memory");
    sum += a; sq += b;
  }
  if (p.csplit > 1) l_cluster_stats(c, sum, sq);
  const float inv = 1.f / (float)p.ln_width;
  const float mean = sum * inv;
  const float var = fmaxf(sq * inv - mean * mean, 0.f);
  const float rstd = rsqrtf(var + 1e-5f);
  const float nmr = -mean * rstd;
  if (p.save && c.q == 0 && c.nt == 0 && p.f0) p.f0[(long long)c.mt * 128 + c.row] = rstd;
  const size_t ch0 = (size_t)c.nt * (W / 8) * kLTileChunk;   // this tile's first chunk in the row tile's panels
  uint8_t* xrow = p.save ? p.aux[0] + (size_t)c.mt * p.aux_tile_bytes + (size_t)c.row * 16 + ch0 : nullptr;
  uint8_t* const orow = c.out_row + ch0;
  for (int g = c.q; g < ngrp; g += 4) {
    float v[32];
    tmem_ld32(c.tmem_lane + g * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float xh[8];
#pragma unroll
      for (int i = 0; i < 8; i += 4) {
        const int f = g * 32 + k * 8 + i;
        const float4 b4 = lds4(bias + f * 4), g4 = lds4(gam + f * 4), e4 = lds4(bet + f * 4);
        xh[i] = fmaf(v[k * 8 + i] + b4.x, rstd, nmr);
        xh[i + 1] = fmaf(v[k * 8 + i + 1] + b4.y, rstd, nmr);
        xh[i + 2] = fmaf(v[k * 8 + i + 2] + b4.z, rstd, nmr);
        xh[i + 3] = fmaf(v[k * 8 + i + 3] + b4.w, rstd, nmr);
        v[k * 8 + i] = act_scaled<kAct>(fmaf(xh[i], g4.x, e4.x));
        v[k * 8 + i + 1] = act_scaled<kAct>(fmaf(xh[i + 1], g4.y, e4.y));
        v[k * 8 + i + 2] = act_scaled<kAct>(fmaf(xh[i + 2], g4.z, e4.z));
        v[k * 8 + i + 3] = act_scaled<kAct>(fmaf(xh[i + 3], g4.w, e4.w));
      }
      if (p.save) st_chunk_g(xrow + (size_t)(g * 4 + k) * kLTileChunk, xh);
      st_chunk_g(orow + (size_t)(g * 4 + k) * kLTileChunk, v + k * 8);
    }
  }
}

// bias + ELU -> panel (embed rssm.py:65, prior hidden rssm.py:33-35); column tile nt covers [nt*n, +n)
__device__ __forceinline__ void lepi_bias_elu(const LCtx& c) {
  const LayerArgs& p = *c.p;
  for (int g = c.q; g < p.n / 16; g += 4) {
    const int col = c.nt * p.n + g * 16;
    if (col >= p.out_cols) break;
    float v[16];
    tmem_ld16(c.tmem_lane + g * 16, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      const float4 b4 = lds4(c.prm + (g * 16 + i) * 4);
      v[i] = elu_fast(v[i] + b4.x); v[i + 1] = elu_fast(v[i + 1] + b4.y);
      v[i + 2] = elu_fast(v[i + 2] + b4.z); v[i + 3] = elu_fast(v[i + 3] + b4.w);
    }
    st_chunk_g(c.out_row + (size_t)(col / 8) * kLTileChunk, v);
    st_chunk_g(c.out_row + (size_t)(col / 8 + 1) * kLTileChunk, v + 8);
  }
}

// a = tanh(mu / action_scale) * bound (actor.py:49-50) -> actions[B,H,A] + the 16-column action panel
__device__ __forceinline__ void lepi_mu(const LCtx& c) {
  const LayerArgs& p = *c.p;
  if (c.q != 0) return;
  float v[16];
  tmem_ld16(c.tmem_lane, v);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 16; i += 4) {
    const float4 b4 = lds4(c.prm + i * 4), d4 = lds4(c.prm + (16 + i) * 4);
    v[i] = tanh_fast((v[i] + b4.x) * p.inv_action_scale) * d4.x;
    v[i + 1] = tanh_fast((v[i + 1] + b4.y) * p.inv_action_scale) * d4.y;
    v[i + 2] = tanh_fast((v[i + 2] + b4.z) * p.inv_action_scale) * d4.z;
    v[i + 3] = tanh_fast((v[i + 3] + b4.w) * p.inv_action_scale) * d4.w;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    if (i >= p.A) v[i] = 0.f;
    else if (c.valid) p.actions[(c.b * p.H + p.t) * p.A + i] = v[i];
  }
  st_chunk_g(c.out_row, v);
  st_chunk_g(c.out_row + kLTileChunk, v + 8);
}

// output Linear of an MLP head (models.py:22, critic.py:26): out[row][0..O) = acc + bias, fp32, O = p.A <= 16
__device__ __forceinline__ void lepi_out(const LCtx& c) {
  const LayerArgs& p = *c.p;
  if (c.q != 0) return;
  float v[16];
  tmem_ld16(c.tmem_lane, v);
  tmem_ld_wait();
  if (!c.valid) return;
#pragma unroll
  for (int i = 0; i < 16; i += 4) {
    const float4 b4 = lds4(c.prm + i * 4);
    v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i)
    if (i < p.A) p.f0[c.b * p.A + i] = v[i];
}

// GRUCell gates (rssm.py:66) of one 64-column tile: TMEM [i_n | r | z | h_n]; parameters [b_r | b_z | b_in | b_hn]
// (b_r = b_ir + b_hr, b_z likewise).  h_prev is the fp32 master copy deter_seq[:, t]; h' goes to deter_seq[:, t+1]
// and, as bf16, to the other h panel.  Warp q owns the 16-column group q.
struct LGruPre { float hp[16]; };
__device__ __forceinline__ void lgru_prefetch(const LCtx& c, LGruPre& pf) {
  const LayerArgs& p = *c.p;
  const int col = c.nt * kGruTile + c.q * 16;
  const float* hprev = p.deter_seq + (c.b * (p.H + 1) + p.t) * p.D;
  if (c.valid && col + 16 <= p.D && (p.D & 3) == 0) {
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      const float4 h4 = *reinterpret_cast<const float4*>(hprev + col + i);
      pf.hp[i] = h4.x; pf.hp[i + 1] = h4.y; pf.hp[i + 2] = h4.z; pf.hp[i + 3] = h4.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) pf.hp[i] = (c.valid && col + i < p.D) ? hprev[col + i] : 0.f;
  }
}
__device__ __forceinline__ void lepi_gru(const LCtx& c, const LGruPre& pf) {
  const LayerArgs& p = *c.p;
  constexpr int cw = kGruTile;
  const int col0 = c.nt * cw + c.q * 16;
  if (col0 >= p.Dp) return;
  float* hnext = p.deter_seq + (c.b * (p.H + 1) + p.t + 1) * p.D;
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    const int lc = c.q * 16 + hf * 8;   // column inside the tile
    const int col = c.nt * cw + lc;
    float gi[8], r[8], z[8], gh[8], hn[8];
    float sr[8], sz[8], sn[8], sh[8];   // recorded for the BPTT: r, z, n, W_hn h + b_hn
    tmem_ld8(c.tmem_lane + lc, gi);
    tmem_ld8(c.tmem_lane + cw + lc, r);
    tmem_ld8(c.tmem_lane + 2 * cw + lc, z);
    tmem_ld8(c.tmem_lane + 3 * cw + lc, gh);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 8; i += 4) {
      const float4 br = lds4(c.prm + (lc + i) * 4), bz = lds4(c.prm + (cw + lc + i) * 4);
      const float4 bi = lds4(c.prm + (2 * cw + lc + i) * 4), bh = lds4(c.prm + (3 * cw + lc + i) * 4);
      const float brv[4] = {br.x, br.y, br.z, br.w}, bzv[4] = {bz.x, bz.y, bz.z, bz.w};
      const float biv[4] = {bi.x, bi.y, bi.z, bi.w}, bhv[4] = {bh.x, bh.y, bh.z, bh.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float rr = sigmoid_fast(r[i + e] + brv[e]);
        const float zz = sigmoid_fast(z[i + e] + bzv[e]);
        const float nn = tanh_fast(gi[i + e] + biv[e] + rr * (gh[i + e] + bhv[e]));
        const float hprev = pf.hp[hf * 8 + i + e];
        const float h = (col + i + e < p.D) ? fmaf(zz, hprev - nn, nn) : 0.f;   // (1-z) n + z h
        hn[i + e] = h;
        sr[i + e] = rr; sz[i + e] = zz; sn[i + e] = nn; sh[i + e] = gh[i + e] + bhv[e];
      }
    }
    if (c.valid) {
      if (col + 8 <= p.D && (p.D & 3) == 0) {
        *reinterpret_cast<float4*>(hnext + col) = make_float4(hn[0], hn[1], hn[2], hn[3]);
        *reinterpret_cast<float4*>(hnext + col + 4) = make_float4(hn[4], hn[5], hn[6], hn[7]);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (col + i < p.D) hnext[col + i] = hn[i];
      }
    }
    if (col < p.out_cols) st_chunk_g(c.out_row + (size_t)(col / 8) * kLTileChunk, hn);
    if (p.save && col < p.out_cols) {
      const size_t ao = (size_t)c.mt * p.aux_tile_bytes + (size_t)(col / 8) * kLTileChunk + (size_t)c.row * 16;
      st_chunk_g(p.aux[0] + ao, sr); st_chunk_g(p.aux[1] + ao, sz); st_chunk_g(p.aux[2] + ao, sn); st_chunk_g(p.aux[3] + ao, sh);
    }
  }
}

// Categorical straight-through head (rssm.py:222-227, state/discrete.py:20-37).  A column tile holds n / S groups of
// S classes; warp q owns the 32-column blocks q, q+4, ...; a thread holds all classes of its groups in registers.
//   p = Categorical(logits).probs = softmax(logits - logsumexp(logits));  idx = first argmax(p / q), q ~ Exp(1)
// (torch.multinomial's single-draw rule); forward value of the straight-through sample is exactly one_hot(idx).
// Global I/O goes through a per-warp 2 KB staging transpose (16 rows x 128 B, 16-byte chunks XOR-swizzled by the row)
// so that every warp-level LDG / STG touches 4 rows x 128 contiguous bytes instead of 32 different lines: the
// row-per-thread pattern cost 39K cycles per work item in LSU wavefronts alone (tools/trace_layer.py).
__device__ __forceinline__ uint32_t stage_addr(uint32_t base, int r16, int chunk) {
  return base + (uint32_t)r16 * 128u + (uint32_t)((chunk ^ (r16 & 7)) * 16);
}
__device__ __forceinline__ void load_rows32(const LCtx& c, const float* row0, long long ld, int nrows_valid, int ncols_valid,
                                            float (&out)[32]);
template <int S>
__device__ __forceinline__ void lepi_sample_disc(const LCtx& c) {
  const LayerArgs& p = *c.p;
  const long long T1 = p.H + 1;
  const int P = p.C * p.S;
  const int lane = c.lane;
  const long long b0 = c.b - lane;             // global row of lane 0 of this warp
  const int sub = lane >> 3, chunk = lane & 7;  // transposed role: row (4 * i + sub) of a 16-row half, 16-byte chunk
  for (int blk = c.q; blk < p.n / 32; blk += 4) {
    const int col = c.nt * p.n + blk * 32;     // first logit column of this block
    if (col >= p.out_cols) break;
    float l[32];
    LPROBE(c, 0);
    tmem_ld32(c.tmem_lane + blk * 32, l);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 b4 = lds4(c.prm + (blk * 32 + i) * 4);
      l[i] += b4.x; l[i + 1] += b4.y; l[i + 2] += b4.z; l[i + 3] += b4.w;
    }
    LPROBE(c, 1);
    // ---- raw logits out (state/discrete.py:20-37 stores them un-normalised): staging -> coalesced stores
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      if ((lane >> 4) == half) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(stage_addr(c.stage, lane & 15, k)), "f"(l[4 * k]),
                       "f"(l[4 * k + 1]), "f"(l[4 * k + 2]), "f"(l[4 * k + 3]) : "memory");
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r16 = i * 4 + sub;
        const long long br = b0 + half * 16 + r16;
        const float4 v = lds4(stage_addr(c.stage, r16, chunk));
        if (br < p.B && col + chunk * 4 < P)
          *reinterpret_cast<float4*>(p.dist_a + (br * T1 + p.t + 1) * P + col + chunk * 4) = v;
      }
      __syncwarp();
    }
    LPROBE(c, 2);
    // ---- q ~ Exp(1) of this block: coalesced loads (all eight in flight) -> staging -> the row's owner
    float qv[32];
    load_rows32(c, p.noise + ((long long)p.t * p.B + b0) * P + col, P, (int)max(0LL, min(32LL, (long long)p.B - b0)),
                min(32, P - col), qv);
    LPROBE(c, 3);
    // ---- per group: p = softmax(l - logsumexp(l)) = exp(l - max) / sum; idx = first argmax(p / q).  The common
    // factor 1 / sum does not move the argmax and p_j / q_j > p_b / q_b <=> e_j q_b > e_b q_j (all positive), so the
    // draw costs one ex2 per class and no divisions.
    int bi[32 / S];
#pragma unroll
    for (int g = 0; g < 32 / S; ++g) {
      float mx = l[g * S];
#pragma unroll
      for (int j = 1; j < S; ++j) mx = fmaxf(mx, l[g * S + j]);
      const float mxs = mx * 1.4426950408889634f;
      float eb = ex2_fast(fmaf(l[g * S], 1.4426950408889634f, -mxs)), qb = qv[g * S];
      int b = 0;
#pragma unroll
      for (int j = 1; j < S; ++j) {
        const float e = ex2_fast(fmaf(l[g * S + j], 1.4426950408889634f, -mxs));
        const bool better = e * qb > eb * qv[g * S + j];
        eb = better ? e : eb;
        qb = better ? qv[g * S + j] : qb;
        b = better ? j : b;
      }
      bi[g] = b;
      const int grp = (col / S) + g;
      if (c.valid && grp < p.C && p.idx) p.idx[(c.b * p.H + p.t) * p.C + grp] = b;
    }
    LPROBE(c, 4);
    // ---- forward value of the straight-through sample, exactly one_hot(idx): the bf16 panel chunk of the row's
    // owner feeds the next step; the fp32 `stoch` sequence the reference returns is expanded from the indices by
    // tcl_onehot_kernel after the scan (a streaming kernel on all SMs instead of 8 KB per row on this critical path)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float oh[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) oh[e] = (col + k * 8 < P && ((k * 8 + e) % S) == bi[(k * 8 + e) / S]) ? 1.f : 0.f;
      if (col + k * 8 < p.out_cols) st_chunk_g(c.out_row + (size_t)(col / 8 + k) * kLTileChunk, oh);
    }
    LPROBE(c, 5);
  }
}

// ==========================================================================
// BPTT epilogues.  Recorded activations are panels of the same layout, read 16 B per row and chunk;
// every output is ZERO for rows past B (the weight-gradient GEMMs contract over all 128 rows of a tile).
// ==========================================================================
__device__ __forceinline__ uint4 l_ldg16(const void* ptr) { return *reinterpret_cast<const uint4*>(ptr); }
__device__ __forceinline__ void l_unpack8(const uint4& u, float (&v)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
}

// fp32 carry rows live in a panel layout of their own, [tile][col/4][128 rows][4 floats]: a warp's access to four
// columns of its 32 rows is 512 contiguous bytes (the row-major form cost 32 LSU wavefronts per instruction)
__device__ __forceinline__ float* fpanel(float* base, int mt, int Wp, int col, int row) {
  return base + ((size_t)mt * (Wp / 4) + (col >> 2)) * 512 + row * 4;
}
// a 32-row x 32-column block of a row-major fp32 tensor (row r of this warp at row0 + r * ld) -> every thread gets
// its own row's 32 values.  Coalesced: each LDG.128 covers 4 rows x 128 B; transposed through the warp's staging area.
__device__ __forceinline__ void load_rows32(const LCtx& c, const float* row0, long long ld, int nrows_valid, int ncols_valid,
                                            float (&out)[32]) {
  const int sub = c.lane >> 3, chunk = c.lane & 7;
  float4 t[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i * 4 + sub;
    t[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < nrows_valid && chunk * 4 < ncols_valid) {
      const float* src = row0 + (long long)r * ld + chunk * 4;
      if (chunk * 4 + 4 <= ncols_valid && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) t[i] = *reinterpret_cast<const float4*>(src);
      else {
        t[i].x = src[0];
        if (chunk * 4 + 1 < ncols_valid) t[i].y = src[1];
        if (chunk * 4 + 2 < ncols_valid) t[i].z = src[2];
        if (chunk * 4 + 3 < ncols_valid) t[i].w = src[3];
      }
    }
  }
#pragma unroll
  for (int half = 0; half < 2; ++half) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 v = t[half * 4 + i];
      asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(stage_addr(c.stage, i * 4 + sub, chunk)), "f"(v.x), "f"(v.y),
                   "f"(v.z), "f"(v.w) : "memory");
    }
    __syncwarp();
    if ((c.lane >> 4) == half) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float4 v = lds4(stage_addr(c.stage, c.lane & 15, k));
        out[4 * k] = v.x; out[4 * k + 1] = v.y; out[4 * k + 2] = v.z; out[4 * k + 3] = v.w;
      }
    }
    __syncwarp();
  }
}

// g = acc * ELU'(a), a = the recorded post-ELU activation (ELU'(z) = 1 for z > 0, else a + 1) -> panel.
// 32 columns per iteration with every global load in flight before the first use.
__device__ __forceinline__ void lepib_elu(const LCtx& c) {
  const LayerArgs& p = *c.p;
  const uint8_t* arow = p.aux[0] + (size_t)c.mt * p.aux_tile_bytes + (size_t)c.row * 16;
  for (int blk = c.q; blk * 32 < p.n; blk += 4) {
    const int col = c.nt * p.n + blk * 32;
    if (col >= p.out_cols) break;
    float v[32];
    tmem_ld32(c.tmem_lane + blk * 32, v);
    uint4 au[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const bool in = blk * 32 + k * 8 < p.n && col + k * 8 < p.out_cols;
      au[k] = in ? l_ldg16(arow + (size_t)(col / 8 + k) * kLTileChunk) : make_uint4(0, 0, 0, 0);
    }
    tmem_ld_wait();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (blk * 32 + k * 8 < p.n && col + k * 8 < p.out_cols) {
        float a[8], o[8];
        l_unpack8(au[k], a);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = c.valid ? v[k * 8 + i] * ((a[i] > 0.f) ? 1.f : (a[i] + 1.f)) : 0.f;
        st_chunk_g(c.out_row + (size_t)(col / 8 + k) * kLTileChunk, o);
      }
    }
  }
}

// GRUCell backward (h' = (1-z) n + z h, n = tanh(i_n + r h_n)): accumulator = g_hid W1; f0 = the fp32 carry panel
// holding the rest of dL/dh'.  aux[0..3] = recorded r, z, n, h_n; aux[4] = the recorded h_t panel.  Writes the
// gate-gradient panel [g_n | g_r | g_z | g_n r] and leaves the direct term dL/dh' z in the carry panel.
__device__ __forceinline__ void lepib_gru(const LCtx& c) {
  const LayerArgs& p = *c.p;
  const size_t arow = (size_t)c.mt * p.aux_tile_bytes + (size_t)c.row * 16;
  const size_t fstride = (size_t)(p.Dp / 8) * kLTileChunk;
  for (int g = c.q; g * 16 < p.n; g += 4) {
    const int col = c.nt * p.n + g * 16;
    if (col >= p.Dp) break;
    float acc[16];
    tmem_ld16(c.tmem_lane + g * 16, acc);
    uint4 ur[2], uz[2], un[2], uh[2], up[2];
    float4 cr[4];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const size_t ao = arow + (size_t)(col / 8 + k) * kLTileChunk;
      ur[k] = l_ldg16(p.aux[0] + ao); uz[k] = l_ldg16(p.aux[1] + ao); un[k] = l_ldg16(p.aux[2] + ao); uh[k] = l_ldg16(p.aux[3] + ao);
      up[k] = l_ldg16(p.aux[4] + ao);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) cr[k] = *reinterpret_cast<const float4*>(fpanel(p.f0, c.mt, p.Dp, col + 4 * k, c.row));
    tmem_ld_wait();
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      float r[8], z[8], n[8], hn[8], hp[8];
      l_unpack8(ur[k], r); l_unpack8(uz[k], z); l_unpack8(un[k], n); l_unpack8(uh[k], hn); l_unpack8(up[k], hp);
      const float crv[8] = {cr[2 * k].x, cr[2 * k].y, cr[2 * k].z, cr[2 * k].w, cr[2 * k + 1].x, cr[2 * k + 1].y, cr[2 * k + 1].z, cr[2 * k + 1].w};
      float on[8], orr[8], oz[8], oh[8], od[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float gh = c.valid ? acc[k * 8 + i] + crv[i] : 0.f;
        const float gn = gh * (1.f - z[i]);
        const float gz = gh * (hp[i] - n[i]);
        od[i] = gh * z[i];
        const float gnp = gn * (1.f - n[i] * n[i]);
        on[i] = gnp;
        oz[i] = gz * z[i] * (1.f - z[i]);
        oh[i] = gnp * r[i];
        orr[i] = gnp * hn[i] * r[i] * (1.f - r[i]);
      }
      uint8_t* dst = c.out_row + (size_t)(col / 8 + k) * kLTileChunk;
      st_chunk_g(dst, on);
      st_chunk_g(dst + fstride, orr);
      st_chunk_g(dst + 2 * fstride, oz);
      st_chunk_g(dst + 3 * fstride, oh);
      *reinterpret_cast<float4*>(fpanel(p.f0, c.mt, p.Dp, col + 8 * k, c.row)) = make_float4(od[0], od[1], od[2], od[3]);
      *reinterpret_cast<float4*>(fpanel(p.f0, c.mt, p.Dp, col + 8 * k + 4, c.row)) = make_float4(od[4], od[5], od[6], od[7]);
    }
  }
}

// dL/dh_t = acc + dL/dh' z (carry panel) + upstream dL/ddeter[:, t]: carried on, or - at t = 0 - the start-state gradient
__device__ __forceinline__ void lepib_carry(const LCtx& c) {
  const LayerArgs& p = *c.p;
  const long long b0 = c.b - c.lane;
  for (int blk = c.q; blk * 32 < p.n; blk += 4) {
    const int col = c.nt * p.n + blk * 32;
    if (col >= p.Dp) break;
    // two 32-float arrays live at a time (96-register cap: acc + carry + upstream together spilled 260 B): the carried
    // rows are folded into the accumulator before the upstream rows arrive
    float acc[32];
    LPROBE(c, 0);
    tmem_ld32(c.tmem_lane + blk * 32, acc);
    {
      float4 cr[8];
#pragma unroll
      for (int k = 0; k < 8; ++k)
        cr[k] = (col + 4 * k < p.Dp) ? *reinterpret_cast<const float4*>(fpanel(p.f0, c.mt, p.Dp, col + 4 * k, c.row))
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
      LPROBE(c, 1);
      tmem_ld_wait();
      LPROBE(c, 2);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        acc[4 * k] += cr[k].x; acc[4 * k + 1] += cr[k].y; acc[4 * k + 2] += cr[k].z; acc[4 * k + 3] += cr[k].w;
      }
    }
    if (p.f1) {
      float up[32];
      const int nr = (int)max(0LL, min(32LL, (long long)p.B - b0));
      load_rows32(c, p.f1 + (b0 * (p.H + 1) + p.t) * p.D + col, (long long)(p.H + 1) * p.D, nr, min(32, p.D - col), up);
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] += up[i];
    }
    if (!c.valid) {
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = 0.f;
    }
    if (p.t > 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (col + 4 * k < p.Dp && blk * 32 + 4 * k < p.n)
          *reinterpret_cast<float4*>(fpanel(p.f0, c.mt, p.Dp, col + 4 * k, c.row)) =
              make_float4(acc[4 * k], acc[4 * k + 1], acc[4 * k + 2], acc[4 * k + 3]);
    } else if (c.valid && p.f2) {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col + i < p.D && blk * 32 + i < p.n) p.f2[c.b * p.D + col + i] = acc[i];
    }
    LPROBE(c, 3);
  }
}

// [dL/ds_t | dL/da_t] = g_x W_e.  Latent tiles: fp32 carry panel f0 (read by the previous step's sample backward)
// or, at t = 0, the start-state gradient f2 (+ upstream f1[:, 0]).  Last tile: the action columns through the tanh
// squash (actor.py:49-50: a = tanh(mu / scale) bound) -> g_mu panel.
__device__ __forceinline__ void lepib_emb(const LCtx& c) {
  const LayerArgs& p = *c.p;
  const int S_in = p.S * p.C;
  if (c.nt < p.NT - 1) {
    for (int blk = c.q; blk * 32 < p.n; blk += 4) {
      const int col = c.nt * p.n + blk * 32;
      if (col >= p.Sip) break;
      float acc[32];
      tmem_ld32(c.tmem_lane + blk * 32, acc);
      tmem_ld_wait();
      if (p.t > 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (col + 4 * k < p.Sip && blk * 32 + 4 * k < p.n)
            *reinterpret_cast<float4*>(fpanel(p.f0, c.mt, p.Sip, col + 4 * k, c.row)) =
                make_float4(acc[4 * k], acc[4 * k + 1], acc[4 * k + 2], acc[4 * k + 3]);
      } else if (c.valid && p.f2) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (col + i < S_in && blk * 32 + i < p.n) {
            float gv = acc[i];
            if (p.f1) gv += p.f1[(c.b * (p.H + 1)) * S_in + col + i];
            p.f2[c.b * S_in + col + i] = gv;
          }
      }
    }
  } else if (c.q == 0 && p.mode == 1) {
    // filtering: the actions are inputs - their gradient goes out as is (f3 = dL/d actions[B][T][A], index t)
    float acc[16];
    tmem_ld16(c.tmem_lane, acc);
    tmem_ld_wait();
    if (p.f3 && c.valid)
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i < p.A) p.f3[(c.b * (p.H + 1) + p.t) * p.A + i] = acc[i];
  } else if (c.q == 0) {
    float acc[16], gm[16];
    tmem_ld16(c.tmem_lane, acc);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      gm[i] = 0.f;
      if (i < p.A && c.valid) {
        const float bd = __ldg(p.ep + i);
        const float a = p.actions[(c.b * p.H + p.t) * p.A + i];
        const float th = (bd != 0.f) ? a / bd : 0.f;
        gm[i] = acc[i] * bd * (1.f - th * th) * p.inv_action_scale;
      }
    }
    st_chunk_g(c.out_row, gm);
    st_chunk_g(c.out_row + kLTileChunk, gm + 8);
  }
}

// the inverse of load_rows32: every thread hands over its row's 32 values, the warp stores them coalesced
__device__ __forceinline__ void store_rows32(const LCtx& c, float* row0, long long ld, int nrows_valid, int ncols_valid,
                                             const float (&v)[32]) {
  const int sub = c.lane >> 3, chunk = c.lane & 7;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    if ((c.lane >> 4) == half) {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(stage_addr(c.stage, c.lane & 15, k)), "f"(v[4 * k]),
                     "f"(v[4 * k + 1]), "f"(v[4 * k + 2]), "f"(v[4 * k + 3]) : "memory");
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r16 = i * 4 + sub, r = half * 16 + r16;
      const float4 t = lds4(stage_addr(c.stage, r16, chunk));
      if (r < nrows_valid && chunk * 4 < ncols_valid) {
        float* dst = row0 + (long long)r * ld + chunk * 4;
        if (chunk * 4 + 4 <= ncols_valid && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) *reinterpret_cast<float4*>(dst) = t;
        else {
          dst[0] = t.x;
          if (chunk * 4 + 1 < ncols_valid) dst[1] = t.y;
          if (chunk * 4 + 2 < ncols_valid) dst[2] = t.z;
          if (chunk * 4 + 3 < ncols_valid) dst[3] = t.w;
        }
      }
    }
    __syncwarp();
  }
}

// posterior hidden layer dgrad (rssm.py:80): [dL/dh' | dL/de_{t+1}] = g_hidq W_post0.  Column tiles below nt_split
// cover h' and ADD into the fp32 carry panel f0 (the prior path and the carried gradient are already there); the
// tiles from nt_split on cover the observation embedding and go to f3 = dL/d obs_embeds[:, t+1] (row-major fp32).
__device__ __forceinline__ void lepib_post(const LCtx& c) {
  const LayerArgs& p = *c.p;
  const long long b0 = c.b - c.lane;
  const int E = p.E;
  for (int blk = c.q; blk * 32 < p.n; blk += 4) {
    float acc[32];
    if (c.nt < p.nt_split) {
      const int col = c.nt * p.n + blk * 32;
      if (col >= p.Dp) break;
      tmem_ld32(c.tmem_lane + blk * 32, acc);
      float4 cr[8];
#pragma unroll
      for (int k = 0; k < 8; ++k)
        cr[k] = (col + 4 * k < p.Dp && blk * 32 + 4 * k < p.n) ? *reinterpret_cast<const float4*>(fpanel(p.f0, c.mt, p.Dp, col + 4 * k, c.row))
                                                              : make_float4(0.f, 0.f, 0.f, 0.f);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (col + 4 * k < p.Dp && blk * 32 + 4 * k < p.n)
          *reinterpret_cast<float4*>(fpanel(p.f0, c.mt, p.Dp, col + 4 * k, c.row)) =
              c.valid ? make_float4(cr[k].x + acc[4 * k], cr[k].y + acc[4 * k + 1], cr[k].z + acc[4 * k + 2], cr[k].w + acc[4 * k + 3])
                      : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      const int col = (c.nt - p.nt_split) * p.n + blk * 32;
      if (col >= E) break;
      tmem_ld32(c.tmem_lane + blk * 32, acc);
      tmem_ld_wait();
      if (p.f3) {
        const int nr = (int)max(0LL, min(32LL, (long long)p.B - b0));
        store_rows32(c, p.f3 + (b0 * (p.H + 1) + p.t + 1) * E + col, (long long)(p.H + 1) * E, nr, min(min(32, E - col), p.n - blk * 32), acc);
      }
    }
  }
}

// Gaussian head (rssm.py:179-181): TMEM [mean (Sp) | raw (Sp)], parameters [b_mean | b_raw].  Warp q owns the 32-column
// blocks q, q + 4, ...; the noise rows come in and mean / std / sample rows go out through the staging transpose (4 rows x
// 128 contiguous bytes per warp-level access; the row-per-thread scalar form cost ~50 us per step at 65 536 rows).
__device__ __forceinline__ void lepi_sample_cont(const LCtx& c) {
  const LayerArgs& p = *c.p;
  const long long T1 = p.H + 1;
  const long long b0 = c.b - c.lane;
  const int nr = (int)max(0LL, min(32LL, (long long)p.B - b0));
  for (int blk = c.q; blk * 32 < p.Sp; blk += 4) {
    const int col = blk * 32;
    const int nc = min(32, p.S - col);
    const long long row0 = (b0 * T1 + p.t + 1) * p.S + col;
    float ee[32];
    load_rows32(c, p.noise + ((long long)p.t * p.B + b0) * p.S + col, p.S, nr, nc, ee);
    {   // std = softplus(raw) + 0.1 -> out; ee <- std * eps   (two 32-float arrays live at a time: 96-register cap)
      float sd[32];
      tmem_ld32(c.tmem_lane + p.Sp + col, sd);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float4 bs = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col + i < p.Sp) bs = lds4(c.prm + (p.Sp + col + i) * 4);
        const float bsv[4] = {bs.x, bs.y, bs.z, bs.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float raw = sd[i + e] + bsv[e];
          sd[i + e] = fmaxf(raw, 0.f) + __logf(1.f + ex2_fast(-fabsf(raw) * 1.4426950408889634f)) + 0.1f;
          ee[i + e] *= sd[i + e];
        }
      }
      store_rows32(c, p.dist_b + row0, T1 * p.S, nr, nc, sd);
    }
    {   // mean -> out; sample = mean + std * eps -> out and the next step's panel (zeros past S / past B)
      float m[32];
      tmem_ld32(c.tmem_lane + col, m);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float4 bm = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col + i < p.Sp) bm = lds4(c.prm + (col + i) * 4);
        const float bmv[4] = {bm.x, bm.y, bm.z, bm.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          m[i + e] += bmv[e];
          ee[i + e] = (c.valid && col + i + e < p.S) ? ee[i + e] + m[i + e] : 0.f;
        }
      }
      store_rows32(c, p.dist_a + row0, T1 * p.S, nr, nc, m);
    }
    store_rows32(c, p.stoch_seq + row0, T1 * p.S, nr, nc, ee);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (col + k * 8 < p.Sp) st_chunk_g(c.out_row + (size_t)(col / 8 + k) * kLTileChunk, ee + k * 8);
  }
}

// input gradient of an MLP head: accumulator = g_pre_0 W_0 -> fp32 rows of g_x[R][F] (p.D = F), coalesced through the
// staging transpose; p.mode != 0 adds onto what g_x holds (several heads sharing one feature tensor)
__device__ __forceinline__ void lepib_xgrad(const LCtx& c) {
  const LayerArgs& p = *c.p;
  const long long b0 = (long long)c.mt * 128 + c.quarter * 32;
  const int nr = (int)max(0LL, min(32LL, (long long)p.B - b0));
  for (int blk = c.q; blk * 32 < p.n; blk += 4) {
    const int col = c.nt * p.n + blk * 32;
    if (col >= p.D) break;
    float acc[32];
    tmem_ld32(c.tmem_lane + blk * 32, acc);
    tmem_ld_wait();
    const int nc = min(32, p.D - col);
    float* dst = p.f0 + b0 * p.D + col;
    if (p.mode) {
      float old[32];
      load_rows32(c, dst, p.D, nr, nc, old);
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] += old[i];
    }
    store_rows32(c, dst, p.D, nr, nc, acc);
  }
}

// LayerNorm(eps 1e-5, affine) + ELU backward of one actor block (actor.py:23-36) from the recorded x-hat / rstd / y:
//   g_u = g_y ELU'(u) with ELU'(u) = 1 (y > 0) | y + 1 from the recorded post-ELU y (no transcendental op: the exp form
//   was MUFU-bound, 2 x 64K ex2 per work item), g_xh = g_u gamma, g_pre = rstd (g_xh - mean(g_xh) - xh mean(g_xh xh)).
// parameters [gamma log2e | beta log2e | gamma] (only gamma is read); aux[0] = x-hat, aux[2] = y (read), aux[1] = g_u
// panel (written), out = g_pre.  32 columns per iteration, all loads of a block in flight together.
// activation derivative from the recorded tensors: ELU / ReLU from the post-activation y, SiLU from u = x-hat gamma +
// beta (recomputed in the log2e-scaled domain the forward used): sigma (1 + u (1 - sigma))
template <int kAct>
__device__ __forceinline__ float act_grad(float y, float xh, float gs, float bs) {
  if (kAct == kActElu) return fminf(y + 1.f, 1.f);
  if (kAct == kActRelu) return y > 0.f ? 1.f : 0.f;
  const float us = fmaf(xh, gs, bs);
  const float sg = __fdividef(1.f, 1.f + ex2_fast(-us));
  return sg * fmaf(us * 0.6931471805599453f, 1.f - sg, 1.f);
}
template <int kAct>
__device__ __forceinline__ void lepib_ln(const LCtx& c) {
  const LayerArgs& p = *c.p;
  const int W = p.n, ngrp = W / 32;
  const uint32_t gsc = c.prm, bsc = c.prm + W * 4;
  const uint32_t gam = c.prm + 2 * W * 4;
  const size_t ch0 = (size_t)c.nt * (W / 8) * kLTileChunk;   // this column tile's first chunk (csplit > 1: nt > 0)
  const size_t ro = (size_t)c.mt * p.aux_tile_bytes + (size_t)c.row * 16 + ch0;
  const uint8_t* xrow = p.aux[0] + ro;
  const uint8_t* yrow = p.aux[2] + ro;
  uint8_t* gurow = p.aux[1] + ro;
  uint8_t* const orow = c.out_row + ch0;
  const float rstd = c.valid ? p.f0[(long long)c.mt * 128 + c.row] : 0.f;
  float s1 = 0.f, s2 = 0.f;
  LPROBE(c, 0);
  for (int g = c.q; g < ngrp; g += 4) {
    float v[32];
    tmem_ld32(c.tmem_lane + g * 32, v);
    uint4 xu[4], yu[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      xu[k] = l_ldg16(xrow + (size_t)(g * 4 + k) * kLTileChunk);
      yu[k] = (kAct == kActSilu) ? make_uint4(0, 0, 0, 0) : l_ldg16(yrow + (size_t)(g * 4 + k) * kLTileChunk);
    }
    tmem_ld_wait();
    if (g == c.q) LPROBE(c, 1);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float xh[8], y[8], gu[8];
      l_unpack8(xu[k], xh);
      l_unpack8(yu[k], y);
#pragma unroll
      for (int i = 0; i < 8; i += 4) {
        const float4 w4 = lds4(gam + (g * 32 + k * 8 + i) * 4);
        const float gw[4] = {w4.x, w4.y, w4.z, w4.w};
        float gs[4] = {0.f, 0.f, 0.f, 0.f}, bs[4] = {0.f, 0.f, 0.f, 0.f};
        if (kAct == kActSilu) {
          const float4 a4 = lds4(gsc + (g * 32 + k * 8 + i) * 4), b4 = lds4(bsc + (g * 32 + k * 8 + i) * 4);
          gs[0] = a4.x; gs[1] = a4.y; gs[2] = a4.z; gs[3] = a4.w;
          bs[0] = b4.x; bs[1] = b4.y; bs[2] = b4.z; bs[3] = b4.w;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float g_u = c.valid ? v[k * 8 + i + e] * act_grad<kAct>(y[i + e], xh[i + e], gs[e], bs[e]) : 0.f;
          const float g_x = g_u * gw[e];
          gu[i + e] = g_u;
          s1 += g_x;
          s2 = fmaf(g_x, xh[i + e], s2);
        }
      }
      if (p.aux[1]) st_chunk_g(gurow + (size_t)(g * 4 + k) * kLTileChunk, gu);
    }
  }
  LPROBE(c, 2);
  asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(c.scratch + (c.q * 128 + c.row) * 8), "f"(s1), "f"(s2) : "memory");
  l_quarter_sync(c.quarter);
  LPROBE(c, 3);
  s1 = 0.f; s2 = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float a, b;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(a), "=f"(b) : "r"(c.scratch + (k * 128 + c.row) * 8) : "memory");
    s1 += a; s2 += b;
  }
  if (p.csplit > 1) l_cluster_stats(c, s1, s2);
  const float inv = 1.f / (float)p.ln_width;
  const float m1 = s1 * inv, m2 = s2 * inv;
  for (int g = c.q; g < ngrp; g += 4) {
    float v[32];
    tmem_ld32(c.tmem_lane + g * 32, v);
    uint4 xu[4], yu[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      xu[k] = l_ldg16(xrow + (size_t)(g * 4 + k) * kLTileChunk);
      yu[k] = (kAct == kActSilu) ? make_uint4(0, 0, 0, 0) : l_ldg16(yrow + (size_t)(g * 4 + k) * kLTileChunk);
    }
    tmem_ld_wait();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float xh[8], y[8], gp[8];
      l_unpack8(xu[k], xh);
      l_unpack8(yu[k], y);
#pragma unroll
      for (int i = 0; i < 8; i += 4) {
        const float4 w4 = lds4(gam + (g * 32 + k * 8 + i) * 4);
        const float gw[4] = {w4.x, w4.y, w4.z, w4.w};
        float gs[4] = {0.f, 0.f, 0.f, 0.f}, bs[4] = {0.f, 0.f, 0.f, 0.f};
        if (kAct == kActSilu) {
          const float4 a4 = lds4(gsc + (g * 32 + k * 8 + i) * 4), b4 = lds4(bsc + (g * 32 + k * 8 + i) * 4);
          gs[0] = a4.x; gs[1] = a4.y; gs[2] = a4.z; gs[3] = a4.w;
          bs[0] = b4.x; bs[1] = b4.y; bs[2] = b4.z; bs[3] = b4.w;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float g_x = v[k * 8 + i + e] * act_grad<kAct>(y[i + e], xh[i + e], gs[e], bs[e]) * gw[e];
          gp[i + e] = c.valid ? rstd * (g_x - m1 - xh[i + e] * m2) : 0.f;
        }
      }
      st_chunk_g(orow + (size_t)(g * 4 + k) * kLTileChunk, gp);
    }
  }
  LPROBE(c, 4);
}

// What the backward epilogues read from HBM (recorded panels written a whole forward ago, upstream gradients read
// once) is pulled into L2 while the work item's MMAs run: issued before the accumulator wait, no registers held.
__device__ __forceinline__ void l_prefetch(const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }
__device__ __forceinline__ void lepib_prefetch(const LCtx& c, int type) {
  const LayerArgs& p = *c.p;
  const bool lead = (c.lane & 7) == 0;   // 8 rows x 16 B share a 128-byte line of a panel chunk
  const size_t ro = (size_t)c.mt * p.aux_tile_bytes + (size_t)c.row * 16;
  if (type == LB_ELU) {
    if (lead)
      for (int ch = c.q * 4; ch * 8 < p.n; ch += 16)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int col = c.nt * p.n + (ch + k) * 8;
          if ((ch + k) * 8 < p.n && col < p.out_cols) l_prefetch(p.aux[0] + ro + (size_t)(col / 8) * kLTileChunk);
        }
  } else if (type == LB_GRU) {
    if (lead)
      for (int g = c.q; g * 16 < p.n; g += 4) {
        const int col = c.nt * p.n + g * 16;
        if (col >= p.Dp) break;
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
          for (int f = 0; f < 5; ++f) l_prefetch(p.aux[f] + ro + (size_t)(col / 8 + k) * kLTileChunk);
      }
  } else if (type == LB_CARRY) {
    if (p.f1 && c.valid)
      for (int blk = c.q; blk * 32 < p.n; blk += 4) {
        const int col = c.nt * p.n + blk * 32;
        if (col < p.D) l_prefetch(p.f1 + (c.b * (p.H + 1) + p.t) * p.D + col);
      }
  } else if (type == LB_LN || type == LB_LN_SILU || type == LB_LN_RELU) {
    if (lead)
      for (int g = c.q; g < p.n / 32; g += 4)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const size_t ch = (size_t)c.nt * (p.n / 8) + (size_t)(g * 4 + k);
          l_prefetch(p.aux[0] + ro + ch * kLTileChunk);
          if (type != LB_LN_SILU) l_prefetch(p.aux[2] + ro + ch * kLTileChunk);
        }
  }
}

// --------------------------------------------------------------------------
// the kernel
// --------------------------------------------------------------------------
// kType: the fused epilogue (LE_*), kS: classes per categorical group (LE_SAMPLE_DISC) - one instantiation each, so
// that no epilogue pays for another's registers (96 per thread at 576 threads)
template <int kType, int kS>
__global__ void __launch_bounds__(kLThreads, 1) tc_layer_kernel(const __grid_constant__ LayerArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t ring = smem_base;
  const uint32_t stage_base = smem_base + kLSlots * kLSlotBytes;
  const uint32_t prm = stage_base + kLStageBytes;
  const uint32_t scratch = prm + kLParamBytes;
  const uint32_t xchg = scratch + kLScratchBytes;
  const uint32_t bars = xchg + kLXchgBytes;
  auto full_bar = [&](uint32_t s) { return bars + 8u * s; };
  auto empty_bar = [&](uint32_t s) { return bars + 64u + 8u * s; };
  auto accf_bar = [&](uint32_t j) { return bars + 128u + 8u * j; };
  auto acce_bar = [&](uint32_t j) { return bars + 144u + 8u * j; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + (bars - smem_base) + 160);
  const uint32_t xbar = bars + 176u;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int nitems = p.MT * p.NT;

  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < (uint32_t)kLSlots; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (uint32_t j = 0; j < 2; ++j) { mbar_init(accf_bar(j), 1); mbar_init(acce_bar(j), kLEpiWarps); }
    if (p.csplit > 1) mbar_init(xbar, 4u * (uint32_t)p.csplit);   // lane 0 of the q == 0 warp of every quarter of every CTA
    fence_barrier_init();
  }
  if (warp == kLMmaWarp) { tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  if (p.csplit > 1) cluster_sync_all();   // the peers' exchange barriers exist before anybody signals them
  tc_fence_after();
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation) overlapped the previous layer's
  // tail; from here on this grid reads what that layer wrote.  The next layer may start its own set-up right away
  // (every grid of the chain is a single wave, so no CTA of this grid is still waiting for an SM).
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t w_in_slot = (uint32_t)p.kc * 4096u;   // the weight K-slice follows the activation K-slice in a slot
  const bool tr = p.trace != nullptr && blockIdx.x == 0;
#define LTRACE(k) do { if (tr) p.trace[k] = clock64(); } while (0)
  if (tr && threadIdx.x == 0) p.trace[0] = clock64();

  if (warp == kLProducerWarp) {
    // kLSlots issuing lanes: lane j owns ring slot j (stage numbers j, j + 4, ...).  One lane sustains only one bulk
    // copy per ~415 cycles whatever its size; four lanes keep the ring fed.
    if (lane < kLSlots) {
      uint32_t sidx = 0, phase = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int mt = item / p.NT, nt = item % p.NT;
        const uint8_t* wt = p.w + (size_t)nt * p.w_tile_bytes;
        for (int sg = 0; sg < p.nseg; ++sg) {
          const LSeg& s = p.seg[sg];
          const uint8_t* a = s.a + (size_t)mt * s.a_tile_bytes;
          const uint8_t* w = wt + s.w_off;
          for (int k0 = 0; k0 < (int)s.k16; k0 += p.kc, ++sidx) {
            if ((int)(sidx & (kLSlots - 1)) != lane) continue;
            const uint32_t nk = (uint32_t)min(p.kc, (int)s.k16 - k0);
            const uint32_t ab = nk * 4096u, wb = nk * (uint32_t)s.n * 32u;
            mbar_wait_backoff_free(empty_bar(lane), phase ^ 1u);
            if (sidx == 0) LTRACE(1);
            mbar_arrive_expect_tx(full_bar(lane), ab + wb);
            bulk_g2s(ring + lane * kLSlotBytes, a + (size_t)k0 * 4096u, ab, full_bar(lane));
            bulk_g2s(ring + lane * kLSlotBytes + w_in_slot, w + (size_t)k0 * s.n * 32u, wb, full_bar(lane));
            phase ^= 1u;
          }
        }
        if (item == (int)blockIdx.x && lane == 0) LTRACE(2);    // lane 0's copies of the first work item issued
      }
    }
  } else if (warp == kLMmaWarp) {
    if (lane == 0) {
      uint32_t slot = 0, phase = 0, it = 0;
      constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);   // SBO = 128 B, descriptor version 1
      for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
        const uint32_t ab = p.dbuf ? (it & 1u) : 0u;
        const uint32_t use = p.dbuf ? (it >> 1) : it;      // how many times this accumulator buffer was used before
        if (use > 0) { mbar_wait_backoff_free(acce_bar(ab), (use - 1u) & 1u); tc_fence_after(); }
        const uint32_t tm0 = tmem_base + ab * 256u;
        for (int sg = 0; sg < p.nseg; ++sg) {
          // everything that does not change inside the K loop is computed here: the issuing lane's own instruction
          // stream is the critical path (~4 cycles of issue latency per dependent integer instruction)
          const LSeg& s = p.seg[sg];
          const uint32_t n = s.n;
          const uint32_t n_a = n > 256u ? 256u : n, n_b = n - n_a;          // pieces of <= 256 columns
          const uint32_t id_a = make_idesc_bf16(128, n_a), id_b = make_idesc_bf16(128, n_b ? n_b : 16u);
          const uint32_t tm_a = tm0 + s.tmem_col, tm_b = tm_a + 256u;
          const uint32_t bstep = n * 2u;                                    // n * 32 bytes >> 4
          const uint32_t blbo = n << 16;                                    // LBO = n * 16 bytes
          const int k16 = (int)s.k16, kc = p.kc;
          for (int k0 = 0; k0 < k16; k0 += kc) {
            const int nk = min(kc, k16 - k0);
            mbar_wait_backoff_free(full_bar(slot), phase);
            tc_fence_after();
            if (it == 0 && sg == 0 && k0 == 0) LTRACE(3);    // first stage landed
            const uint32_t a_addr = ring + slot * kLSlotBytes;
            uint32_t alo = ((a_addr >> 4) & 0x3FFFu) | ((2048u >> 4) << 16);
            uint32_t blo = (((a_addr + w_in_slot) >> 4) & 0x3FFFu) | blbo;
            int i = 0;
            if (k0 == 0) {   // first K step of the segment: fresh / split accumulate rules
              if (s.split) {
                mma_bf16_lohi(tm_a, alo, kDescHi, blo, kDescHi, make_idesc_bf16(128, s.split), 1u);
                mma_bf16_lohi(tm_a + s.split, alo, kDescHi, blo + (uint32_t)s.split, kDescHi,
                              make_idesc_bf16(128, n - s.split), 0u);
              } else {
                const uint32_t acc = s.fresh ? 0u : 1u;
                mma_bf16_lohi(tm_a, alo, kDescHi, blo, kDescHi, id_a, acc);
                if (n_b) mma_bf16_lohi(tm_b, alo, kDescHi, blo + 256u, kDescHi, id_b, acc);
              }
              alo += 256u; blo += bstep; i = 1;
            }
            if (n_b) {
#pragma unroll 2
              for (; i < nk; ++i) {
                mma_bf16_lohi(tm_a, alo, kDescHi, blo, kDescHi, id_a, 1u);
                mma_bf16_lohi(tm_b, alo, kDescHi, blo + 256u, kDescHi, id_b, 1u);
                alo += 256u; blo += bstep;
              }
            } else {
#pragma unroll 4
              for (; i < nk; ++i) {
                mma_bf16_lohi(tm_a, alo, kDescHi, blo, kDescHi, id_a, 1u);
                alo += 256u; blo += bstep;
              }
            }
            tc_commit(empty_bar(slot));
            if (++slot == (uint32_t)kLSlots) { slot = 0; phase ^= 1u; }
          }
        }
        tc_commit(accf_bar(ab));
        if (it == 0) LTRACE(4);                      // first work item's MMAs issued
      }
      LTRACE(5);                                     // all MMAs issued
    }
  } else {
    LCtx c;
    c.p = &p;
    c.prm = prm;
    c.scratch = scratch;
    c.stage = stage_base + (uint32_t)warp * 2048u;
    c.xchg = xchg; c.xbar = xbar;
    c.quarter = warp & 3;
    c.q = warp >> 2;
    c.lane = lane;
    c.row = c.quarter * 32 + lane;
    uint32_t it = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
      const uint32_t ab = p.dbuf ? (it & 1u) : 0u;
      const uint32_t use = p.dbuf ? (it >> 1) : it;
      c.mt = item / p.NT; c.nt = item % p.NT;
      c.xphase = it & 1u;
      c.b = (long long)c.mt * 128 + c.row;
      c.valid = c.b < p.B;
      c.tmem_lane = tmem_base + ab * 256u + ((uint32_t)(c.quarter * 32) << 16);
      c.out_row = p.out + (size_t)c.mt * p.out_tile_bytes + (size_t)c.row * 16;
      c.probe = (tr && it == 0 && threadIdx.x == 0) ? p.trace + 16 : nullptr;
      // this work item's epilogue parameters -> shared memory (the previous item's readers are done)
      l_epi_sync();
      {
        const float4* src = reinterpret_cast<const float4*>(p.ep + (size_t)c.nt * p.ep_floats);
        for (uint32_t i = threadIdx.x; i < p.ep_floats / 4; i += kLEpiThreads) {
          const float4 v = src[i];
          asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(prm + i * 16), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        }
      }
      l_epi_sync();
      LGruPre gpf;
      if (kType == LE_GRU) lgru_prefetch(c, gpf);
      if (kType == LE_SAMPLE_DISC && c.valid) {
        // the Exp(1) draws of this row are read once, from HBM: pull them into L2 while the MMAs run
        const int P = p.C * p.S;
        const float* qrow = p.noise + ((long long)p.t * p.B + c.b) * P + c.nt * p.n;
        for (int o = c.q * 32; o < p.n && c.nt * p.n + o < P; o += 128)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(qrow + o));
      }
      if (l_is_bwd(kType)) lepib_prefetch(c, kType);
      if (it == 0 && threadIdx.x == 0) LTRACE(6);    // parameters staged, waiting for the accumulator
      mbar_wait_backoff_quiet(accf_bar(ab), use & 1u);
      tc_fence_after();
      if (it == 0 && threadIdx.x == 0) LTRACE(7);    // accumulator complete
      if (kType == LE_LN_ELU) lepi_ln_act<kActElu>(c);
      else if (kType == LE_LN_SILU) lepi_ln_act<kActSilu>(c);
      else if (kType == LE_LN_RELU) lepi_ln_act<kActRelu>(c);
      else if (kType == LE_OUT) lepi_out(c);
      else if (kType == LB_LN_SILU) lepib_ln<kActSilu>(c);
      else if (kType == LB_LN_RELU) lepib_ln<kActRelu>(c);
      else if (kType == LB_XGRAD) lepib_xgrad(c);
      else if (kType == LE_BIAS_ELU) lepi_bias_elu(c);
      else if (kType == LE_MU) lepi_mu(c);
      else if (kType == LE_GRU) lepi_gru(c, gpf);
      else if (kType == LE_SAMPLE_CONT) lepi_sample_cont(c);
      else if (kType == LB_ELU) lepib_elu(c);
      else if (kType == LB_GRU) lepib_gru(c);
      else if (kType == LB_CARRY) lepib_carry(c);
      else if (kType == LB_EMB) lepib_emb(c);
      else if (kType == LB_LN) lepib_ln<kActElu>(c);
      else if (kType == LB_POST) lepib_post(c);
      else lepi_sample_disc<(kS > 0 ? kS : 4)>(c);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acce_bar(ab));
      if (it == 0 && threadIdx.x == 0) LTRACE(8);    // first epilogue done
    }
    if (threadIdx.x == 0) { LTRACE(9); if (tr) p.trace[10] = (long long)it; }
  }
  tc_fence_before();
  __syncthreads();
  if (p.csplit > 1) cluster_sync_all();   // no CTA leaves while a peer may still write into its exchange area
  if (warp == kLMmaWarp) tmem_dealloc(tmem_base, 512);
  if (tr && threadIdx.x == 0) p.trace[11] = clock64();
#undef LTRACE
}

typedef void (*LayerKernel)(const LayerArgs);
static LayerKernel layer_kernel_for(int type, int S) {
  switch (type) {
    case LE_LN_ELU: return tc_layer_kernel<LE_LN_ELU, 0>;
    case LE_BIAS_ELU: return tc_layer_kernel<LE_BIAS_ELU, 0>;
    case LE_MU: return tc_layer_kernel<LE_MU, 0>;
    case LE_GRU: return tc_layer_kernel<LE_GRU, 0>;
    case LE_SAMPLE_CONT: return tc_layer_kernel<LE_SAMPLE_CONT, 0>;
    case LB_ELU: return tc_layer_kernel<LB_ELU, 0>;
    case LB_GRU: return tc_layer_kernel<LB_GRU, 0>;
    case LB_CARRY: return tc_layer_kernel<LB_CARRY, 0>;
    case LB_EMB: return tc_layer_kernel<LB_EMB, 0>;
    case LB_LN: return tc_layer_kernel<LB_LN, 0>;
    case LB_POST: return tc_layer_kernel<LB_POST, 0>;
    case LE_LN_SILU: return tc_layer_kernel<LE_LN_SILU, 0>;
    case LE_LN_RELU: return tc_layer_kernel<LE_LN_RELU, 0>;
    case LE_OUT: return tc_layer_kernel<LE_OUT, 0>;
    case LB_LN_SILU: return tc_layer_kernel<LB_LN_SILU, 0>;
    case LB_LN_RELU: return tc_layer_kernel<LB_LN_RELU, 0>;
    case LB_XGRAD: return tc_layer_kernel<LB_XGRAD, 0>;
    default: break;
  }
  return S == 32 ? tc_layer_kernel<LE_SAMPLE_DISC, 32> : S == 16 ? tc_layer_kernel<LE_SAMPLE_DISC, 16>
       : S == 8 ? tc_layer_kernel<LE_SAMPLE_DISC, 8> : tc_layer_kernel<LE_SAMPLE_DISC, 4>;
}

// start states -> index 0 of the sequence outputs and the bf16 panels (all 128 rows of every tile, zeros past B)
__global__ void tcl_init_kernel(const float* __restrict__ deter0, const float* __restrict__ stoch0, float* deter_seq,
                                float* stoch_seq, uint8_t* hpanel, uint8_t* spanel, int B, int H, int D, int S_in,
                                int Dp, int Sip) {
  const int mt = blockIdx.y;
  const int chunks = Dp / 8 + Sip / 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < chunks * 128; i += gridDim.x * blockDim.x) {
    const int row = i % 128, ch = i / 128;
    const long long b = (long long)mt * 128 + row;
    const bool isd = ch < Dp / 8;
    const int k8 = isd ? ch : ch - Dp / 8;
    const int W = isd ? D : S_in;
    const float* src = isd ? deter0 : stoch0;
    float* seq = isd ? deter_seq : stoch_seq;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int col = k8 * 8 + e;
      v[e] = (src && b < B && col < W) ? src[b * W + col] : 0.f;
      if (b < B && col < W) seq[(b * (H + 1)) * W + col] = v[e];
    }
    uint8_t* panel = isd ? hpanel + (size_t)mt * (Dp / 8) * kLTileChunk : spanel + (size_t)mt * (Sip / 8) * kLTileChunk;
    st_chunk_g(panel + (size_t)k8 * kLTileChunk + row * 16, v);
  }
}

// --------------------------------------------------------------------------
// BPTT helpers outside the GEMM chain
// --------------------------------------------------------------------------
// stoch[b][t+1][c*S + j] = (j == idx[b][t][c]) for every step of a discrete rollout (state/discrete.py:27-32 forward value)
__global__ void tcl_onehot_kernel(const int32_t* __restrict__ idx, float* __restrict__ stoch, long long B, int H, int C, int S) {
  const long long n4 = B * H * (long long)(C * S / 4);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int per = C * S / 4;
    const long long bt = i / per;
    const int e0 = (int)(i % per) * 4;
    const long long b = bt / H;
    const int t = (int)(bt % H);
    const int v = idx[bt * C + e0 / S], j0 = e0 % S;
    *reinterpret_cast<float4*>(stoch + (b * (H + 1) + t + 1) * (long long)(C * S) + e0) =
        make_float4(j0 == v ? 1.f : 0.f, j0 + 1 == v ? 1.f : 0.f, j0 + 2 == v ? 1.f : 0.f, j0 + 3 == v ? 1.f : 0.f);
  }
}

// carry rows of dL/dh: the upstream gradient of the LAST deter state (zero where absent / padding)
__global__ void tcl_carry_init_kernel(float* carry_h, const float* g_deter, int B, int H, int D, int Dp, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // panel order: [tile][col/4][row][4]
  if (i >= n) return;
  const int e = (int)(i & 3), row = (int)((i >> 2) & 127);
  const long long ch = i >> 9;
  const long long mt = ch / (Dp / 4);
  const int col = (int)(ch % (Dp / 4)) * 4 + e;
  const long long b = mt * 128 + row;
  carry_h[i] = (g_deter && b < B && col < D) ? g_deter[(b * (H + 1) + H) * D + col] : 0.f;
}

struct SbArgs {
  const float *carry_s, *g_stoch, *g_da, *g_db, *dist_a, *dist_b, *noise;
  uint8_t* out;   // g_out panel [MT][Pp/8][128][8] bf16
  int B, H, t, S, C, Sp, Sip, Pp, MT;
};
__device__ __forceinline__ void panel_store_bf16(uint8_t* panel, int Pp, long long prow, int col, float v) {
  const long long mt = prow >> 7;
  const int row = (int)(prow & 127);
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(panel + (size_t)mt * (Pp / 8) * kLTileChunk + (size_t)(col / 8) * kLTileChunk +
                                                        (size_t)row * 16) + (col & 7);
  *dst = __float2bfloat16_rn(v);
}
// straight-through categorical backward (state/discrete.py:27-30): for the state produced by step t,
//   G = carried dL/ds + dL/dstoch[:, t+1];  dlogits = p (G - sum_k p_k G_k) + dL/dlogits[:, t+1],  p = softmax(logits[:, t+1])
// one warp per (padded row, kSbGroups consecutive groups), lane = class: the loads of all its groups are in flight together
// (one group per warp left the kernel latency-bound at ~2.3 TB/s)
constexpr int kSbGroups = 4;
__global__ void tcl_sample_bwd_disc_kernel(const SbArgs a) {
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int gpr = (a.C + kSbGroups - 1) / kSbGroups;      // warps per row
  if (warp >= (long long)a.MT * 128 * gpr) return;
  const long long prow = warp / gpr;
  const int c0 = (int)(warp % gpr) * kSbGroups;
  const int P = a.C * a.S;
  const bool valid = prow < a.B;
  const long long o0 = (prow * (a.H + 1) + a.t + 1) * P;
  float l[kSbGroups], G[kSbGroups], gl[kSbGroups];
#pragma unroll
  for (int k = 0; k < kSbGroups; ++k) {
    const int c = c0 + k;
    l[k] = -INFINITY; G[k] = 0.f; gl[k] = 0.f;
    if (valid && c < a.C && lane < a.S) {
      const long long o = o0 + (long long)c * a.S + lane;
      l[k] = a.dist_a[o];
      if (a.carry_s) { const int cc = c * a.S + lane; G[k] = a.carry_s[((prow >> 7) * (a.Sip / 4) + (cc >> 2)) * 512 + (prow & 127) * 4 + (cc & 3)]; }
      if (a.g_stoch) G[k] += a.g_stoch[o];
      if (a.g_da) gl[k] = a.g_da[o];
    }
  }
#pragma unroll
  for (int k = 0; k < kSbGroups; ++k) {
    const int c = c0 + k;
    if (c >= a.C) break;
    float mx = l[k];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
    const float e = (valid && lane < a.S) ? expf(l[k] - mx) : 0.f;
    float se = e, sg = e * G[k];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) { se += __shfl_xor_sync(0xffffffffu, se, s); sg += __shfl_xor_sync(0xffffffffu, sg, s); }
    if (lane < a.S) {
      const float v = valid ? (e / se) * (G[k] - sg / se) + gl[k] : 0.f;
      panel_store_bf16(a.out, a.Pp, prow, c * a.S + lane, v);
    }
  }
  // zero the padding columns [P, Pp) once per row
  if (c0 == 0 && P + lane < a.Pp) panel_store_bf16(a.out, a.Pp, prow, P + lane, 0.f);
}
// Gaussian reparameterised sample backward (rssm.py:179-181): panel [g_mean (Sp) | g_raw (Sp)],
//   g_mean = dL/dmean + G,  g_raw = (dL/dstd + G eps) sigmoid(raw),  sigmoid(raw) = 1 - exp(-(std - 0.1))
// one thread per (padded row, 8-column chunk)
__global__ void tcl_sample_bwd_cont_kernel(const SbArgs a) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nch = a.Sp / 8;
  if (i >= (long long)a.MT * 128 * nch) return;
  const long long prow = i / nch;
  const int col = (int)(i % nch) * 8;
  const bool valid = prow < a.B;
  const long long o = (prow * (a.H + 1) + a.t + 1) * a.S + col;
  float gm[8], gr[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    gm[k] = gr[k] = 0.f;
    if (valid && col + k < a.S) {
      float G = a.carry_s ? a.carry_s[((prow >> 7) * (a.Sip / 4) + ((col + k) >> 2)) * 512 + (prow & 127) * 4 + ((col + k) & 3)] : 0.f;
      if (a.g_stoch) G += a.g_stoch[o + k];
      const float eps = a.noise[((long long)a.t * a.B + prow) * a.S + col + k];
      const float sig = 1.f - __expf(-(a.dist_b[o + k] - 0.1f));
      gm[k] = (a.g_da ? a.g_da[o + k] : 0.f) + G;
      gr[k] = ((a.g_db ? a.g_db[o + k] : 0.f) + G * eps) * sig;
    }
  }
  const long long mt = prow >> 7;
  const int row = (int)(prow & 127);
  uint8_t* base = a.out + (size_t)mt * (a.Pp / 8) * kLTileChunk + (size_t)row * 16;
  st_chunk_g(base + (size_t)(col / 8) * kLTileChunk, gm);
  st_chunk_g(base + (size_t)((a.Sp + col) / 8) * kLTileChunk, gr);
}

// weight gradients: dW[n][m] += sum over 128-row blocks of A_blk[m][rows] . B_blk[n][rows]
//   A = in-feature panel (actor block input), 128 in-features per CTA tile (TMEM lanes)
//   B = out-feature panel (g_pre / g_mu), up to 256 out-features per CTA tile (TMEM columns)
// For a contraction over ROWS a panel tile [feature/8][128 rows][8] is exactly the canonical no-swizzle MN-major
// operand: 8 features contiguous, 8-row groups 128 B apart, feature chunks 2 KB apart.  Split-K over the SMs, fp32 atomics.
struct Wg2Params {
  const uint8_t* a_base; size_t a_blk;
  const uint8_t* b_base; size_t b_blk;
  int nblocks, m_cols, m_valid, n_cols, n_valid;
  float* dW; int ldw;
  int mtiles, ntiles, splits;
};
constexpr int kWg2Stages = 2, kWg2Threads = 192;
constexpr uint32_t kWg2A = 32768, kWg2B = 65536, kWg2Stage = kWg2A + kWg2B;
constexpr uint32_t kWg2Smem = kWg2Stages * kWg2Stage + 1024;
__global__ void __launch_bounds__(kWg2Threads, 1) tcl_wgrad_kernel(const Wg2Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bars = smem_base + kWg2Stages * kWg2Stage;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 64u + 8u * s; };
  const uint32_t acc_bar = bars + 128u;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kWg2Stages * kWg2Stage + 136);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, split = blockIdx.y;
  const int mt = tile / p.ntiles, nt = tile % p.ntiles;
  const int m0 = mt * 128, n0 = nt * 256;
  const int mc = min(128, p.m_cols - m0), nc = min(256, p.n_cols - n0);
  const uint32_t a_bytes = (uint32_t)(mc / 8) * kLTileChunk, b_bytes = (uint32_t)(nc / 8) * kLTileChunk;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kWg2Stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
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
        bulk_g2s(smem_base + slot * kWg2Stage, p.a_base + blk * p.a_blk + (size_t)(m0 / 8) * kLTileChunk, a_bytes, full_bar(slot));
        bulk_g2s(smem_base + slot * kWg2Stage + kWg2A, p.b_base + blk * p.b_blk + (size_t)(n0 / 8) * kLTileChunk, b_bytes,
                 full_bar(slot));
        if (++slot == kWg2Stages) { slot = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, (uint32_t)nc) | (1u << 15) | (1u << 16);   // A and B MN-major
      uint32_t slot = 0, phase = 0;
      for (int i = 0; i < nmine; ++i) {
        mbar_wait(full_bar(slot), phase);
        tc_fence_after();
        const uint32_t a0 = smem_base + slot * kWg2Stage, b0 = a0 + kWg2A;
#pragma unroll
        for (int k = 0; k < 8; ++k) {   // 16 rows (K) per MMA: two 8-row groups = 256 B further on
          const uint64_t da = make_smem_desc(a0 + k * 256, 128u, kLTileChunk);
          const uint64_t db = make_smem_desc(b0 + k * 256, 128u, kLTileChunk);
          mma_bf16_ss(tmem_base, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        tc_commit(empty_bar(slot));
        if (++slot == kWg2Stages) { slot = 0; phase ^= 1u; }
      }
      tc_commit(acc_bar);
    }
  } else {
    // epilogue: thread = TMEM lane = in-feature m; consecutive lanes -> consecutive dW columns (coalesced atomics)
    mbar_wait_backoff(acc_bar, 0);
    tc_fence_after();
    if (nmine > 0) {
      const int m = m0 + warp * 32 + lane;
      const bool ok = m < p.m_valid;
      const uint32_t tl = tmem_base + ((uint32_t)(warp * 32) << 16);
      for (int j0 = 0; j0 < nc; j0 += 16) {
        float v[16];
        tmem_ld16(tl + j0, v);
        tmem_ld_wait();
        if (ok) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int n = n0 + j0 + j;
            if (n < p.n_valid) atomicAdd(p.dW + (size_t)n * p.ldw + m, v[j]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, 256);
}

// column sums over panel tiles: out[c] += sum over blocks and their 128 rows of X[c] (* Y[c])
struct Cs2Params {
  const uint8_t* x_base; size_t x_blk;
  const uint8_t* y_base; size_t y_blk;   // null: plain column sum
  // fused form (z_base != null): out += sum X (ln_b), out2 += sum X Y (ln_g), out3 += sum Z (lin_b) in ONE pass over the
  // g_u, x-hat and g_pre panels of a block (three separate passes read four panels)
  const uint8_t* z_base; size_t z_blk;
  int nblocks, valid;
  float* out; float* out2; float* out3;
};
constexpr int kCs2Warps = 8;
__device__ __forceinline__ void cs2_reduce_store(float (&acc)[8], float* out, int chunk, int valid, float (*part)[8]) {
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
    for (int w2 = 0; w2 < kCs2Warps; ++w2) s += part[w2][threadIdx.x];
    const int col = chunk * 8 + threadIdx.x;
    if (col < valid) atomicAdd(out + col, s);
  }
}
__global__ void __launch_bounds__(kCs2Warps * 32) tcl_colsum_kernel(const Cs2Params p) {
  const int chunk = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[8], acc2[8], acc3[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { acc[i] = 0.f; acc2[i] = 0.f; acc3[i] = 0.f; }
  const size_t off = (size_t)chunk * kLTileChunk + (size_t)lane * 16;
  const bool fused = p.z_base != nullptr;
  for (size_t blk = (size_t)blockIdx.y * kCs2Warps + warp; blk < (size_t)p.nblocks; blk += (size_t)gridDim.y * kCs2Warps) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float x[8];
      ld_chunk_g(p.x_base + blk * p.x_blk + off + r * 512, x);
      if (p.y_base) {
        float y[8];
        ld_chunk_g(p.y_base + blk * p.y_blk + off + r * 512, y);
        if (fused) {
          float z[8];
          ld_chunk_g(p.z_base + blk * p.z_blk + off + r * 512, z);
#pragma unroll
          for (int i = 0; i < 8; ++i) { acc[i] += x[i]; acc2[i] = fmaf(x[i], y[i], acc2[i]); acc3[i] += z[i]; }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = fmaf(x[i], y[i], acc[i]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += x[i];
      }
    }
  }
  __shared__ float part[kCs2Warps][8];
  cs2_reduce_store(acc, p.out, chunk, p.valid, part);
  if (fused) {
    cs2_reduce_store(acc2, p.out2, chunk, p.valid, part);
    cs2_reduce_store(acc3, p.out3, chunk, p.valid, part);
  }
}

// --------------------------------------------------------------------------
// host: plan, packing, memory layout, launch chains
// --------------------------------------------------------------------------
// logical panel buffers.  Forward fields first (activations; recorded per step when training), then the BPTT's
// gradient panels.  F_Y / F_XH / F_RSTD / G_GP / G_GU take one slot per actor block.
enum {
  F_H = 0, F_HN, F_S, F_SN, F_A, F_X, F_HID, F_GR, F_GZ, F_GN, F_GHN,
  F_OBS, F_ACT, F_HIDQ, F_SPRI,   // posterior filtering: observation-embedding / action panels per step, posterior hidden, prior-sample sink
  F_Y0, F_XH0 = F_Y0 + WMRL_MAX_ACTOR_LAYERS, F_RSTD0 = F_XH0 + WMRL_MAX_ACTOR_LAYERS,
  G_GO = F_RSTD0 + WMRL_MAX_ACTOR_LAYERS, G_GHID, G_GATES_I, G_GATES_H, G_GX, G_GMU, G_OQ, G_HIDQ,
  G_GP0, G_GU0 = G_GP0 + WMRL_MAX_ACTOR_LAYERS, BUF_COUNT = G_GU0 + WMRL_MAX_ACTOR_LAYERS
};

struct LSegH { int buf, k16, n, tmem_col, split, fresh; uint32_t w_off; };
struct LLayerH {
  int type = 0, NT = 1, n = 0, nseg = 0, kc = 1, dbuf = 0, out = 0, layer = 0;
  int csplit = 0;            // > 1: the NT = csplit column tiles of a LayerNorm layer run as one cluster
  LSegH seg[3];
  size_t w_off = 0;          // into the data area
  uint32_t w_tile_bytes = 0;
  size_t ep_off = 0;         // float offset into eparams
  uint32_t ep_floats = 0;
};
struct LPlan {
  bool ok = false;
  Plan pk;                   // pack ops / vec ops / image offsets (stage and job tables stay empty)
  std::vector<LLayerH> layers;    // forward chain of one step
  std::vector<LLayerH> blayers;   // backward chain of one step (after the sample backward)
  // the same chains with every LayerNorm layer split into `csplit` column tiles (one cluster per row tile, statistics
  // exchanged through DSMEM): used when a call has so few row tiles that one CTA per row tile leaves most SMs idle
  std::vector<LLayerH> layers_s, blayers_s;   // split over kLMaxSplit CTAs: few row tiles
  // split over CTA PAIRS for wide blocks (Ha > 256) at ANY row count: a 512-column block fills TMEM, so its epilogue was
  // exposed; two 256-column halves double-buffer their accumulators and the next item's MMAs run under the epilogue
  std::vector<LLayerH> layers_p, blayers_p;
  int csplit = 0, cpair = 0;
  int Dp = 0, Sip = 0, Sp = 0, Hdp = 0, Ha = 0, Pp = 0, Ep = 0;
};

static int ceil_to(int x, int m) { return (x + m - 1) / m * m; }

// one K segment of one column tile: rows `rs` of src[., ld], K range [k0, k0 + kvalid) padded to kpad
static void l_add_pack(LPlan& lp, const RowSpec& rs, const float* src, int ld, int k0, int kvalid, int kpad) {
  PackOp op{};
  op.src = src; op.ld = ld; op.trans = 0;
  op.nseg = rs.nseg;
  for (int q = 0; q < rs.nseg; ++q) { op.row0[q] = rs.seg[q].row0; op.valid[q] = rs.seg[q].valid; op.pad[q] = rs.seg[q].pad; }
  op.k0 = k0; op.kvalid = kvalid; op.nk8 = kpad / 8;
  op.nks = 0; op.kbegin = 0;
  op.split = 1;
  op.dst_off = lp.pk.data_bytes;
  op.bias = nullptr; op.bias_k = -1;
  lp.pk.pack_ops.push_back(op);
  lp.pk.data_bytes += (size_t)kpad * rs.N() * 2;
}
// dgrad operand: B rows are IN-features `rs` of the torch weight src[out][in] (ld = in-features), the K index runs over
// OUT-features, given as up to 4 ranges {first, valid, padded}
struct KRange { int k0, valid, pad; };
static void l_add_pack_t(LPlan& lp, const RowSpec& rs, const float* src, int ld, std::initializer_list<KRange> ks) {
  PackOp op{};
  op.src = src; op.ld = ld; op.trans = 1;
  op.nseg = rs.nseg;
  for (int q = 0; q < rs.nseg; ++q) { op.row0[q] = rs.seg[q].row0; op.valid[q] = rs.seg[q].valid; op.pad[q] = rs.seg[q].pad; }
  int kpad = 0, i = 0;
  for (const KRange& k : ks) { op.kk0[i] = k.k0; op.kkvalid[i] = k.valid; op.kkpad[i] = k.pad; kpad += k.pad; ++i; }
  op.nks = i; op.kbegin = 0;
  op.k0 = 0; op.kvalid = kpad; op.nk8 = kpad / 8;
  op.split = 1;
  op.dst_off = lp.pk.data_bytes;
  op.bias = nullptr; op.bias_k = -1;
  lp.pk.pack_ops.push_back(op);
  lp.pk.data_bytes += (size_t)kpad * rs.N() * 2;
}

static int buf_width(const LPlan& lp, int buf) {
  if (buf == F_H || buf == F_HN || buf == F_X || buf == F_GR || buf == F_GZ || buf == F_GN || buf == F_GHN || buf == G_GX) return lp.Dp;
  if (buf == F_S || buf == F_SN || buf == F_SPRI) return lp.Sip;
  if (buf == F_A || buf == G_GMU || buf == F_ACT) return 16;
  if (buf == F_HID || buf == G_GHID || buf == F_HIDQ || buf == G_HIDQ) return lp.Hdp;
  if (buf == F_OBS) return lp.Ep;
  if (buf == G_GO || buf == G_OQ) return lp.Pp;
  if (buf == G_GATES_I || buf == G_GATES_H) return 4 * lp.Dp;
  return lp.Ha;   // F_Y, F_XH, G_GP, G_GU
}

static LPlan tcl_make_plan(const Dims& d, const wmrl_rssm_params* w, const wmrl_actor_params* aw) {
  LPlan lp;
  const bool disc = d.variant == WMRL_DISCRETE;
  if (d.Ha % 32 != 0 || d.Ha > 512 || d.Ha < 32 || d.A > 16 || d.La < 1) return lp;
  if (disc && !(d.S == 4 || d.S == 8 || d.S == 16 || d.S == 32)) return lp;
  const int Dp = ceil16(d.D), Sip = ceil16(d.S_in), Sp = ceil16(d.S), Hdp = ceil16(d.Hd);
  if (!disc && 2 * Sp > 256) return lp;
  const int Pp = disc ? Sip : 2 * Sp;
  lp.Dp = Dp; lp.Sip = Sip; lp.Sp = Sp; lp.Hdp = Hdp; lp.Ha = d.Ha; lp.Pp = Pp;
  const float* nul = nullptr;
  Plan& pk = lp.pk;
  auto kc_for = [&](int nmax) { int kc = (int)(kLSlotBytes / (4096u + (uint32_t)nmax * 32u)); return std::max(1, std::min(kc, 8)); };
  // even column tiles of at most `cap` columns, multiples of `unit`
  auto tiles = [&](int total, int cap, int unit, int& nt, int& wt) {
    nt = (total + cap - 1) / cap;
    wt = ceil_to((total + nt - 1) / nt, unit);
  };
  // ====================================================================== forward chain
  // ---- actor blocks: [h | s] (block 0) or y_{l-1} -> Linear -> LayerNorm -> ELU
  for (int l = 0; l < d.La; ++l) {
    LLayerH L;
    L.type = LE_LN_ELU; L.NT = 1; L.n = d.Ha; L.dbuf = d.Ha <= 256 ? 1 : 0; L.out = F_Y0 + l; L.layer = l;
    L.kc = kc_for(d.Ha);
    L.w_off = pk.data_bytes;
    const float* lw = aw ? aw->lin_w[l] : nul;
    RowSpec rs(0, d.Ha, d.Ha);
    if (l == 0) {
      L.nseg = 2;
      L.seg[0] = {F_H, Dp / 16, d.Ha, 0, 0, 1, 0u};
      l_add_pack(lp, rs, lw, d.F, 0, d.D, Dp);
      L.seg[1] = {F_S, Sip / 16, d.Ha, 0, 0, 0, (uint32_t)(pk.data_bytes - L.w_off)};
      l_add_pack(lp, rs, lw, d.F, d.D, d.S_in, Sip);
    } else {
      L.nseg = 1;
      L.seg[0] = {F_Y0 + l - 1, d.Ha / 16, d.Ha, 0, 0, 1, 0u};
      l_add_pack(lp, rs, lw, d.Ha, 0, d.Ha, d.Ha);
    }
    L.w_tile_bytes = (uint32_t)(pk.data_bytes - L.w_off);
    L.ep_off = add_vec(pk, aw ? aw->lin_b[l] : nul, nul, 0, d.Ha, d.Ha);
    add_vec(pk, aw ? aw->ln_g[l] : nul, nul, 0, d.Ha, d.Ha, 1.4426950408889634f);
    add_vec(pk, aw ? aw->ln_b[l] : nul, nul, 0, d.Ha, d.Ha, 1.4426950408889634f);
    L.ep_floats = 3u * d.Ha;
    lp.layers.push_back(L);
  }
  const int ylast = F_Y0 + d.La - 1;
  // ---- mu head
  {
    LLayerH L;
    L.type = LE_MU; L.NT = 1; L.n = 16; L.dbuf = 1; L.out = F_A; L.kc = kc_for(16); L.nseg = 1;
    L.w_off = pk.data_bytes;
    L.seg[0] = {ylast, d.Ha / 16, 16, 0, 0, 1, 0u};
    l_add_pack(lp, RowSpec(0, d.A, 16), aw ? aw->mu_w : nul, d.Ha, 0, d.Ha, d.Ha);
    L.w_tile_bytes = (uint32_t)(pk.data_bytes - L.w_off);
    L.ep_off = add_vec(pk, aw ? aw->mu_b : nul, nul, 0, d.A, 16);
    add_vec(pk, aw ? aw->bound : nul, nul, 0, d.A, 16);
    L.ep_floats = 32;
    lp.layers.push_back(L);
  }
  // ---- state-action embed: [s | a] -> Linear -> ELU
  {
    LLayerH L;
    L.type = LE_BIAS_ELU; L.dbuf = 1; L.out = F_X; L.nseg = 2;
    tiles(Dp, 256, 16, L.NT, L.n);
    L.kc = kc_for(L.n);
    L.w_off = pk.data_bytes;
    const int Ke = d.S_in + d.A;
    for (int j = 0; j < L.NT; ++j) {
      const size_t t0 = pk.data_bytes;
      RowSpec rs(j * L.n, d.D - j * L.n, L.n);
      L.seg[0] = {F_S, Sip / 16, L.n, 0, 0, 1, 0u};
      l_add_pack(lp, rs, w ? w->embed_w : nul, Ke, 0, d.S_in, Sip);
      L.seg[1] = {F_A, 1, L.n, 0, 0, 0, (uint32_t)(pk.data_bytes - t0)};
      l_add_pack(lp, rs, w ? w->embed_w : nul, Ke, d.S_in, d.A, 16);
      L.w_tile_bytes = (uint32_t)(pk.data_bytes - t0);
      const uint32_t at = add_vec(pk, w ? w->embed_b : nul, nul, j * L.n, std::max(0, std::min(L.n, d.D - j * L.n)), L.n);
      if (j == 0) L.ep_off = at;
    }
    L.ep_floats = (uint32_t)L.n;
    lp.layers.push_back(L);
  }
  // ---- GRU, 64 state columns per work item: TMEM [i_n | r | z | h_n]
  {
    LLayerH L;
    constexpr int cw = kGruTile;
    L.type = LE_GRU; L.dbuf = 1; L.out = F_HN; L.nseg = 2; L.n = 4 * cw;
    L.NT = (Dp + cw - 1) / cw;
    L.kc = kc_for(3 * cw);
    L.w_off = pk.data_bytes;
    for (int j = 0; j < L.NT; ++j) {
      const size_t t0 = pk.data_bytes;
      const int c0 = j * cw, v = std::max(0, std::min(cw, d.D - c0));
      RowSpec rx;   // W_ih rows [n | r | z] of the tile's state columns
      rx.add(2 * d.D + c0, v, cw); rx.add(c0, v, cw); rx.add(d.D + c0, v, cw);
      L.seg[0] = {F_X, Dp / 16, 3 * cw, 0, 0, 1, 0u};
      l_add_pack(lp, rx, w ? w->w_ih : nul, d.D, 0, d.D, Dp);
      RowSpec rh;   // W_hh rows [r | z | n]
      rh.add(c0, v, cw); rh.add(d.D + c0, v, cw); rh.add(2 * d.D + c0, v, cw);
      L.seg[1] = {F_H, Dp / 16, 3 * cw, cw, 2 * cw, 0, (uint32_t)(pk.data_bytes - t0)};
      l_add_pack(lp, rh, w ? w->w_hh : nul, d.D, 0, d.D, Dp);
      L.w_tile_bytes = (uint32_t)(pk.data_bytes - t0);
      const uint32_t at = add_vec(pk, w ? w->b_ih : nul, w ? w->b_hh : nul, c0, v, cw);       // b_r
      add_vec(pk, w ? w->b_ih : nul, w ? w->b_hh : nul, d.D + c0, v, cw);                    // b_z
      add_vec(pk, w ? w->b_ih : nul, nul, 2 * d.D + c0, v, cw);                              // b_in
      add_vec(pk, w ? w->b_hh : nul, nul, 2 * d.D + c0, v, cw);                              // b_hn
      if (j == 0) L.ep_off = at;
    }
    L.ep_floats = 4 * cw;
    lp.layers.push_back(L);
  }
  // ---- prior hidden: h' -> Linear -> ELU
  {
    LLayerH L;
    L.type = LE_BIAS_ELU; L.dbuf = 1; L.out = F_HID; L.nseg = 1;
    tiles(Hdp, 256, 16, L.NT, L.n);
    L.kc = kc_for(L.n);
    L.w_off = pk.data_bytes;
    for (int j = 0; j < L.NT; ++j) {
      const size_t t0 = pk.data_bytes;
      L.seg[0] = {F_HN, Dp / 16, L.n, 0, 0, 1, 0u};
      l_add_pack(lp, RowSpec(j * L.n, d.Hd - j * L.n, L.n), w ? w->prior0_w : nul, d.D, 0, d.D, Dp);
      L.w_tile_bytes = (uint32_t)(pk.data_bytes - t0);
      const uint32_t at = add_vec(pk, w ? w->prior0_b : nul, nul, j * L.n, std::max(0, std::min(L.n, d.Hd - j * L.n)), L.n);
      if (j == 0) L.ep_off = at;
    }
    L.ep_floats = (uint32_t)L.n;
    lp.layers.push_back(L);
  }
  // ---- prior head + sample
  {
    LLayerH L;
    L.dbuf = 1; L.out = F_SN; L.nseg = 1;
    L.w_off = pk.data_bytes;
    if (disc) {
      L.type = LE_SAMPLE_DISC;
      tiles(ceil_to(d.P, 32), 256, 32, L.NT, L.n);
      for (int j = 0; j < L.NT; ++j) {
        const size_t t0 = pk.data_bytes;
        L.seg[0] = {F_HID, Hdp / 16, L.n, 0, 0, 1, 0u};
        l_add_pack(lp, RowSpec(j * L.n, d.P - j * L.n, L.n), w ? w->prior2_w : nul, d.Hd, 0, d.Hd, Hdp);
        L.w_tile_bytes = (uint32_t)(pk.data_bytes - t0);
        const uint32_t at = add_vec(pk, w ? w->prior2_b : nul, nul, j * L.n, std::max(0, std::min(L.n, d.P - j * L.n)), L.n);
        if (j == 0) L.ep_off = at;
      }
      L.ep_floats = (uint32_t)L.n;
    } else {
      L.type = LE_SAMPLE_CONT; L.NT = 1; L.n = 2 * Sp;
      RowSpec rs;
      rs.add(0, d.S, Sp); rs.add(d.S, d.S, Sp);
      L.seg[0] = {F_HID, Hdp / 16, L.n, 0, 0, 1, 0u};
      l_add_pack(lp, rs, w ? w->prior2_w : nul, d.Hd, 0, d.Hd, Hdp);
      L.w_tile_bytes = (uint32_t)(pk.data_bytes - L.w_off);
      L.ep_off = add_vec(pk, w ? w->prior2_b : nul, nul, 0, d.S, Sp);
      add_vec(pk, w ? w->prior2_b : nul, nul, d.S, d.S, Sp);
      L.ep_floats = 2u * Sp;
    }
    L.kc = kc_for(L.n);
    lp.layers.push_back(L);
  }
  // ====================================================================== backward chain (dgrad through the frozen
  // world model and the actor; transposed weights; world_model.py:172, rssm.py:135)
  const uint32_t zero16 = add_vec(pk, nul, nul, 0, 0, 16);   // epilogues without parameters still load 16 floats
  // ---- dL/d(prior hidden) = g_out W2 (x ELU'(hid))
  {
    LLayerH L;
    L.type = LB_ELU; L.dbuf = 1; L.out = G_GHID; L.nseg = 1; L.layer = F_HID;
    tiles(Hdp, 256, 16, L.NT, L.n);
    L.kc = kc_for(L.n);
    L.w_off = pk.data_bytes;
    for (int j = 0; j < L.NT; ++j) {
      const size_t t0 = pk.data_bytes;
      L.seg[0] = {G_GO, Pp / 16, L.n, 0, 0, 1, 0u};
      RowSpec rs(j * L.n, d.Hd - j * L.n, L.n);
      if (disc) l_add_pack_t(lp, rs, w ? w->prior2_w : nul, d.Hd, {{0, d.P, Pp}});
      else l_add_pack_t(lp, rs, w ? w->prior2_w : nul, d.Hd, {{0, d.S, Sp}, {d.S, d.S, Sp}});
      L.w_tile_bytes = (uint32_t)(pk.data_bytes - t0);
    }
    L.ep_off = zero16; L.ep_floats = 0;
    lp.blayers.push_back(L);
  }
  // ---- dL/dh' = g_hid W1 + carried; GRU gate backward -> [g_n | g_r | g_z | g_n r]
  {
    LLayerH L;
    L.type = LB_GRU; L.dbuf = 1; L.out = G_GATES_I; L.nseg = 1;
    tiles(Dp, 128, 16, L.NT, L.n);
    L.kc = kc_for(L.n);
    L.w_off = pk.data_bytes;
    for (int j = 0; j < L.NT; ++j) {
      const size_t t0 = pk.data_bytes;
      L.seg[0] = {G_GHID, Hdp / 16, L.n, 0, 0, 1, 0u};
      l_add_pack_t(lp, RowSpec(j * L.n, d.D - j * L.n, L.n), w ? w->prior0_w : nul, d.D, {{0, d.Hd, Hdp}});
      L.w_tile_bytes = (uint32_t)(pk.data_bytes - t0);
    }
    L.ep_off = zero16; L.ep_floats = 0;
    lp.blayers.push_back(L);
  }
  // ---- dL/dx = [g_n | g_r | g_z] W_ih (x ELU'(x))
  {
    LLayerH L;
    L.type = LB_ELU; L.dbuf = 1; L.out = G_GX; L.nseg = 1; L.layer = F_X;
    tiles(Dp, 256, 16, L.NT, L.n);
    L.kc = kc_for(L.n);
    L.w_off = pk.data_bytes;
    for (int j = 0; j < L.NT; ++j) {
      const size_t t0 = pk.data_bytes;
      L.seg[0] = {G_GATES_I, 3 * Dp / 16, L.n, 0, 0, 1, 0u};
      l_add_pack_t(lp, RowSpec(j * L.n, d.D - j * L.n, L.n), w ? w->w_ih : nul, d.D,
                   {{2 * d.D, d.D, Dp}, {0, d.D, Dp}, {d.D, d.D, Dp}});
      L.w_tile_bytes = (uint32_t)(pk.data_bytes - t0);
    }
    L.ep_off = zero16; L.ep_floats = 0;
    lp.blayers.push_back(L);
  }
  // ---- dL/dh_t = [g_r | g_z | g_n r] W_hh + dL/dh' z (+ upstream)
  {
    LLayerH L;
    L.type = LB_CARRY; L.dbuf = 1; L.out = G_GX; L.nseg = 1;
    tiles(Dp, 256, 16, L.NT, L.n);
    L.kc = kc_for(L.n);
    L.w_off = pk.data_bytes;
    for (int j = 0; j < L.NT; ++j) {
      const size_t t0 = pk.data_bytes;
      L.seg[0] = {G_GATES_H, 3 * Dp / 16, L.n, 0, 0, 1, 0u};
      l_add_pack_t(lp, RowSpec(j * L.n, d.D - j * L.n, L.n), w ? w->w_hh : nul, d.D,
                   {{0, d.D, Dp}, {d.D, d.D, Dp}, {2 * d.D, d.D, Dp}});
      L.w_tile_bytes = (uint32_t)(pk.data_bytes - t0);
    }
    L.ep_off = zero16; L.ep_floats = 0;
    lp.blayers.push_back(L);
  }
  // ---- [dL/ds_t | dL/da_t] = g_x W_e: latent column tiles, then ONE action tile (tanh-squash backward -> g_mu)
  {
    LLayerH L;
    L.type = LB_EMB; L.dbuf = 1; L.out = G_GMU; L.nseg = 1;
    int nts;
    tiles(Sip, 256, 16, nts, L.n);
    L.NT = nts + 1;
    L.kc = kc_for(L.n);
    L.w_off = pk.data_bytes;
    const int Ke = d.S_in + d.A;
    for (int j = 0; j < L.NT; ++j) {
      const size_t t0 = pk.data_bytes;
      L.seg[0] = {G_GX, Dp / 16, L.n, 0, 0, 1, 0u};
      RowSpec rs = (j < nts) ? RowSpec(j * L.n, d.S_in - j * L.n, L.n) : RowSpec(d.S_in, d.A, L.n);
      l_add_pack_t(lp, rs, w ? w->embed_w : nul, Ke, {{0, d.D, Dp}});
      L.w_tile_bytes = (uint32_t)(pk.data_bytes - t0);
    }
    L.ep_off = add_vec(pk, aw ? aw->bound : nul, nul, 0, d.A, 16); L.ep_floats = 0;   // ep_floats 0: same block for every tile
    lp.blayers.push_back(L);
  }
  // ---- actor: dL/dy_{L-1} = g_mu W_mu, then per block LayerNorm+ELU backward and dL/dy_{l-1} = g_pre_l W_l
  for (int l = d.La - 1; l >= 0; --l) {
    LLayerH L;
    L.type = LB_LN; L.NT = 1; L.n = d.Ha; L.dbuf = d.Ha <= 256 ? 1 : 0; L.out = G_GP0 + l; L.layer = l; L.nseg = 1;
    L.kc = kc_for(d.Ha);
    L.w_off = pk.data_bytes;
    if (l == d.La - 1) {
      L.seg[0] = {G_GMU, 1, d.Ha, 0, 0, 1, 0u};
      l_add_pack_t(lp, RowSpec(0, d.Ha, d.Ha), aw ? aw->mu_w : nul, d.Ha, {{0, d.A, 16}});
    } else {
      L.seg[0] = {G_GP0 + l + 1, d.Ha / 16, d.Ha, 0, 0, 1, 0u};
      l_add_pack_t(lp, RowSpec(0, d.Ha, d.Ha), aw ? aw->lin_w[l + 1] : nul, d.Ha, {{0, d.Ha, d.Ha}});
    }
    L.w_tile_bytes = (uint32_t)(pk.data_bytes - L.w_off);
    L.ep_off = add_vec(pk, aw ? aw->ln_g[l] : nul, nul, 0, d.Ha, d.Ha, 1.4426950408889634f);
    add_vec(pk, aw ? aw->ln_b[l] : nul, nul, 0, d.Ha, d.Ha, 1.4426950408889634f);
    add_vec(pk, aw ? aw->ln_g[l] : nul, nul, 0, d.Ha, d.Ha);
    L.ep_floats = 3u * d.Ha;
    lp.blayers.push_back(L);
  }
  // ====================================================================== LayerNorm layers split over a cluster
  auto make_split = [&](int f, std::vector<LLayerH>& fl, std::vector<LLayerH>& bl) {
    constexpr float kLog2e = 1.4426950408889634f;
    const int ns = d.Ha / f;
    fl = lp.layers;
    bl = lp.blayers;
    for (LLayerH& L : fl) {
      if (L.type != LE_LN_ELU) continue;
      const int l = L.layer;
      const float* lw = aw ? aw->lin_w[l] : nul;
      L.NT = f; L.n = ns; L.dbuf = 1; L.kc = kc_for(ns); L.csplit = f;
      L.w_off = pk.data_bytes;
      for (int j = 0; j < L.NT; ++j) {
        const size_t t0 = pk.data_bytes;
        RowSpec rs(j * ns, ns, ns);
        if (l == 0) {
          L.seg[0] = {F_H, Dp / 16, ns, 0, 0, 1, 0u};
          l_add_pack(lp, rs, lw, d.F, 0, d.D, Dp);
          L.seg[1] = {F_S, Sip / 16, ns, 0, 0, 0, (uint32_t)(pk.data_bytes - t0)};
          l_add_pack(lp, rs, lw, d.F, d.D, d.S_in, Sip);
        } else {
          L.seg[0] = {F_Y0 + l - 1, d.Ha / 16, ns, 0, 0, 1, 0u};
          l_add_pack(lp, rs, lw, d.Ha, 0, d.Ha, d.Ha);
        }
        L.w_tile_bytes = (uint32_t)(pk.data_bytes - t0);
        const uint32_t at = add_vec(pk, aw ? aw->lin_b[l] : nul, nul, j * ns, ns, ns);
        add_vec(pk, aw ? aw->ln_g[l] : nul, nul, j * ns, ns, ns, kLog2e);
        add_vec(pk, aw ? aw->ln_b[l] : nul, nul, j * ns, ns, ns, kLog2e);
        if (j == 0) L.ep_off = at;
      }
      L.ep_floats = 3u * ns;
    }
    for (LLayerH& L : bl) {
      if (L.type != LB_LN) continue;
      const int l = L.layer;
      L.NT = f; L.n = ns; L.dbuf = 1; L.kc = kc_for(ns); L.csplit = f;
      L.w_off = pk.data_bytes;
      for (int j = 0; j < L.NT; ++j) {
        const size_t t0 = pk.data_bytes;
        RowSpec rs(j * ns, ns, ns);
        if (l == d.La - 1) {
          L.seg[0] = {G_GMU, 1, ns, 0, 0, 1, 0u};
          l_add_pack_t(lp, rs, aw ? aw->mu_w : nul, d.Ha, {{0, d.A, 16}});
        } else {
          L.seg[0] = {G_GP0 + l + 1, d.Ha / 16, ns, 0, 0, 1, 0u};
          l_add_pack_t(lp, rs, aw ? aw->lin_w[l + 1] : nul, d.Ha, {{0, d.Ha, d.Ha}});
        }
        L.w_tile_bytes = (uint32_t)(pk.data_bytes - t0);
        const uint32_t at = add_vec(pk, aw ? aw->ln_g[l] : nul, nul, j * ns, ns, ns, kLog2e);
        add_vec(pk, aw ? aw->ln_b[l] : nul, nul, j * ns, ns, ns, kLog2e);
        add_vec(pk, aw ? aw->ln_g[l] : nul, nul, j * ns, ns, ns);
        if (j == 0) L.ep_off = at;
      }
      L.ep_floats = 3u * ns;
    }
  };
  if (d.Ha % (32 * kLMaxSplit) == 0) { lp.csplit = kLMaxSplit; make_split(kLMaxSplit, lp.layers_s, lp.blayers_s); }
  if (d.Ha > 256 && d.Ha % 64 == 0) { lp.cpair = 2; make_split(2, lp.layers_p, lp.blayers_p); }
  for (const std::vector<LLayerH>* v : {&lp.layers, &lp.blayers, &lp.layers_s, &lp.blayers_s, &lp.layers_p, &lp.blayers_p})
    for (const LLayerH& L : *v) {
      if (L.ep_floats * 4 > kLParamBytes || (L.ep_floats & 3)) return lp;
      for (int s = 0; s < L.nseg; ++s)
        if ((uint32_t)L.kc * (4096u + (uint32_t)L.seg[s].n * 32u) > kLSlotBytes) return lp;
    }
  auto al = [](size_t x) { return (x + 255) / 256 * 256; };
  pk.off_stage_tab = 256;
  pk.off_job_tab = 512;
  pk.off_packops = 768;
  pk.off_eparams = al(pk.off_packops + pk.pack_ops.size() * sizeof(PackOp) + pk.vec_ops.size() * sizeof(VecOp));
  pk.off_data = al(pk.off_eparams + pk.eparams_floats * sizeof(float));
  pk.total_bytes = al(pk.off_data + pk.data_bytes);
  pk.ok = true;
  lp.ok = true;
  return lp;
}

// ---- memory layout.  Every panel buffer is [steps][MT][cols/8][2 KB]; the forward fields live in the workspace
// (one step, the h / s panels two) or, when recording for the BPTT, in `saved` with one entry per step.
struct Region { size_t off = 0; uint32_t tile_bytes = 0; size_t step_bytes = 0; int steps = 0; int where = 0; };   // where: 0 ws, 1 saved
struct LMem {
  Region r[BUF_COUNT];
  size_t carry_h = 0, carry_s = 0;   // fp32 carry panels in the backward workspace
  size_t idx_pri = 0, idx_post = 0;  // filtering: int32 [B][T-1][C] draws of the two heads (discrete)
  size_t ws_bytes = 0, saved_bytes = 0;
};
constexpr size_t kLTraceBytes = 8192;
// mode 0: forward, nothing recorded; 1: forward, recording; 2: backward (forward fields resolve into `saved`)
static LMem tcl_layout(const LPlan& lp, const Dims& d, int mode) {
  LMem m;
  const int MT = (d.B + 127) / 128, T = d.T;
  size_t off[2] = {kLTraceBytes, 0};
  auto place = [&](int buf, int steps, int where, size_t row_bytes = 0) {
    Region& r = m.r[buf];
    r.tile_bytes = row_bytes ? (uint32_t)(128 * row_bytes) : (uint32_t)(buf_width(lp, buf) / 8) * kLTileChunk;
    r.step_bytes = (size_t)MT * r.tile_bytes;
    r.steps = steps; r.where = where;
    r.off = off[where];
    off[where] += ((size_t)steps * r.step_bytes + 1023) / 1024 * 1024;
  };
  const bool rec = mode != 0;
  const int sv = rec ? 1 : 0;
  place(F_H, rec ? T + 1 : 2, sv);
  place(F_S, rec ? T + 1 : 1, sv);
  if (mode != 2) place(F_A, 1, 0);
  place(F_X, rec ? T : 1, sv);
  place(F_HID, rec ? T : 1, sv);
  for (int l = 0; l < d.La; ++l) place(F_Y0 + l, rec ? T : 1, sv);
  if (rec) {
    for (int f : {F_GR, F_GZ, F_GN, F_GHN}) place(f, T, 1);
    for (int l = 0; l < d.La; ++l) { place(F_XH0 + l, T, 1); place(F_RSTD0 + l, T, 1, sizeof(float)); }
  }
  if (mode == 2) {
    place(G_GO, 1, 0); place(G_GHID, 1, 0); place(G_GATES_I, 1, 0); place(G_GX, 1, 0);
    m.r[G_GATES_H] = m.r[G_GATES_I];
    m.r[G_GATES_H].off += (size_t)(lp.Dp / 8) * kLTileChunk;
    place(G_GMU, T, 0);
    for (int l = 0; l < d.La; ++l) { place(G_GP0 + l, T, 0); place(G_GU0 + l, T, 0); }
    m.carry_h = off[0]; off[0] += ((size_t)MT * 128 * lp.Dp * 4 + 1023) / 1024 * 1024;
    m.carry_s = off[0]; off[0] += ((size_t)MT * 128 * lp.Sip * 4 + 1023) / 1024 * 1024;
  }
  m.r[F_HN] = m.r[F_H]; m.r[F_SN] = m.r[F_S];
  m.ws_bytes = off[0] + 1024;
  m.saved_bytes = off[1] + 1024;
  return m;
}

bool tcl_supported(const Dims& d) { return tcl_make_plan(d, nullptr, nullptr).ok; }
bool tcl_train_supported(const Dims& d) { return tcl_supported(d); }
size_t tcl_packed_bytes(const Dims& d) {
  LPlan lp = tcl_make_plan(d, nullptr, nullptr);
  return lp.ok ? lp.pk.total_bytes : 0;
}
size_t tcl_workspace_bytes(const Dims& d) {
  LPlan lp = tcl_make_plan(d, nullptr, nullptr);
  if (!lp.ok) return 0;
  const int mode = (d.flags & WMRL_FLAG_BACKWARD) ? 2 : ((d.flags & WMRL_FLAG_SAVE_FOR_BWD) ? 1 : 0);
  return tcl_layout(lp, d, mode).ws_bytes;
}
size_t tcl_saved_bytes(const Dims& d) {
  LPlan lp = tcl_make_plan(d, nullptr, nullptr);
  return lp.ok ? tcl_layout(lp, d, 1).saved_bytes : 0;
}
int tcl_pack_weights(const Dims& d, const wmrl_rssm_params* w, const wmrl_actor_params* aw, void* packed,
                     cudaStream_t st) {
  LPlan lp = tcl_make_plan(d, w, aw);
  WMRL_REQUIRE(lp.ok, "pack_weights: shape not supported by the layer-wise tcgen05 path");
  return tc_pack_plan(lp.pk, (uint8_t*)packed, st);
}

static int g_max_clusters[64][kLMaxSplit + 1] = {};   // [device][f]: co-resident clusters of f layer-kernel CTAs (tcl_device_setup)
static int tcl_device_setup(int& nsm) {
  static std::mutex mu;
  static bool attr_set[64] = {false};
  static int sm_cached[64] = {0};
  int dev = 0;
  WMRL_CUDA_OK(cudaGetDevice(&dev));
  WMRL_REQUIRE(dev >= 0 && dev < 64, "device ordinal out of range");
  std::lock_guard<std::mutex> lk(mu);
  if (!attr_set[dev]) {
    for (int ty = LE_LN_ELU; ty <= kLTypeLast; ++ty)
      for (int S : {4, 8, 16, 32}) {
        if (ty != LE_SAMPLE_DISC && S != 4) continue;
        WMRL_CUDA_OK(cudaFuncSetAttribute(layer_kernel_for(ty, S), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)kLSmemBytes));
      }
    WMRL_CUDA_OK(cudaFuncSetAttribute(tcl_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWg2Smem));
    WMRL_CUDA_OK(cudaDeviceGetAttribute(&sm_cached[dev], cudaDevAttrMultiProcessorCount, dev));
    for (int f : {2, kLMaxSplit}) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3((unsigned)(sm_cached[dev] / f * f));
      cfg.blockDim = dim3(kLThreads);
      cfg.dynamicSmemBytes = kLSmemBytes;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = (unsigned)f; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      int nc = 0;
      if (cudaOccupancyMaxActiveClusters(&nc, layer_kernel_for(LE_LN_ELU, 4), &cfg) != cudaSuccess) { nc = 0; cudaGetLastError(); }
      g_max_clusters[dev][f] = nc;
    }
    attr_set[dev] = true;
  }
  nsm = sm_cached[dev];
  return 0;
}

static int tcl_max_clusters(int f) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || f < 2 || f > kLMaxSplit) return 0;
  return g_max_clusters[dev][f];
}
// which tiling of the LayerNorm layers a call with MT row tiles uses: kLMaxSplit-CTA clusters when every row tile's
// cluster is co-resident (few row tiles), CTA pairs for wide blocks otherwise, 0 = one CTA per row tile
static int tcl_split_choice(int csplit, int cpair, int MT) {
  // WMRL_LN_SPLIT = 0: one CTA per row tile; 2 / 4: force that cluster form wherever it is possible (tests); unset: automatic
  const char* e = getenv("WMRL_LN_SPLIT");
  if (e && e[0] == '0') return 0;
  if (e && e[0] == '2') return (cpair > 1 && tcl_max_clusters(cpair) > 0) ? cpair : 0;
  if (e && e[0] == '4') return (csplit > 1 && MT <= tcl_max_clusters(csplit)) ? csplit : 0;
  if (csplit > 1 && MT <= tcl_max_clusters(csplit)) return csplit;
  // pairs re-read the block's input panel from both CTAs: a win while that panel is L2-resident (<= 256 row tiles =
  // 32 MB at 512 columns; measured: configs[4] shard 20.2 -> 19.1 ms forward), a loss once it streams from HBM
  // (983 040-row heads: 5.9 -> 6.3 ms)
  if (cpair > 1 && MT <= 256 && tcl_max_clusters(cpair) > 0) return cpair;
  return 0;
}
struct LCall {            // what the launch helper needs to resolve a layer description into kernel arguments
  const LPlan* lp; const LMem* m; const Dims* d;
  uint8_t* ws; uint8_t* saved;
  const uint8_t* data; const float* ep;
  int MT, nsm;
  cudaStream_t st;
  uint8_t* at(int buf, int step) const {
    const Region& r = m->r[buf];
    return (r.where ? saved : ws) + r.off + (size_t)step * r.step_bytes;
  }
};
static int tcl_launch(const LCall& c, const LLayerH& L, LayerArgs& a, const int* seg_step, int out_step) {
  const LPlan& lp = *c.lp;
  const Dims& d = *c.d;
  a.nseg = L.nseg;
  for (int s = 0; s < L.nseg; ++s) {
    const int buf = L.seg[s].buf;
    a.seg[s].a = c.at(buf, seg_step[s]);
    a.seg[s].a_tile_bytes = c.m->r[buf].tile_bytes;
    a.seg[s].w_off = L.seg[s].w_off;
    a.seg[s].k16 = (uint16_t)L.seg[s].k16; a.seg[s].n = (uint16_t)L.seg[s].n;
    a.seg[s].tmem_col = (uint16_t)L.seg[s].tmem_col; a.seg[s].split = (uint16_t)L.seg[s].split;
    a.seg[s].fresh = (uint16_t)L.seg[s].fresh;
  }
  a.w = c.data + L.w_off; a.w_tile_bytes = L.w_tile_bytes;
  a.ep = c.ep + L.ep_off; a.ep_floats = L.ep_floats;
  a.kc = L.kc; a.dbuf = L.dbuf; a.type = L.type; a.n = L.n; a.NT = L.NT; a.MT = c.MT;
  a.csplit = L.csplit; a.ln_width = L.csplit > 1 ? L.n * L.NT : L.n;
  a.B = d.B; a.H = d.T; a.D = d.D; a.S = d.S; a.C = d.C; a.A = d.A; a.Sp = lp.Sp; a.Dp = lp.Dp; a.Sip = lp.Sip; a.E = d.E;
  a.inv_action_scale = 1.f / d.action_scale;
  a.out = c.at(L.out, out_step); a.out_tile_bytes = c.m->r[L.out].tile_bytes; a.out_cols = buf_width(lp, L.out);
  const int items = c.MT * L.NT;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)std::min(items, c.nsm));
  cfg.blockDim = dim3(kLThreads);
  cfg.dynamicSmemBytes = kLSmemBytes;
  cfg.stream = c.st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (L.csplit > 1) {   // clusters of csplit CTAs, all co-resident (single wave): persistent over the row tiles
    cfg.gridDim = dim3((unsigned)(std::min(c.MT, tcl_max_clusters(L.csplit)) * L.csplit));
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = (unsigned)L.csplit; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
    cfg.numAttrs = 2;
  }
  WMRL_CUDA_OK(cudaLaunchKernelEx(&cfg, layer_kernel_for(L.type, d.S), a));
  return 0;
}
int tcl_imagine_fwd(const Dims& d, const void* packed, const float* deter0, const float* stoch0, const float* noise,
                    float* deter_seq, float* stoch_seq, float* dist_a, float* dist_b, float* actions, int32_t* idx,
                    void* saved, void* ws, cudaStream_t st) {
  LPlan lp = tcl_make_plan(d, nullptr, nullptr);
  WMRL_REQUIRE(lp.ok, "imagine_fwd: shape not supported by the layer-wise tcgen05 path");
  const bool rec = saved != nullptr;
  const LMem m = tcl_layout(lp, d, rec ? 1 : 0);
  LCall c{&lp, &m, &d, (uint8_t*)ws, (uint8_t*)saved, nullptr, nullptr, (d.B + 127) / 128, 148, st};
  if (int r = tcl_device_setup(c.nsm)) return r;
  const uint8_t* img = (const uint8_t*)packed;
  c.data = img + lp.pk.off_data;
  c.ep = (const float*)(img + lp.pk.off_eparams);
  long long* trace_base = (long long*)ws;   // the first 8 KB of the workspace
  tcl_init_kernel<<<dim3(8, c.MT), 256, 0, st>>>(deter0, stoch0, deter_seq, stoch_seq, c.at(F_H, 0), c.at(F_S, 0), d.B, d.T,
                                                 d.D, d.S_in, lp.Dp, lp.Sip);
  WMRL_CUDA_OK(cudaGetLastError());
  const int sc = tcl_split_choice(lp.csplit, lp.cpair, c.MT);
  const std::vector<LLayerH>& fl = sc == 0 ? lp.layers : (sc == 2 ? lp.layers_p : lp.layers_s);
  for (int t = 0; t < d.T; ++t) {
    // step index of each field: recorded fields are indexed by t (h, s: t -> t + 1); otherwise one entry (h: parity)
    auto step_of = [&](int buf) {
      if (buf == F_H) return rec ? t : (t & 1);
      if (buf == F_HN) return rec ? t + 1 : ((t + 1) & 1);
      if (buf == F_S) return rec ? t : 0;
      if (buf == F_SN) return rec ? t + 1 : 0;
      if (buf == F_A) return 0;
      return rec ? t : 0;
    };
    for (const LLayerH& L : fl) {
      LayerArgs a{};
      a.t = t;
      a.deter_seq = deter_seq; a.stoch_seq = stoch_seq; a.dist_a = dist_a; a.dist_b = dist_b; a.actions = actions;
      a.idx = idx; a.noise = noise;
      a.save = rec ? 1 : 0;
      if (rec && L.type == LE_LN_ELU) {
        a.aux[0] = c.at(F_XH0 + L.layer, t); a.aux_tile_bytes = m.r[F_XH0 + L.layer].tile_bytes;
        a.f0 = (float*)c.at(F_RSTD0 + L.layer, t);
      }
      if (rec && L.type == LE_GRU) {
        a.aux[0] = c.at(F_GR, t); a.aux[1] = c.at(F_GZ, t); a.aux[2] = c.at(F_GN, t); a.aux[3] = c.at(F_GHN, t);
        a.aux_tile_bytes = m.r[F_GR].tile_bytes;
      }
      a.trace = nullptr;
      if ((d.flags & WMRL_FLAG_TRACE) && t == std::min(1, d.T - 1)) a.trace = trace_base + 32 * (&L - &fl[0]);
      int ss[3];
      for (int s = 0; s < L.nseg; ++s) ss[s] = step_of(L.seg[s].buf);
      if (int r = tcl_launch(c, L, a, ss, step_of(L.out))) return r;
    }
  }
  if (d.variant == WMRL_DISCRETE) {
    tcl_onehot_kernel<<<4 * c.nsm, 256, 0, st>>>(idx, stoch_seq, d.B, d.T, d.C, d.S);
    WMRL_CUDA_OK(cudaGetLastError());
  }
  return 0;
}

// BPTT over the recorded rollout (autograd of rssm.py:123-140 with the world model frozen and the actor input
// detached).  Per step, backwards in time: sample backward (SIMT) -> 5 dgrad GEMMs through the dynamics with fused
// ELU' / GRU-gate / tanh-squash backward -> La dgrad GEMMs through the actor with fused LayerNorm+ELU backward; then the
// actor weight gradients as big-K tcgen05 GEMMs over the recorded panels and column sums for the bias / affine terms.
int tcl_imagine_bwd(const Dims& d, const void* packed, const float* noise, const float* deter_seq, const float* dist_a,
                    const float* dist_b, const float* actions, const void* saved, const float* g_deter_seq,
                    const float* g_stoch_seq, const float* g_dist_a, const float* g_dist_b, const wmrl_actor_grads* gaw,
                    float* g_deter0, float* g_stoch0, void* ws, cudaStream_t st) {
  LPlan lp = tcl_make_plan(d, nullptr, nullptr);
  WMRL_REQUIRE(lp.ok, "imagine_bwd: shape not supported by the layer-wise tcgen05 path");
  const LMem m = tcl_layout(lp, d, 2);
  LCall c{&lp, &m, &d, (uint8_t*)ws, (uint8_t*)const_cast<void*>(saved), nullptr, nullptr, (d.B + 127) / 128, 148, st};
  if (int r = tcl_device_setup(c.nsm)) return r;
  const uint8_t* img = (const uint8_t*)packed;
  c.data = img + lp.pk.off_data;
  c.ep = (const float*)(img + lp.pk.off_eparams);
  const int T = d.T, MT = c.MT;
  const bool disc = d.variant == WMRL_DISCRETE;
  float* carry_h = (float*)((uint8_t*)ws + m.carry_h);
  float* carry_s = (float*)((uint8_t*)ws + m.carry_s);
  {
    const long long n = (long long)MT * 128 * lp.Dp;
    tcl_carry_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(carry_h, g_deter_seq, d.B, T, d.D, lp.Dp, n);
    WMRL_CUDA_OK(cudaGetLastError());
  }
  const int sc = tcl_split_choice(lp.csplit, lp.cpair, MT);
  const std::vector<LLayerH>& bl = sc == 0 ? lp.blayers : (sc == 2 ? lp.blayers_p : lp.blayers_s);
  for (int t = T - 1; t >= 0; --t) {
    // dL/d(head output) of the step that produced state t+1
    {
      SbArgs s{};
      s.carry_s = (t < T - 1) ? carry_s : nullptr; s.g_stoch = g_stoch_seq; s.g_da = g_dist_a; s.g_db = g_dist_b;
      s.dist_a = dist_a; s.dist_b = dist_b; s.noise = noise; s.out = c.at(G_GO, 0);
      s.B = d.B; s.H = T; s.t = t; s.S = d.S; s.C = d.C; s.Sp = lp.Sp; s.Sip = lp.Sip; s.Pp = lp.Pp; s.MT = MT;
      if (disc) {
        const long long nw = (long long)MT * 128 * ((d.C + kSbGroups - 1) / kSbGroups);
        tcl_sample_bwd_disc_kernel<<<(unsigned)((nw * 32 + 255) / 256), 256, 0, st>>>(s);
      } else {
        const long long n = (long long)MT * 128 * (lp.Sp / 8);
        tcl_sample_bwd_cont_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s);
      }
      WMRL_CUDA_OK(cudaGetLastError());
    }
    for (const LLayerH& L : bl) {
      LayerArgs a{};
      a.t = t;
      a.deter_seq = const_cast<float*>(deter_seq); a.actions = const_cast<float*>(actions);
      int out_step = 0;
      switch (L.type) {
        case LB_ELU:
          a.aux[0] = c.at(L.layer, t); a.aux_tile_bytes = m.r[L.layer].tile_bytes;
          break;
        case LB_GRU:
          a.aux[0] = c.at(F_GR, t); a.aux[1] = c.at(F_GZ, t); a.aux[2] = c.at(F_GN, t); a.aux[3] = c.at(F_GHN, t);
          a.aux[4] = c.at(F_H, t);
          a.aux_tile_bytes = m.r[F_GR].tile_bytes;
          a.f0 = carry_h;
          break;
        case LB_CARRY:
          a.f0 = carry_h; a.f1 = const_cast<float*>(g_deter_seq); a.f2 = g_deter0;
          break;
        case LB_EMB:
          a.f0 = carry_s; a.f1 = const_cast<float*>(g_stoch_seq); a.f2 = g_stoch0;
          out_step = t;
          break;
        case LB_LN:
          a.aux[0] = c.at(F_XH0 + L.layer, t); a.aux[1] = c.at(G_GU0 + L.layer, t); a.aux[2] = c.at(F_Y0 + L.layer, t);
          a.aux_tile_bytes = m.r[F_XH0 + L.layer].tile_bytes;
          a.f0 = (float*)c.at(F_RSTD0 + L.layer, t);
          out_step = t;
          break;
      }
      int ss[3] = {0, 0, 0};
      if (L.seg[0].buf == G_GMU || L.seg[0].buf >= G_GP0) ss[0] = t;
      a.trace = nullptr;
      if ((d.flags & WMRL_FLAG_TRACE) && t == std::max(0, T - 2)) a.trace = (long long*)ws + 32 * (&L - &bl[0]);
      if (int r = tcl_launch(c, L, a, ss, out_step)) return r;
    }
  }
  // ---- weight gradients: dW[out][in] += sum over (step, row tile) of G^T X, both operands straight from the panels
  const int nblk = T * MT;
  auto wgrad = [&](int abuf, int m_cols, int m_valid, int bbuf, int n_cols, int n_valid, float* dW, int ldw) -> int {
    if (!dW) return 0;
    Wg2Params q{};
    q.a_base = c.at(abuf, 0); q.a_blk = m.r[abuf].tile_bytes; q.b_base = c.at(bbuf, 0); q.b_blk = m.r[bbuf].tile_bytes;
    q.nblocks = nblk; q.m_cols = m_cols; q.m_valid = m_valid; q.n_cols = n_cols; q.n_valid = n_valid;
    q.dW = dW; q.ldw = ldw;
    q.mtiles = (m_cols + 127) / 128; q.ntiles = (n_cols + 255) / 256;
    const int tl = q.mtiles * q.ntiles;
    q.splits = std::max(1, std::min(nblk, c.nsm / tl));
    tcl_wgrad_kernel<<<dim3(tl, q.splits), kWg2Threads, kWg2Smem, st>>>(q);
    WMRL_CUDA_OK(cudaGetLastError());
    return 0;
  };
  auto colsum = [&](int xbuf, int ybuf, int cols, int valid, float* out) -> int {
    if (!out) return 0;
    Cs2Params q{};
    q.x_base = c.at(xbuf, 0); q.x_blk = m.r[xbuf].tile_bytes;
    q.y_base = ybuf >= 0 ? c.at(ybuf, 0) : nullptr; q.y_blk = ybuf >= 0 ? m.r[ybuf].tile_bytes : 0;
    q.nblocks = nblk; q.valid = valid; q.out = out;
    const int chunks = cols / 8;
    const int ysplit = std::max(1, std::min((nblk + kCs2Warps - 1) / kCs2Warps, (8 * c.nsm + chunks - 1) / chunks));
    tcl_colsum_kernel<<<dim3(chunks, ysplit), kCs2Warps * 32, 0, st>>>(q);
    WMRL_CUDA_OK(cudaGetLastError());
    return 0;
  };
  auto colsum3 = [&](int xbuf, int ybuf, int zbuf, int cols, float* o_x, float* o_xy, float* o_z) -> int {
    Cs2Params q{};
    q.x_base = c.at(xbuf, 0); q.x_blk = m.r[xbuf].tile_bytes; q.y_base = c.at(ybuf, 0); q.y_blk = m.r[ybuf].tile_bytes;
    q.z_base = c.at(zbuf, 0); q.z_blk = m.r[zbuf].tile_bytes;
    q.nblocks = nblk; q.valid = cols; q.out = o_x; q.out2 = o_xy; q.out3 = o_z;
    const int chunks = cols / 8;
    const int ysplit = std::max(1, std::min((nblk + kCs2Warps - 1) / kCs2Warps, (8 * c.nsm + chunks - 1) / chunks));
    tcl_colsum_kernel<<<dim3(chunks, ysplit), kCs2Warps * 32, 0, st>>>(q);
    WMRL_CUDA_OK(cudaGetLastError());
    return 0;
  };
  for (int l = 0; l < d.La; ++l) {
    if (l == 0) {
      if (int r = wgrad(F_H, lp.Dp, d.D, G_GP0, d.Ha, d.Ha, gaw->lin_w[0], d.F)) return r;
      if (int r = wgrad(F_S, lp.Sip, d.S_in, G_GP0, d.Ha, d.Ha, gaw->lin_w[0] ? gaw->lin_w[0] + d.D : nullptr, d.F)) return r;
    } else {
      if (int r = wgrad(F_Y0 + l - 1, d.Ha, d.Ha, G_GP0 + l, d.Ha, d.Ha, gaw->lin_w[l], d.Ha)) return r;
    }
    if (gaw->lin_b[l] && gaw->ln_b[l] && gaw->ln_g[l]) {
      if (int r = colsum3(G_GU0 + l, F_XH0 + l, G_GP0 + l, d.Ha, gaw->ln_b[l], gaw->ln_g[l], gaw->lin_b[l])) return r;
    } else {
      if (int r = colsum(G_GP0 + l, -1, d.Ha, d.Ha, gaw->lin_b[l])) return r;
      if (int r = colsum(G_GU0 + l, -1, d.Ha, d.Ha, gaw->ln_b[l])) return r;
      if (int r = colsum(G_GU0 + l, F_XH0 + l, d.Ha, d.Ha, gaw->ln_g[l])) return r;
    }
  }
  if (int r = wgrad(F_Y0 + d.La - 1, d.Ha, d.Ha, G_GMU, 16, d.A, gaw->mu_w, d.Ha)) return r;
  if (int r = colsum(G_GMU, -1, 16, d.A, gaw->mu_b)) return r;
  return 0;
}

// ==========================================================================
// Posterior filtering on the chain (RSSMBase.observe_rollout, rssm.py:93-121), forward: per step
//   prior_{t+1} = prior(a_t, post_t)  (embed -> GRU -> prior hidden -> prior head + sample)
//   post_{t+1}  = posterior(e_{t+1}, prior_{t+1})  ([h' | e] -> hidden -> head + sample; its sample feeds step t+1)
// obs_embeds / actions are converted once into per-step bf16 panels.
// ==========================================================================
static LPlan tcl_make_observe_plan(const Dims& d, const wmrl_rssm_params* w) {
  LPlan lp;
  const bool disc = d.variant == WMRL_DISCRETE;
  if (d.A > 16) return lp;
  if (disc && !(d.S == 4 || d.S == 8 || d.S == 16 || d.S == 32)) return lp;
  const int Dp = ceil16(d.D), Sip = ceil16(d.S_in), Sp = ceil16(d.S), Hdp = ceil16(d.Hd), Ep = ceil16(d.E);
  if (!disc && 2 * Sp > 256) return lp;
  lp.Dp = Dp; lp.Sip = Sip; lp.Sp = Sp; lp.Hdp = Hdp; lp.Ha = 32; lp.Pp = disc ? Sip : 2 * Sp; lp.Ep = Ep;
  const float* nul = nullptr;
  Plan& pk = lp.pk;
  auto kc_for = [&](int nmax) { int kc = (int)(kLSlotBytes / (4096u + (uint32_t)nmax * 32u)); return std::max(1, std::min(kc, 8)); };
  auto tiles = [&](int total, int cap, int unit, int& nt, int& wt) {
    nt = (total + cap - 1) / cap;
    wt = ceil_to((total + nt - 1) / nt, unit);
  };
  // ---- state-action embed: [post s_t | a_t] -> Linear -> ELU
  {
    LLayerH L;
    L.type = LE_BIAS_ELU; L.dbuf = 1; L.out = F_X; L.nseg = 2;
    tiles(Dp, 256, 16, L.NT, L.n);
    L.kc = kc_for(L.n);
    L.w_off = pk.data_bytes;
    const int Ke = d.S_in + d.A;
    for (int j = 0; j < L.NT; ++j) {
      const size_t t0 = pk.data_bytes;
      RowSpec rs(j * L.n, d.D - j * L.n, L.n);
      L.seg[0] = {F_S, Sip / 16, L.n, 0, 0, 1, 0u};
      l_add_pack(lp, rs, w ? w->embed_w : nul, Ke, 0, d.S_in, Sip);
      L.seg[1] = {F_ACT, 1, L.n, 0, 0, 0, (uint32_t)(pk.data_bytes - t0)};
      l_add_pack(lp, rs, w ? w->embed_w : nul, Ke, d.S_in, d.A, 16);
      L.w_tile_bytes = (uint32_t)(pk.data_bytes - t0);
      const uint32_t at = add_vec(pk, w ? w->embed_b : nul, nul, j * L.n, std::max(0, std::min(L.n, d.D - j * L.n)), L.n);
      if (j == 0) L.ep_off = at;
    }
    L.ep_floats = (uint32_t)L.n;
    lp.layers.push_back(L);
  }
  // ---- GRU
  {
    LLayerH L;
    constexpr int cw = kGruTile;
    L.type = LE_GRU; L.dbuf = 1; L.out = F_HN; L.nseg = 2; L.n = 4 * cw;
    L.NT = (Dp + cw - 1) / cw;
    L.kc = kc_for(3 * cw);
    L.w_off = pk.data_bytes;
    for (int j = 0; j < L.NT; ++j) {
      const size_t t0 = pk.data_bytes;
      const int c0 = j * cw, v = std::max(0, std::min(cw, d.D - c0));
      RowSpec rx;
      rx.add(2 * d.D + c0, v, cw); rx.add(c0, v, cw); rx.add(d.D + c0, v, cw);
      L.seg[0] = {F_X, Dp / 16, 3 * cw, 0, 0, 1, 0u};
      l_add_pack(lp, rx, w ? w->w_ih : nul, d.D, 0, d.D, Dp);
      RowSpec rh;
      rh.add(c0, v, cw); rh.add(d.D + c0, v, cw); rh.add(2 * d.D + c0, v, cw);
      L.seg[1] = {F_H, Dp / 16, 3 * cw, cw, 2 * cw, 0, (uint32_t)(pk.data_bytes - t0)};
      l_add_pack(lp, rh, w ? w->w_hh : nul, d.D, 0, d.D, Dp);
      L.w_tile_bytes = (uint32_t)(pk.data_bytes - t0);
      const uint32_t at = add_vec(pk, w ? w->b_ih : nul, w ? w->b_hh : nul, c0, v, cw);
      add_vec(pk, w ? w->b_ih : nul, w ? w->b_hh : nul, d.D + c0, v, cw);
      add_vec(pk, w ? w->b_ih : nul, nul, 2 * d.D + c0, v, cw);
      add_vec(pk, w ? w->b_hh : nul, nul, 2 * d.D + c0, v, cw);
      if (j == 0) L.ep_off = at;
    }
    L.ep_floats = 4 * cw;
    lp.layers.push_back(L);
  }
  // ---- hidden layers of the two heads: prior h' -> ELU; posterior [h' | e_{t+1}] -> ELU
  for (int post = 0; post < 2; ++post) {
    LLayerH L;
    L.type = LE_BIAS_ELU; L.dbuf = 1; L.out = post ? F_HIDQ : F_HID; L.nseg = post ? 2 : 1;
    tiles(Hdp, 256, 16, L.NT, L.n);
    L.kc = kc_for(L.n);
    L.w_off = pk.data_bytes;
    const float* W0 = w ? (post ? w->post0_w : w->prior0_w) : nul;
    const float* b0 = w ? (post ? w->post0_b : w->prior0_b) : nul;
    const int ld = post ? d.D + d.E : d.D;
    for (int j = 0; j < L.NT; ++j) {
      const size_t t0 = pk.data_bytes;
      RowSpec rs(j * L.n, d.Hd - j * L.n, L.n);
      L.seg[0] = {F_HN, Dp / 16, L.n, 0, 0, 1, 0u};
      l_add_pack(lp, rs, W0, ld, 0, d.D, Dp);
      if (post) {
        L.seg[1] = {F_OBS, Ep / 16, L.n, 0, 0, 0, (uint32_t)(pk.data_bytes - t0)};
        l_add_pack(lp, rs, W0, ld, d.D, d.E, Ep);
      }
      L.w_tile_bytes = (uint32_t)(pk.data_bytes - t0);
      const uint32_t at = add_vec(pk, b0, nul, j * L.n, std::max(0, std::min(L.n, d.Hd - j * L.n)), L.n);
      if (j == 0) L.ep_off = at;
    }
    L.ep_floats = (uint32_t)L.n;
    lp.layers.push_back(L);
  }
  // ---- heads + samples (layer = 0 prior, 1 posterior)
  for (int post = 0; post < 2; ++post) {
    LLayerH L;
    L.dbuf = 1; L.out = post ? F_SN : F_SPRI; L.nseg = 1; L.layer = post;
    L.w_off = pk.data_bytes;
    const float* W2 = w ? (post ? w->post2_w : w->prior2_w) : nul;
    const float* b2 = w ? (post ? w->post2_b : w->prior2_b) : nul;
    const int in = post ? F_HIDQ : F_HID;
    if (disc) {
      L.type = LE_SAMPLE_DISC;
      tiles(ceil_to(d.P, 32), 256, 32, L.NT, L.n);
      for (int j = 0; j < L.NT; ++j) {
        const size_t t0 = pk.data_bytes;
        L.seg[0] = {in, Hdp / 16, L.n, 0, 0, 1, 0u};
        l_add_pack(lp, RowSpec(j * L.n, d.P - j * L.n, L.n), W2, d.Hd, 0, d.Hd, Hdp);
        L.w_tile_bytes = (uint32_t)(pk.data_bytes - t0);
        const uint32_t at = add_vec(pk, b2, nul, j * L.n, std::max(0, std::min(L.n, d.P - j * L.n)), L.n);
        if (j == 0) L.ep_off = at;
      }
      L.ep_floats = (uint32_t)L.n;
    } else {
      L.type = LE_SAMPLE_CONT; L.NT = 1; L.n = 2 * Sp;
      RowSpec rs;
      rs.add(0, d.S, Sp); rs.add(d.S, d.S, Sp);
      L.seg[0] = {in, Hdp / 16, L.n, 0, 0, 1, 0u};
      l_add_pack(lp, rs, W2, d.Hd, 0, d.Hd, Hdp);
      L.w_tile_bytes = (uint32_t)(pk.data_bytes - L.w_off);
      L.ep_off = add_vec(pk, b2, nul, 0, d.S, Sp);
      add_vec(pk, b2, nul, d.S, d.S, Sp);
      L.ep_floats = 2u * Sp;
    }
    L.kc = kc_for(L.n);
    lp.layers.push_back(L);
  }
  // execution order of a step: embed, GRU, prior hidden, prior head, posterior hidden, posterior head
  std::swap(lp.layers[3], lp.layers[4]);
  // ====================================================================== backward chain (full backward of filtering:
  // dgrad through both heads, the GRU and the embed; weight gradients after the scan, world_model.py:120-163)
  {
    const int Pp = lp.Pp;
    const uint32_t zero16 = add_vec(pk, nul, nul, 0, 0, 16);
    auto head_dgrad = [&](int in_buf, const float* W2, int out_buf, int stash) {   // g_hid = g_out W2 (x ELU'(hid))
      LLayerH L;
      L.type = LB_ELU; L.dbuf = 1; L.out = out_buf; L.nseg = 1; L.layer = stash;
      tiles(Hdp, 256, 16, L.NT, L.n);
      L.kc = kc_for(L.n);
      L.w_off = pk.data_bytes;
      for (int j = 0; j < L.NT; ++j) {
        const size_t t0 = pk.data_bytes;
        L.seg[0] = {in_buf, Pp / 16, L.n, 0, 0, 1, 0u};
        RowSpec rs(j * L.n, d.Hd - j * L.n, L.n);
        if (disc) l_add_pack_t(lp, rs, W2, d.Hd, {{0, d.P, Pp}});
        else l_add_pack_t(lp, rs, W2, d.Hd, {{0, d.S, Sp}, {d.S, d.S, Sp}});
        L.w_tile_bytes = (uint32_t)(pk.data_bytes - t0);
      }
      L.ep_off = zero16; L.ep_floats = 0;
      lp.blayers.push_back(L);
    };
    // ---- posterior head, then the posterior hidden layer: [dL/dh' += | dL/de =] g_hidq W_post0
    head_dgrad(G_OQ, w ? w->post2_w : nul, G_HIDQ, F_HIDQ);
    {
      LLayerH L;
      L.type = LB_POST; L.dbuf = 1; L.out = G_GX; L.nseg = 1;
      int nth;
      tiles(Dp, 256, 16, nth, L.n);
      const int nte = (Ep + L.n - 1) / L.n;
      L.NT = nth + nte; L.layer = nth;
      L.kc = kc_for(L.n);
      L.w_off = pk.data_bytes;
      for (int j = 0; j < L.NT; ++j) {
        const size_t t0 = pk.data_bytes;
        L.seg[0] = {G_HIDQ, Hdp / 16, L.n, 0, 0, 1, 0u};
        RowSpec rs = (j < nth) ? RowSpec(j * L.n, d.D - j * L.n, L.n) : RowSpec(d.D + (j - nth) * L.n, d.E - (j - nth) * L.n, L.n);
        l_add_pack_t(lp, rs, w ? w->post0_w : nul, d.D + d.E, {{0, d.Hd, Hdp}});
        L.w_tile_bytes = (uint32_t)(pk.data_bytes - t0);
      }
      L.ep_off = zero16; L.ep_floats = 0;
      lp.blayers.push_back(L);
    }
    // ---- prior head, prior hidden + GRU gate backward
    head_dgrad(G_GO, w ? w->prior2_w : nul, G_GHID, F_HID);
    {
      LLayerH L;
      L.type = LB_GRU; L.dbuf = 1; L.out = G_GATES_I; L.nseg = 1;
      tiles(Dp, 128, 16, L.NT, L.n);
      L.kc = kc_for(L.n);
      L.w_off = pk.data_bytes;
      for (int j = 0; j < L.NT; ++j) {
        const size_t t0 = pk.data_bytes;
        L.seg[0] = {G_GHID, Hdp / 16, L.n, 0, 0, 1, 0u};
        l_add_pack_t(lp, RowSpec(j * L.n, d.D - j * L.n, L.n), w ? w->prior0_w : nul, d.D, {{0, d.Hd, Hdp}});
        L.w_tile_bytes = (uint32_t)(pk.data_bytes - t0);
      }
      L.ep_off = zero16; L.ep_floats = 0;
      lp.blayers.push_back(L);
    }
    // ---- dL/dx = [g_n | g_r | g_z] W_ih (x ELU'(x));  dL/dh_t = [g_r | g_z | g_n r] W_hh + dL/dh' z
    for (int hh = 0; hh < 2; ++hh) {
      LLayerH L;
      L.type = hh ? LB_CARRY : LB_ELU; L.dbuf = 1; L.out = G_GX; L.nseg = 1; L.layer = F_X;
      tiles(Dp, 256, 16, L.NT, L.n);
      L.kc = kc_for(L.n);
      L.w_off = pk.data_bytes;
      for (int j = 0; j < L.NT; ++j) {
        const size_t t0 = pk.data_bytes;
        L.seg[0] = {hh ? G_GATES_H : G_GATES_I, 3 * Dp / 16, L.n, 0, 0, 1, 0u};
        if (hh) l_add_pack_t(lp, RowSpec(j * L.n, d.D - j * L.n, L.n), w ? w->w_hh : nul, d.D, {{0, d.D, Dp}, {d.D, d.D, Dp}, {2 * d.D, d.D, Dp}});
        else l_add_pack_t(lp, RowSpec(j * L.n, d.D - j * L.n, L.n), w ? w->w_ih : nul, d.D, {{2 * d.D, d.D, Dp}, {0, d.D, Dp}, {d.D, d.D, Dp}});
        L.w_tile_bytes = (uint32_t)(pk.data_bytes - t0);
      }
      L.ep_off = zero16; L.ep_floats = 0;
      lp.blayers.push_back(L);
    }
    // ---- [dL/d post_s_t | dL/da_t] = g_x W_e
    {
      LLayerH L;
      L.type = LB_EMB; L.dbuf = 1; L.out = G_GMU; L.nseg = 1;
      int nts;
      tiles(Sip, 256, 16, nts, L.n);
      L.NT = nts + 1;
      L.kc = kc_for(L.n);
      L.w_off = pk.data_bytes;
      const int Ke = d.S_in + d.A;
      for (int j = 0; j < L.NT; ++j) {
        const size_t t0 = pk.data_bytes;
        L.seg[0] = {G_GX, Dp / 16, L.n, 0, 0, 1, 0u};
        RowSpec rs = (j < nts) ? RowSpec(j * L.n, d.S_in - j * L.n, L.n) : RowSpec(d.S_in, d.A, L.n);
        l_add_pack_t(lp, rs, w ? w->embed_w : nul, Ke, {{0, d.D, Dp}});
        L.w_tile_bytes = (uint32_t)(pk.data_bytes - t0);
      }
      L.ep_off = zero16; L.ep_floats = 0;
      lp.blayers.push_back(L);
    }
  }
  for (const std::vector<LLayerH>* v : {&lp.layers, &lp.blayers})
    for (const LLayerH& L : *v) {
      if (L.ep_floats * 4 > kLParamBytes || (L.ep_floats & 3)) return lp;
      for (int s = 0; s < L.nseg; ++s)
        if ((uint32_t)L.kc * (4096u + (uint32_t)L.seg[s].n * 32u) > kLSlotBytes) return lp;
    }
  auto al = [](size_t x) { return (x + 255) / 256 * 256; };
  pk.off_stage_tab = 256;
  pk.off_job_tab = 512;
  pk.off_packops = 768;
  pk.off_eparams = al(pk.off_packops + pk.pack_ops.size() * sizeof(PackOp) + pk.vec_ops.size() * sizeof(VecOp));
  pk.off_data = al(pk.off_eparams + pk.eparams_floats * sizeof(float));
  pk.total_bytes = al(pk.off_data + pk.data_bytes);
  pk.ok = true;
  lp.ok = true;
  return lp;
}

// mode 0: forward, nothing recorded; 1: forward, recording (world-model training); 2: backward
static LMem tcl_observe_layout(const LPlan& lp, const Dims& d, int mode) {
  LMem m;
  const int MT = (d.B + 127) / 128, T = std::max(1, d.T), H = std::max(1, d.T - 1);
  size_t off[2] = {kLTraceBytes, 0};
  auto place = [&](int buf, int steps, int where) {
    Region& r = m.r[buf];
    r.tile_bytes = (uint32_t)(buf_width(lp, buf) / 8) * kLTileChunk;
    r.step_bytes = (size_t)MT * r.tile_bytes;
    r.steps = steps; r.where = where; r.off = off[where];
    off[where] += ((size_t)steps * r.step_bytes + 1023) / 1024 * 1024;
  };
  const bool rec = mode != 0;
  const int sv = rec ? 1 : 0;
  place(F_H, rec ? T : 2, sv); place(F_S, rec ? T : 1, sv);
  place(F_X, rec ? H : 1, sv); place(F_HID, rec ? H : 1, sv); place(F_HIDQ, rec ? H : 1, sv);
  place(F_OBS, T, sv); place(F_ACT, T, sv);
  if (rec) for (int f : {F_GR, F_GZ, F_GN, F_GHN}) place(f, H, 1);
  if (mode != 2) place(F_SPRI, 1, 0);
  if (mode == 2) {
    place(G_OQ, H, 0); place(G_HIDQ, H, 0); place(G_GO, H, 0); place(G_GHID, H, 0); place(G_GATES_I, H, 0); place(G_GX, H, 0);
    m.r[G_GATES_H] = m.r[G_GATES_I];
    m.r[G_GATES_H].off += (size_t)(lp.Dp / 8) * kLTileChunk;
    place(G_GMU, 1, 0);
    m.carry_h = off[0]; off[0] += ((size_t)MT * 128 * lp.Dp * 4 + 1023) / 1024 * 1024;
    m.carry_s = off[0]; off[0] += ((size_t)MT * 128 * lp.Sip * 4 + 1023) / 1024 * 1024;
  }
  m.r[F_HN] = m.r[F_H]; m.r[F_SN] = m.r[F_S];
  if (mode != 2) {
    const size_t ib = ((size_t)d.B * T * d.C * sizeof(int32_t) + 1023) / 1024 * 1024;
    m.idx_pri = off[0]; off[0] += ib;
    m.idx_post = off[0]; off[0] += ib;
  }
  m.ws_bytes = off[0] + 1024;
  m.saved_bytes = off[1] + 1024;
  return m;
}

// fp32 [B][T][W] row-major -> bf16 panels [T][MT][Wp/8][128][8] (zeros past B / W)
__global__ void tcl_to_panel_kernel(const float* __restrict__ src, uint8_t* dst, int B, int T, int W, int Wp, int MT) {
  const int t = blockIdx.z, mt = blockIdx.y;
  const int chunks = Wp / 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < chunks * 128; i += gridDim.x * blockDim.x) {
    const int row = i & 127, k8 = i >> 7;
    const long long b = (long long)mt * 128 + row;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int col = k8 * 8 + e;
      v[e] = (b < B && col < W) ? src[(b * T + t) * W + col] : 0.f;
    }
    st_chunk_g(dst + (((size_t)t * MT + mt) * chunks + k8) * kLTileChunk + row * 16, v);
  }
}

bool tcl_observe_supported(const Dims& d) { return tcl_make_observe_plan(d, nullptr).ok; }
size_t tcl_observe_packed_bytes(const Dims& d) {
  LPlan lp = tcl_make_observe_plan(d, nullptr);
  return lp.ok ? lp.pk.total_bytes : 0;
}
size_t tcl_observe_workspace_bytes(const Dims& d) {
  LPlan lp = tcl_make_observe_plan(d, nullptr);
  if (!lp.ok) return 0;
  const int mode = (d.flags & WMRL_FLAG_BACKWARD) ? 2 : ((d.flags & WMRL_FLAG_SAVE_FOR_BWD) ? 1 : 0);
  return tcl_observe_layout(lp, d, mode).ws_bytes;
}
size_t tcl_observe_saved_bytes(const Dims& d) {
  LPlan lp = tcl_make_observe_plan(d, nullptr);
  return lp.ok ? tcl_observe_layout(lp, d, 1).saved_bytes : 0;
}
int tcl_pack_observe_weights(const Dims& d, const wmrl_rssm_params* w, void* packed, cudaStream_t st) {
  LPlan lp = tcl_make_observe_plan(d, w);
  WMRL_REQUIRE(lp.ok, "pack_observe_weights: shape not supported by the layer-wise tcgen05 path");
  return tc_pack_plan(lp.pk, (uint8_t*)packed, st);
}

int tcl_observe_fwd(const Dims& d, const void* packed, const float* obs, const float* act, const float* noise_pri,
                    const float* noise_post, float* pri_deter, float* pri_stoch, float* pri_da, float* pri_db,
                    float* post_deter, float* post_stoch, float* post_da, float* post_db, void* saved, void* ws,
                    cudaStream_t st) {
  LPlan lp = tcl_make_observe_plan(d, nullptr);
  WMRL_REQUIRE(lp.ok, "observe_fwd: shape not supported by the layer-wise tcgen05 path");
  const bool rec = saved != nullptr;
  const LMem m = tcl_observe_layout(lp, d, rec ? 1 : 0);
  LCall c{&lp, &m, &d, (uint8_t*)ws, (uint8_t*)saved, nullptr, nullptr, (d.B + 127) / 128, 148, st};
  if (int r = tcl_device_setup(c.nsm)) return r;
  const uint8_t* img = (const uint8_t*)packed;
  c.data = img + lp.pk.off_data;
  c.ep = (const float*)(img + lp.pk.off_eparams);
  const int T = d.T, MT = c.MT, H = T - 1;
  // index 0 of both sequences = the zero initial state (rssm.py:166-172 / 215-220); h / s panels start at zero
  tcl_init_kernel<<<dim3(8, MT), 256, 0, st>>>(nullptr, nullptr, pri_deter, pri_stoch, c.at(F_H, 0), c.at(F_S, 0), d.B, H, d.D,
                                               d.S_in, lp.Dp, lp.Sip);
  tcl_init_kernel<<<dim3(8, MT), 256, 0, st>>>(nullptr, nullptr, post_deter, post_stoch, c.at(F_H, 0), c.at(F_S, 0), d.B, H,
                                               d.D, d.S_in, lp.Dp, lp.Sip);
  WMRL_CUDA_OK(cudaGetLastError());
  if (H < 1) return 0;
  tcl_to_panel_kernel<<<dim3(8, MT, T), 256, 0, st>>>(obs, c.at(F_OBS, 0), d.B, T, d.E, lp.Ep, MT);
  tcl_to_panel_kernel<<<dim3(1, MT, T), 256, 0, st>>>(act, c.at(F_ACT, 0), d.B, T, d.A, 16, MT);
  WMRL_CUDA_OK(cudaGetLastError());
  Dims dstep = d;
  dstep.T = H;   // the sequence tensors hold H + 1 = T entries
  LCall cs = c;
  cs.d = &dstep;
  for (int t = 0; t < H; ++t) {
    auto step_of = [&](int buf) {
      if (buf == F_H) return rec ? t : (t & 1);
      if (buf == F_HN) return rec ? t + 1 : ((t + 1) & 1);
      if (buf == F_S) return rec ? t : 0;
      if (buf == F_SN) return rec ? t + 1 : 0;
      if (buf == F_OBS) return t + 1;
      if (buf == F_ACT) return t;
      if (buf == F_SPRI) return 0;
      return rec ? t : 0;
    };
    for (const LLayerH& L : lp.layers) {
      LayerArgs a{};
      a.t = t;
      a.save = rec ? 1 : 0;
      if (rec && L.type == LE_GRU) {
        a.aux[0] = c.at(F_GR, t); a.aux[1] = c.at(F_GZ, t); a.aux[2] = c.at(F_GN, t); a.aux[3] = c.at(F_GHN, t);
        a.aux_tile_bytes = m.r[F_GR].tile_bytes;
      }
      const bool post = L.layer == 1 && (L.type == LE_SAMPLE_DISC || L.type == LE_SAMPLE_CONT);
      a.deter_seq = pri_deter;
      a.stoch_seq = post ? post_stoch : pri_stoch; a.dist_a = post ? post_da : pri_da; a.dist_b = post ? post_db : pri_db;
      a.actions = nullptr; a.noise = post ? noise_post : noise_pri;
      a.idx = (int32_t*)((uint8_t*)ws + (post ? m.idx_post : m.idx_pri));
      a.trace = nullptr;
      int ss[3];
      for (int s = 0; s < L.nseg; ++s) ss[s] = step_of(L.seg[s].buf);
      if (int r = tcl_launch(cs, L, a, ss, step_of(L.out))) return r;
    }
  }
  if (d.variant == WMRL_DISCRETE) {
    tcl_onehot_kernel<<<4 * c.nsm, 256, 0, st>>>((const int32_t*)((uint8_t*)ws + m.idx_pri), pri_stoch, d.B, H, d.C, d.S);
    tcl_onehot_kernel<<<4 * c.nsm, 256, 0, st>>>((const int32_t*)((uint8_t*)ws + m.idx_post), post_stoch, d.B, H, d.C, d.S);
    WMRL_CUDA_OK(cudaGetLastError());
  }
  // both sequences carry the same deterministic states (rssm.py:84-91)
  WMRL_CUDA_OK(cudaMemcpyAsync(post_deter, pri_deter, (size_t)d.B * T * d.D * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

// Full backward of filtering on the chain (world-model training, world_model.py:120-163): per step, backwards in
// time, the two sample backwards (SIMT) and 7 dgrad GEMMs with fused epilogues; every gradient panel is kept per step
// and the RSSM weight gradients are MN-major big-K tcgen05 GEMMs over (recorded activation, gradient) panel pairs after
// the scan, the bias gradients column sums.  g_deter = dL/d pri_deter + dL/d post_deter (both sequences hold the same h).
int tcl_observe_bwd(const Dims& d, const void* packed, const float* noise_pri, const float* noise_post, const float* pri_deter,
                    const float* pri_da, const float* pri_db, const float* post_da, const float* post_db, const void* saved,
                    const float* g_deter, const float* g_pri_stoch, const float* g_pri_da, const float* g_pri_db,
                    const float* g_post_stoch, const float* g_post_da, const float* g_post_db, const wmrl_rssm_grads* gw,
                    float* g_obs, float* g_act, void* ws, cudaStream_t st) {
  LPlan lp = tcl_make_observe_plan(d, nullptr);
  WMRL_REQUIRE(lp.ok, "observe_bwd: shape not supported by the layer-wise tcgen05 path");
  const LMem m = tcl_observe_layout(lp, d, 2);
  LCall c{&lp, &m, &d, (uint8_t*)ws, (uint8_t*)const_cast<void*>(saved), nullptr, nullptr, (d.B + 127) / 128, 148, st};
  if (int r = tcl_device_setup(c.nsm)) return r;
  const uint8_t* img = (const uint8_t*)packed;
  c.data = img + lp.pk.off_data;
  c.ep = (const float*)(img + lp.pk.off_eparams);
  const int T = d.T, MT = c.MT, H = T - 1;
  if (H < 1) return 0;
  const bool disc = d.variant == WMRL_DISCRETE;
  Dims dstep = d;
  dstep.T = H;
  LCall cs = c;
  cs.d = &dstep;
  float* carry_h = (float*)((uint8_t*)ws + m.carry_h);
  float* carry_s = (float*)((uint8_t*)ws + m.carry_s);
  {
    const long long n = (long long)MT * 128 * lp.Dp;
    tcl_carry_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(carry_h, g_deter, d.B, H, d.D, lp.Dp, n);
    WMRL_CUDA_OK(cudaGetLastError());
  }
  for (int t = H - 1; t >= 0; --t) {
    for (int post = 1; post >= 0; --post) {   // dL/d(head outputs) of the state produced by step t
      SbArgs s{};
      s.carry_s = (post && t < H - 1) ? carry_s : nullptr;
      s.g_stoch = post ? g_post_stoch : g_pri_stoch; s.g_da = post ? g_post_da : g_pri_da; s.g_db = post ? g_post_db : g_pri_db;
      s.dist_a = post ? post_da : pri_da; s.dist_b = post ? post_db : pri_db; s.noise = post ? noise_post : noise_pri;
      s.out = c.at(post ? G_OQ : G_GO, t);
      s.B = d.B; s.H = H; s.t = t; s.S = d.S; s.C = d.C; s.Sp = lp.Sp; s.Sip = lp.Sip; s.Pp = lp.Pp; s.MT = MT;
      if (disc) {
        const long long nw = (long long)MT * 128 * ((d.C + kSbGroups - 1) / kSbGroups);
        tcl_sample_bwd_disc_kernel<<<(unsigned)((nw * 32 + 255) / 256), 256, 0, st>>>(s);
      } else {
        const long long n = (long long)MT * 128 * (lp.Sp / 8);
        tcl_sample_bwd_cont_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s);
      }
      WMRL_CUDA_OK(cudaGetLastError());
    }
    for (const LLayerH& L : lp.blayers) {
      LayerArgs a{};
      a.t = t;
      a.deter_seq = const_cast<float*>(pri_deter);
      switch (L.type) {
        case LB_ELU:
          a.aux[0] = c.at(L.layer, t); a.aux_tile_bytes = m.r[L.layer].tile_bytes;
          break;
        case LB_POST:
          a.f0 = carry_h; a.f3 = g_obs; a.nt_split = L.layer;
          break;
        case LB_GRU:
          a.aux[0] = c.at(F_GR, t); a.aux[1] = c.at(F_GZ, t); a.aux[2] = c.at(F_GN, t); a.aux[3] = c.at(F_GHN, t);
          a.aux[4] = c.at(F_H, t);
          a.aux_tile_bytes = m.r[F_GR].tile_bytes;
          a.f0 = carry_h;
          break;
        case LB_CARRY:
          a.f0 = carry_h; a.f1 = const_cast<float*>(g_deter); a.f2 = nullptr;
          break;
        case LB_EMB:
          a.f0 = carry_s; a.f1 = nullptr; a.f2 = nullptr; a.f3 = g_act; a.mode = 1;
          break;
      }
      int ss[3] = {t, t, t};
      if (int r = tcl_launch(cs, L, a, ss, L.out == G_GMU ? 0 : t)) return r;
    }
  }
  // ---- weight gradients over (recorded activation, gradient) panel pairs, blocks = (step, row tile)
  const int nblk = H * MT;
  auto wgrad = [&](uint8_t* a_base, size_t a_blk, int m_cols, int m_valid, uint8_t* b_base, size_t b_blk, int n_cols, int n_valid,
                   float* dW, int ldw) -> int {
    if (!dW) return 0;
    Wg2Params q{};
    q.a_base = a_base; q.a_blk = a_blk; q.b_base = b_base; q.b_blk = b_blk;
    q.nblocks = nblk; q.m_cols = m_cols; q.m_valid = m_valid; q.n_cols = n_cols; q.n_valid = n_valid;
    q.dW = dW; q.ldw = ldw;
    q.mtiles = (m_cols + 127) / 128; q.ntiles = (n_cols + 255) / 256;
    const int tl = q.mtiles * q.ntiles;
    q.splits = std::max(1, std::min(nblk, c.nsm / tl));
    tcl_wgrad_kernel<<<dim3(tl, q.splits), kWg2Threads, kWg2Smem, st>>>(q);
    WMRL_CUDA_OK(cudaGetLastError());
    return 0;
  };
  auto colsum = [&](uint8_t* x_base, size_t x_blk, int cols, int valid, float* out) -> int {
    if (!out) return 0;
    Cs2Params q{};
    q.x_base = x_base; q.x_blk = x_blk; q.y_base = nullptr; q.y_blk = 0; q.nblocks = nblk; q.valid = valid; q.out = out;
    const int chunks = cols / 8;
    const int ysplit = std::max(1, std::min((nblk + kCs2Warps - 1) / kCs2Warps, (8 * c.nsm + chunks - 1) / chunks));
    tcl_colsum_kernel<<<dim3(chunks, ysplit), kCs2Warps * 32, 0, st>>>(q);
    WMRL_CUDA_OK(cudaGetLastError());
    return 0;
  };
  auto tb = [&](int buf) { return (size_t)m.r[buf].tile_bytes; };
  const int D = d.D, Dp = lp.Dp, Hdp = lp.Hdp, Sp = lp.Sp;
  const size_t fchunk = (size_t)(Dp / 8) * kLTileChunk;   // one Dp-wide field of the gate panel
  // heads: W2 [P, Hd] (continuous: rows [mean | raw] sit in two Sp-wide halves of the gradient panel), W0
  for (int post = 0; post < 2; ++post) {
    const int gout = post ? G_OQ : G_GO, ghid = post ? G_HIDQ : G_GHID, hid = post ? F_HIDQ : F_HID;
    float* dW2 = post ? gw->post2_w : gw->prior2_w;
    float* db2 = post ? gw->post2_b : gw->prior2_b;
    if (disc) {
      if (int r = wgrad(c.at(hid, 0), tb(hid), Hdp, d.Hd, c.at(gout, 0), tb(gout), lp.Pp, d.P, dW2, d.Hd)) return r;
      if (int r = colsum(c.at(gout, 0), tb(gout), lp.Pp, d.P, db2)) return r;
    } else {
      for (int half = 0; half < 2; ++half) {
        uint8_t* gb = c.at(gout, 0) + (size_t)half * (Sp / 8) * kLTileChunk;
        if (int r = wgrad(c.at(hid, 0), tb(hid), Hdp, d.Hd, gb, tb(gout), Sp, d.S, dW2 ? dW2 + (size_t)half * d.S * d.Hd : nullptr, d.Hd)) return r;
        if (int r = colsum(gb, tb(gout), Sp, d.S, db2 ? db2 + half * d.S : nullptr)) return r;
      }
    }
    float* dW0 = post ? gw->post0_w : gw->prior0_w;
    const int ld0 = post ? D + d.E : D;
    if (int r = wgrad(c.at(F_H, 1), tb(F_H), Dp, D, c.at(ghid, 0), tb(ghid), Hdp, d.Hd, dW0, ld0)) return r;
    if (post)
      if (int r = wgrad(c.at(F_OBS, 1), tb(F_OBS), lp.Ep, d.E, c.at(ghid, 0), tb(ghid), Hdp, d.Hd, dW0 ? dW0 + D : nullptr, ld0)) return r;
    if (int r = colsum(c.at(ghid, 0), tb(ghid), Hdp, d.Hd, post ? gw->post0_b : gw->prior0_b)) return r;
  }
  // GRU: gate panel fields [g_n | g_r | g_z | g_n r]; torch rows [r | z | n]
  {
    uint8_t* gates = c.at(G_GATES_I, 0);
    const size_t gblk = tb(G_GATES_I);
    const int f_ih[3] = {1, 2, 0}, f_hh[3] = {1, 2, 3};   // panel field of torch gate block r, z, n
    for (int g = 0; g < 3; ++g) {
      if (int r = wgrad(c.at(F_X, 0), tb(F_X), Dp, D, gates + f_ih[g] * fchunk, gblk, Dp, D, gw->w_ih ? gw->w_ih + (size_t)g * D * D : nullptr, D)) return r;
      if (int r = wgrad(c.at(F_H, 0), tb(F_H), Dp, D, gates + f_hh[g] * fchunk, gblk, Dp, D, gw->w_hh ? gw->w_hh + (size_t)g * D * D : nullptr, D)) return r;
      if (int r = colsum(gates + f_ih[g] * fchunk, gblk, Dp, D, gw->b_ih ? gw->b_ih + g * D : nullptr)) return r;
      if (int r = colsum(gates + f_hh[g] * fchunk, gblk, Dp, D, gw->b_hh ? gw->b_hh + g * D : nullptr)) return r;
    }
  }
  // embed: W_e [D, S_in + A] over [post s_t | a_t]
  {
    const int Ke = d.S_in + d.A;
    if (int r = wgrad(c.at(F_S, 0), tb(F_S), lp.Sip, d.S_in, c.at(G_GX, 0), tb(G_GX), Dp, D, gw->embed_w, Ke)) return r;
    if (int r = wgrad(c.at(F_ACT, 0), tb(F_ACT), 16, d.A, c.at(G_GX, 0), tb(G_GX), Dp, D, gw->embed_w ? gw->embed_w + d.S_in : nullptr, Ke)) return r;
    if (int r = colsum(c.at(G_GX, 0), tb(G_GX), Dp, D, gw->embed_b)) return r;
  }
  return 0;
}

// ==========================================================================
// MLP heads on the chain: the reward / done models (rssm_world_model/models.py:3-37: [Linear, LayerNorm, SiLU] x depth,
// Linear) and the value critic (trainers/value/critic.py:4-29: ReLU) over the features the rollout returns
// (world_model.py:178-180, value_trainer.py:121-123).  R = B * H rows are row tiles of the same tc_layer_kernel:
//   forward   x panel -> depth x (GEMM + fused LayerNorm + activation, recording y / x-hat / rstd when training) ->
//             output Linear as a 16-column GEMM tile -> fp32 out[R][O]
//   backward  g_out -> panel; per block, from the last: dgrad GEMM over the TRANSPOSED next weight with the fused
//             LayerNorm + activation backward; the input gradient g_x[R][F] in fp32 (coalesced); weight gradients as
//             MN-major big-K GEMMs over (recorded activation, gradient) panels, biases / affine as column sums.
// One feature panel (wmrl_mlp_to_panel) feeds every head that consumes the same features.
// ==========================================================================
struct MlpShape {
  long long R; int F, Hh, L, O, act, flags;
  int Fp, MT;
};
static bool mlp_shape(const wmrl_mlp_dims* d, MlpShape& m) {
  if (!d) return false;
  m.R = d->rows; m.F = d->in_dim; m.Hh = d->hidden; m.L = d->depth; m.O = d->out_dim; m.act = d->act; m.flags = d->flags;
  if (m.R < 0 || m.R > 65535ll * 128 || m.F < 1 || m.F > 4096 || m.Hh < 32 || m.Hh > 512 || m.Hh % 32 != 0) return false;
  if (m.L < 1 || m.L > WMRL_MAX_MLP_LAYERS || m.O < 1 || m.O > 16 || m.act < 0 || m.act > 2) return false;
  m.Fp = ceil16(m.F);
  m.MT = (int)((m.R + 127) / 128);
  return true;
}
enum { M_X = 0, M_GOUT, M_Y0, M_XH0 = M_Y0 + WMRL_MAX_MLP_LAYERS, M_RSTD0 = M_XH0 + WMRL_MAX_MLP_LAYERS,
       M_GP0 = M_RSTD0 + WMRL_MAX_MLP_LAYERS, M_GU0 = M_GP0 + WMRL_MAX_MLP_LAYERS, M_COUNT = M_GU0 + WMRL_MAX_MLP_LAYERS };

static LPlan tcm_make_plan(const MlpShape& m, const wmrl_mlp_params* w) {
  LPlan lp;
  Plan& pk = lp.pk;
  const float* nul = nullptr;
  constexpr float kLog2e = 1.4426950408889634f;
  auto kc_for = [&](int nmax) { int kc = (int)(kLSlotBytes / (4096u + (uint32_t)nmax * 32u)); return std::max(1, std::min(kc, 8)); };
  const int fwd_type = m.act == kActElu ? LE_LN_ELU : (m.act == kActSilu ? LE_LN_SILU : LE_LN_RELU);
  const int bwd_type = m.act == kActElu ? LB_LN : (m.act == kActSilu ? LB_LN_SILU : LB_LN_RELU);
  // ---- forward blocks
  for (int l = 0; l < m.L; ++l) {
    LLayerH L;
    L.type = fwd_type; L.NT = 1; L.n = m.Hh; L.dbuf = m.Hh <= 256 ? 1 : 0; L.out = M_Y0 + l; L.layer = l; L.nseg = 1;
    L.kc = kc_for(m.Hh);
    L.w_off = pk.data_bytes;
    const int K = l == 0 ? m.F : m.Hh, Kp = l == 0 ? m.Fp : m.Hh;
    L.seg[0] = {l == 0 ? M_X : M_Y0 + l - 1, Kp / 16, m.Hh, 0, 0, 1, 0u};
    l_add_pack(lp, RowSpec(0, m.Hh, m.Hh), w ? w->lin_w[l] : nul, K, 0, K, Kp);
    L.w_tile_bytes = (uint32_t)(pk.data_bytes - L.w_off);
    L.ep_off = add_vec(pk, w ? w->lin_b[l] : nul, nul, 0, m.Hh, m.Hh);
    add_vec(pk, w ? w->ln_g[l] : nul, nul, 0, m.Hh, m.Hh, kLog2e);
    add_vec(pk, w ? w->ln_b[l] : nul, nul, 0, m.Hh, m.Hh, kLog2e);
    L.ep_floats = 3u * m.Hh;
    lp.layers.push_back(L);
  }
  // ---- output Linear
  {
    LLayerH L;
    L.type = LE_OUT; L.NT = 1; L.n = 16; L.dbuf = 1; L.out = -1; L.kc = kc_for(16); L.nseg = 1;
    L.w_off = pk.data_bytes;
    L.seg[0] = {M_Y0 + m.L - 1, m.Hh / 16, 16, 0, 0, 1, 0u};
    l_add_pack(lp, RowSpec(0, m.O, 16), w ? w->out_w : nul, m.Hh, 0, m.Hh, m.Hh);
    L.w_tile_bytes = (uint32_t)(pk.data_bytes - L.w_off);
    L.ep_off = add_vec(pk, w ? w->out_b : nul, nul, 0, m.O, 16);
    L.ep_floats = 16;
    lp.layers.push_back(L);
  }
  // ---- backward blocks, last first: g_y_l = g_next W_next -> LayerNorm + activation backward -> g_pre_l
  for (int l = m.L - 1; l >= 0; --l) {
    LLayerH L;
    L.type = bwd_type; L.NT = 1; L.n = m.Hh; L.dbuf = m.Hh <= 256 ? 1 : 0; L.out = M_GP0 + l; L.layer = l; L.nseg = 1;
    L.kc = kc_for(m.Hh);
    L.w_off = pk.data_bytes;
    if (l == m.L - 1) {
      L.seg[0] = {M_GOUT, 1, m.Hh, 0, 0, 1, 0u};
      l_add_pack_t(lp, RowSpec(0, m.Hh, m.Hh), w ? w->out_w : nul, m.Hh, {{0, m.O, 16}});
    } else {
      L.seg[0] = {M_GP0 + l + 1, m.Hh / 16, m.Hh, 0, 0, 1, 0u};
      l_add_pack_t(lp, RowSpec(0, m.Hh, m.Hh), w ? w->lin_w[l + 1] : nul, m.Hh, {{0, m.Hh, m.Hh}});
    }
    L.w_tile_bytes = (uint32_t)(pk.data_bytes - L.w_off);
    L.ep_off = add_vec(pk, w ? w->ln_g[l] : nul, nul, 0, m.Hh, m.Hh, kLog2e);
    add_vec(pk, w ? w->ln_b[l] : nul, nul, 0, m.Hh, m.Hh, kLog2e);
    add_vec(pk, w ? w->ln_g[l] : nul, nul, 0, m.Hh, m.Hh);
    L.ep_floats = 3u * m.Hh;
    lp.blayers.push_back(L);
  }
  // ---- input gradient: g_x = g_pre_0 W_0, column tiles of <= 256 input features
  {
    LLayerH L;
    L.type = LB_XGRAD; L.dbuf = 1; L.out = -1; L.nseg = 1;
    const int total = ceil_to(m.F, 32);
    L.NT = (total + 255) / 256;
    L.n = ceil_to((total + L.NT - 1) / L.NT, 32);
    L.kc = kc_for(L.n);
    L.w_off = pk.data_bytes;
    for (int j = 0; j < L.NT; ++j) {
      const size_t t0 = pk.data_bytes;
      L.seg[0] = {M_GP0, m.Hh / 16, L.n, 0, 0, 1, 0u};
      l_add_pack_t(lp, RowSpec(j * L.n, m.F - j * L.n, L.n), w ? w->lin_w[0] : nul, m.F, {{0, m.Hh, m.Hh}});
      L.w_tile_bytes = (uint32_t)(pk.data_bytes - t0);
    }
    L.ep_off = add_vec(pk, nul, nul, 0, 0, 16); L.ep_floats = 0;
    lp.blayers.push_back(L);
  }
  // ---- wide blocks (hidden > 256): the same layers as CTA pairs of hidden / 2 columns each (see LPlan::layers_p)
  if (m.Hh > 256 && m.Hh % 64 == 0) {
    const int ns = m.Hh / 2;
    lp.cpair = 2;
    lp.layers_p = lp.layers;
    lp.blayers_p = lp.blayers;
    for (LLayerH& L : lp.layers_p) {
      if (L.type != fwd_type) continue;
      const int l = L.layer;
      const int K = l == 0 ? m.F : m.Hh, Kp = l == 0 ? m.Fp : m.Hh;
      L.NT = 2; L.n = ns; L.dbuf = 1; L.kc = kc_for(ns); L.csplit = 2;
      L.w_off = pk.data_bytes;
      for (int j = 0; j < 2; ++j) {
        const size_t t0 = pk.data_bytes;
        L.seg[0] = {l == 0 ? M_X : M_Y0 + l - 1, Kp / 16, ns, 0, 0, 1, 0u};
        l_add_pack(lp, RowSpec(j * ns, ns, ns), w ? w->lin_w[l] : nul, K, 0, K, Kp);
        L.w_tile_bytes = (uint32_t)(pk.data_bytes - t0);
        const uint32_t at = add_vec(pk, w ? w->lin_b[l] : nul, nul, j * ns, ns, ns);
        add_vec(pk, w ? w->ln_g[l] : nul, nul, j * ns, ns, ns, kLog2e);
        add_vec(pk, w ? w->ln_b[l] : nul, nul, j * ns, ns, ns, kLog2e);
        if (j == 0) L.ep_off = at;
      }
      L.ep_floats = 3u * ns;
    }
    for (LLayerH& L : lp.blayers_p) {
      if (L.type != bwd_type) continue;
      const int l = L.layer;
      L.NT = 2; L.n = ns; L.dbuf = 1; L.kc = kc_for(ns); L.csplit = 2;
      L.w_off = pk.data_bytes;
      for (int j = 0; j < 2; ++j) {
        const size_t t0 = pk.data_bytes;
        RowSpec rs(j * ns, ns, ns);
        if (l == m.L - 1) {
          L.seg[0] = {M_GOUT, 1, ns, 0, 0, 1, 0u};
          l_add_pack_t(lp, rs, w ? w->out_w : nul, m.Hh, {{0, m.O, 16}});
        } else {
          L.seg[0] = {M_GP0 + l + 1, m.Hh / 16, ns, 0, 0, 1, 0u};
          l_add_pack_t(lp, rs, w ? w->lin_w[l + 1] : nul, m.Hh, {{0, m.Hh, m.Hh}});
        }
        L.w_tile_bytes = (uint32_t)(pk.data_bytes - t0);
        const uint32_t at = add_vec(pk, w ? w->ln_g[l] : nul, nul, j * ns, ns, ns, kLog2e);
        add_vec(pk, w ? w->ln_b[l] : nul, nul, j * ns, ns, ns, kLog2e);
        add_vec(pk, w ? w->ln_g[l] : nul, nul, j * ns, ns, ns);
        if (j == 0) L.ep_off = at;
      }
      L.ep_floats = 3u * ns;
    }
  }
  for (const std::vector<LLayerH>* v : {&lp.layers, &lp.blayers, &lp.layers_p, &lp.blayers_p})
    for (const LLayerH& L : *v) {
      if (L.ep_floats * 4 > kLParamBytes || (L.ep_floats & 3)) return lp;
      for (int s = 0; s < L.nseg; ++s)
        if ((uint32_t)L.kc * (4096u + (uint32_t)L.seg[s].n * 32u) > kLSlotBytes) return lp;
    }
  auto al = [](size_t x) { return (x + 255) / 256 * 256; };
  pk.off_stage_tab = 256;
  pk.off_job_tab = 512;
  pk.off_packops = 768;
  pk.off_eparams = al(pk.off_packops + pk.pack_ops.size() * sizeof(PackOp) + pk.vec_ops.size() * sizeof(VecOp));
  pk.off_data = al(pk.off_eparams + pk.eparams_floats * sizeof(float));
  pk.total_bytes = al(pk.off_data + pk.data_bytes);
  pk.ok = true;
  lp.ok = true;
  return lp;
}

// panel buffers of one MLP call: [MT][cols/8][2 KB] each.  `saved` holds what the backward reads (y, x-hat, rstd per
// block); the forward workspace the y panels when nothing is recorded; the backward workspace g_out, g_pre, g_u.
struct MlpMem {
  size_t off[M_COUNT]; uint32_t tile[M_COUNT]; int where[M_COUNT];   // where: 0 ws, 1 saved, 2 caller's x panel
  size_t ws_bytes = 0, saved_bytes = 0;
};
static MlpMem tcm_layout(const MlpShape& m, int mode) {   // mode 0 forward, 1 forward recording, 2 backward
  MlpMem mm{};
  size_t o[2] = {0, 0};
  auto place = [&](int buf, int cols, int where, size_t row_bytes = 0) {
    mm.tile[buf] = row_bytes ? (uint32_t)(128 * row_bytes) : (uint32_t)(cols / 8) * kLTileChunk;
    mm.where[buf] = where;
    mm.off[buf] = o[where];
    o[where] += ((size_t)m.MT * mm.tile[buf] + 1023) / 1024 * 1024;
  };
  mm.tile[M_X] = (uint32_t)(m.Fp / 8) * kLTileChunk; mm.where[M_X] = 2; mm.off[M_X] = 0;
  const int sv = mode == 0 ? 0 : 1;
  for (int l = 0; l < m.L; ++l) {
    if (mode == 0 && l >= 2) { mm.tile[M_Y0 + l] = mm.tile[M_Y0 + (l & 1)]; mm.where[M_Y0 + l] = 0; mm.off[M_Y0 + l] = mm.off[M_Y0 + (l & 1)]; }
    else place(M_Y0 + l, m.Hh, sv);
    if (mode != 0) { place(M_XH0 + l, m.Hh, 1); place(M_RSTD0 + l, 0, 1, sizeof(float)); }
  }
  if (mode == 2) {
    place(M_GOUT, 16, 0);
    const bool wg = (m.flags & WMRL_MLP_WGRAD) != 0;
    for (int l = 0; l < m.L; ++l) {
      // without weight gradients a block's g_pre is read once, by the next dgrad: two panels ping-pong
      if (!wg && l >= 2) { mm.tile[M_GP0 + l] = mm.tile[M_GP0 + (l & 1)]; mm.where[M_GP0 + l] = 0; mm.off[M_GP0 + l] = mm.off[M_GP0 + (l & 1)]; }
      else place(M_GP0 + l, m.Hh, 0);
      if (wg) place(M_GU0 + l, m.Hh, 0);
    }
  }
  mm.ws_bytes = o[0] + 1024;
  mm.saved_bytes = o[1] + 1024;
  return mm;
}

bool tcm_supported(const wmrl_mlp_dims* d) { MlpShape m; return mlp_shape(d, m) && tcm_make_plan(m, nullptr).ok; }
size_t tcm_packed_bytes(const wmrl_mlp_dims* d) {
  MlpShape m;
  if (!mlp_shape(d, m)) return 0;
  LPlan lp = tcm_make_plan(m, nullptr);
  return lp.ok ? lp.pk.total_bytes : 0;
}
size_t tcm_panel_bytes(const wmrl_mlp_dims* d) {
  MlpShape m;
  return mlp_shape(d, m) ? (size_t)m.MT * (m.Fp / 8) * kLTileChunk + 1024 : 0;
}
size_t tcm_saved_bytes(const wmrl_mlp_dims* d) { MlpShape m; return mlp_shape(d, m) ? tcm_layout(m, 1).saved_bytes : 0; }
size_t tcm_workspace_bytes(const wmrl_mlp_dims* d, int backward) {
  MlpShape m;
  return mlp_shape(d, m) ? tcm_layout(m, backward ? 2 : ((m.flags & WMRL_MLP_SAVE) ? 1 : 0)).ws_bytes : 0;
}
int tcm_pack(const wmrl_mlp_dims* d, const wmrl_mlp_params* w, void* packed, cudaStream_t st) {
  MlpShape m;
  WMRL_REQUIRE(mlp_shape(d, m), "mlp_pack: shape not supported (hidden % 32, hidden <= 512, out_dim <= 16)");
  LPlan lp = tcm_make_plan(m, w);
  WMRL_REQUIRE(lp.ok, "mlp_pack: shape not supported by the layer-wise tcgen05 path");
  return tc_pack_plan(lp.pk, (uint8_t*)packed, st);
}
int tcm_to_panel(const wmrl_mlp_dims* d, const float* x, void* panel, cudaStream_t st) {
  MlpShape m;
  WMRL_REQUIRE(mlp_shape(d, m), "mlp_to_panel: bad dims");
  if (m.MT == 0) return 0;
  tcl_to_panel_kernel<<<dim3(4, (unsigned)m.MT, 1), 256, 0, st>>>(x, (uint8_t*)panel, (int)m.R, 1, m.F, m.Fp, m.MT);
  WMRL_CUDA_OK(cudaGetLastError());
  return 0;
}

struct MlpCall {
  const MlpShape* m; const MlpMem* mm;
  uint8_t* ws; uint8_t* saved; const uint8_t* xpanel;
  const uint8_t* data; const float* ep;
  int nsm; cudaStream_t st;
  int tile0 = 0, ntiles = 0;   // row-tile range of the current chunk (ntiles == 0: all)
  uint8_t* at(int buf) const {
    const int w = mm->where[buf];
    return (w == 2 ? const_cast<uint8_t*>(xpanel) : (w == 1 ? saved : ws)) + mm->off[buf] + (size_t)tile0 * mm->tile[buf];
  }
};
// Optional row chunks of the heads (WMRL_MLP_CHUNK = row tiles per chunk; default: one chunk): all layers of a chunk run before
// the next chunk starts, so a layer's input panel is still in L2 when the next layer reads it.  Measured at 983 040 rows
// (forward / forward+backward of the four heads): no chunks 5.8 / 20.1 ms, 2 048 tiles 6.2 / 21.0, 1 024: 6.5 / 21.6,
// 512: 7.6 / 24.3, 256 (with the CTA-pair form): 9.0 / 26.9 - the ramp and tail of the many short launches cost more than the
// HBM round trips they save, so it stays off.
static int mlp_chunk_tiles() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("WMRL_MLP_CHUNK"); v = e ? atoi(e) : 0; if (v <= 0) v = 1 << 30; }
  return v;
}
#define kMlpChunkTiles mlp_chunk_tiles()
static int tcm_launch(const MlpCall& c, const LLayerH& L, LayerArgs& a) {
  const MlpShape& m = *c.m;
  a.nseg = 1;
  const int buf = L.seg[0].buf;
  a.seg[0].a = c.at(buf);
  a.seg[0].a_tile_bytes = c.mm->tile[buf];
  a.seg[0].w_off = L.seg[0].w_off;
  a.seg[0].k16 = (uint16_t)L.seg[0].k16; a.seg[0].n = (uint16_t)L.seg[0].n;
  a.seg[0].tmem_col = 0; a.seg[0].split = 0; a.seg[0].fresh = 1;
  a.w = c.data + L.w_off; a.w_tile_bytes = L.w_tile_bytes;
  a.ep = c.ep + L.ep_off; a.ep_floats = L.ep_floats;
  const int nt = c.ntiles > 0 ? c.ntiles : m.MT;
  a.kc = L.kc; a.dbuf = L.dbuf; a.type = L.type; a.n = L.n; a.NT = L.NT; a.MT = nt;
  a.csplit = L.csplit; a.ln_width = L.csplit > 1 ? L.n * L.NT : L.n;
  a.B = (int)std::min<long long>(m.R - (long long)c.tile0 * 128, (long long)nt * 128); a.H = 1; a.t = 0; a.D = m.F; a.A = m.O;
  if (L.out >= 0) { a.out = c.at(L.out); a.out_tile_bytes = c.mm->tile[L.out]; a.out_cols = m.Hh; }
  const int items = nt * L.NT;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)std::min(items, c.nsm));
  cfg.blockDim = dim3(kLThreads);
  cfg.dynamicSmemBytes = kLSmemBytes;
  cfg.stream = c.st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (L.csplit > 1) {
    cfg.gridDim = dim3((unsigned)(std::min(nt, tcl_max_clusters(L.csplit)) * L.csplit));
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = (unsigned)L.csplit; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
    cfg.numAttrs = 2;
  }
  WMRL_CUDA_OK(cudaLaunchKernelEx(&cfg, layer_kernel_for(L.type, 4), a));
  return 0;
}

int tcm_fwd(const wmrl_mlp_dims* d, const void* packed, const void* x_panel, float* out, void* saved, void* ws,
            cudaStream_t st) {
  MlpShape m;
  WMRL_REQUIRE(mlp_shape(d, m), "mlp_fwd: shape not supported");
  if (m.MT == 0) return 0;
  LPlan lp = tcm_make_plan(m, nullptr);
  WMRL_REQUIRE(lp.ok, "mlp_fwd: shape not supported by the layer-wise tcgen05 path");
  const bool rec = saved != nullptr;
  const MlpMem mm = tcm_layout(m, rec ? 1 : 0);
  MlpCall c{&m, &mm, (uint8_t*)ws, (uint8_t*)saved, (const uint8_t*)x_panel, nullptr, nullptr, 148, st};
  if (int r = tcl_device_setup(c.nsm)) return r;
  const uint8_t* img = (const uint8_t*)packed;
  c.data = img + lp.pk.off_data;
  c.ep = (const float*)(img + lp.pk.off_eparams);
  for (int t0 = 0; t0 < m.MT; t0 += kMlpChunkTiles) {
    c.tile0 = t0; c.ntiles = std::min(kMlpChunkTiles, m.MT - t0);
    const std::vector<LLayerH>& fl = tcl_split_choice(0, lp.cpair, c.ntiles) == 2 ? lp.layers_p : lp.layers;
    for (const LLayerH& L : fl) {
      LayerArgs a{};
      a.save = rec ? 1 : 0;
      if (L.type == LE_OUT) a.f0 = out + (size_t)t0 * 128 * m.O;
      else if (rec) {
        a.aux[0] = c.at(M_XH0 + L.layer); a.aux_tile_bytes = mm.tile[M_XH0 + L.layer];
        a.f0 = (float*)c.at(M_RSTD0 + L.layer);
      }
      if (int r = tcm_launch(c, L, a)) return r;
    }
  }
  return 0;
}

int tcm_bwd(const wmrl_mlp_dims* d, const void* packed, const void* x_panel, const void* saved, const float* g_out,
            const wmrl_mlp_grads* gw, float* g_x, void* ws, cudaStream_t st) {
  MlpShape m;
  WMRL_REQUIRE(mlp_shape(d, m), "mlp_bwd: shape not supported");
  if (m.MT == 0) return 0;
  LPlan lp = tcm_make_plan(m, nullptr);
  WMRL_REQUIRE(lp.ok, "mlp_bwd: shape not supported by the layer-wise tcgen05 path");
  const bool wg = (m.flags & WMRL_MLP_WGRAD) != 0;
  WMRL_REQUIRE(!wg || gw, "mlp_bwd: WMRL_MLP_WGRAD needs the gradient struct");
  const MlpMem mm = tcm_layout(m, 2);
  MlpCall c{&m, &mm, (uint8_t*)ws, (uint8_t*)const_cast<void*>(saved), (const uint8_t*)x_panel, nullptr, nullptr, 148, st};
  if (int r = tcl_device_setup(c.nsm)) return r;
  const uint8_t* img = (const uint8_t*)packed;
  c.data = img + lp.pk.off_data;
  c.ep = (const float*)(img + lp.pk.off_eparams);
  tcl_to_panel_kernel<<<dim3(1, (unsigned)m.MT, 1), 256, 0, st>>>(g_out, c.at(M_GOUT), (int)m.R, 1, m.O, 16, m.MT);
  WMRL_CUDA_OK(cudaGetLastError());
  for (int t0 = 0; t0 < m.MT; t0 += kMlpChunkTiles) {
    c.tile0 = t0; c.ntiles = std::min(kMlpChunkTiles, m.MT - t0);
    const std::vector<LLayerH>& bl = tcl_split_choice(0, lp.cpair, c.ntiles) == 2 ? lp.blayers_p : lp.blayers;
    for (const LLayerH& L : bl) {
      LayerArgs a{};
      if (L.type == LB_XGRAD) {
        if (!g_x) continue;
        a.f0 = g_x + (size_t)t0 * 128 * m.F;
        a.mode = (m.flags & WMRL_MLP_XGRAD_ACCUMULATE) ? 1 : 0;
      } else {
        a.aux[0] = c.at(M_XH0 + L.layer); a.aux[1] = wg ? c.at(M_GU0 + L.layer) : nullptr; a.aux[2] = c.at(M_Y0 + L.layer);
        a.aux_tile_bytes = mm.tile[M_XH0 + L.layer];
        a.f0 = (float*)c.at(M_RSTD0 + L.layer);
      }
      if (int r = tcm_launch(c, L, a)) return r;
    }
  }
  c.tile0 = 0; c.ntiles = 0;
  if (!wg) return 0;
  auto wgrad = [&](int abuf, int m_cols, int m_valid, int bbuf, int n_cols, int n_valid, float* dW, int ldw) -> int {
    if (!dW) return 0;
    Wg2Params q{};
    q.a_base = c.at(abuf); q.a_blk = mm.tile[abuf]; q.b_base = c.at(bbuf); q.b_blk = mm.tile[bbuf];
    q.nblocks = m.MT; q.m_cols = m_cols; q.m_valid = m_valid; q.n_cols = n_cols; q.n_valid = n_valid;
    q.dW = dW; q.ldw = ldw;
    q.mtiles = (m_cols + 127) / 128; q.ntiles = (n_cols + 255) / 256;
    const int tl = q.mtiles * q.ntiles;
    q.splits = std::max(1, std::min(m.MT, c.nsm / tl));
    tcl_wgrad_kernel<<<dim3(tl, q.splits), kWg2Threads, kWg2Smem, st>>>(q);
    WMRL_CUDA_OK(cudaGetLastError());
    return 0;
  };
  auto colsum = [&](int xbuf, int ybuf, int cols, int valid, float* out) -> int {
    if (!out) return 0;
    Cs2Params q{};
    q.x_base = c.at(xbuf); q.x_blk = mm.tile[xbuf];
    q.y_base = ybuf >= 0 ? c.at(ybuf) : nullptr; q.y_blk = ybuf >= 0 ? mm.tile[ybuf] : 0;
    q.nblocks = m.MT; q.valid = valid; q.out = out;
    const int chunks = cols / 8;
    const int ysplit = std::max(1, std::min((m.MT + kCs2Warps - 1) / kCs2Warps, (8 * c.nsm + chunks - 1) / chunks));
    tcl_colsum_kernel<<<dim3(chunks, ysplit), kCs2Warps * 32, 0, st>>>(q);
    WMRL_CUDA_OK(cudaGetLastError());
    return 0;
  };
  for (int l = 0; l < m.L; ++l) {
    if (l == 0) { if (int r = wgrad(M_X, m.Fp, m.F, M_GP0, m.Hh, m.Hh, gw->lin_w[0], m.F)) return r; }
    else if (int r = wgrad(M_Y0 + l - 1, m.Hh, m.Hh, M_GP0 + l, m.Hh, m.Hh, gw->lin_w[l], m.Hh)) return r;
    if (gw->lin_b[l] && gw->ln_b[l] && gw->ln_g[l]) {   // one pass over g_u, x-hat, g_pre
      Cs2Params q{};
      q.x_base = c.at(M_GU0 + l); q.x_blk = mm.tile[M_GU0 + l]; q.y_base = c.at(M_XH0 + l); q.y_blk = mm.tile[M_XH0 + l];
      q.z_base = c.at(M_GP0 + l); q.z_blk = mm.tile[M_GP0 + l];
      q.nblocks = m.MT; q.valid = m.Hh; q.out = gw->ln_b[l]; q.out2 = gw->ln_g[l]; q.out3 = gw->lin_b[l];
      const int chunks = m.Hh / 8;
      const int ysplit = std::max(1, std::min((m.MT + kCs2Warps - 1) / kCs2Warps, (8 * c.nsm + chunks - 1) / chunks));
      tcl_colsum_kernel<<<dim3(chunks, ysplit), kCs2Warps * 32, 0, st>>>(q);
      WMRL_CUDA_OK(cudaGetLastError());
    } else {
      if (int r = colsum(M_GP0 + l, -1, m.Hh, m.Hh, gw->lin_b[l])) return r;
      if (int r = colsum(M_GU0 + l, -1, m.Hh, m.Hh, gw->ln_b[l])) return r;
      if (int r = colsum(M_GU0 + l, M_XH0 + l, m.Hh, m.Hh, gw->ln_g[l])) return r;
    }
  }
  if (int r = wgrad(M_Y0 + m.L - 1, m.Hh, m.Hh, M_GOUT, 16, m.O, gw->out_w, m.Hh)) return r;
  if (int r = colsum(M_GOUT, -1, 16, m.O, gw->out_b)) return r;
  return 0;
}

}  // namespace wmrl
