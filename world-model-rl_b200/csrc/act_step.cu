// Online acting step of the stateful policy (rssm_world_model/memory_actor.py:34-57) as ONE kernel:
//   post   = posterior(e_t, state)            rssm.py:70-82   [h | e] -> Linear -> ELU -> Linear -> sample
//   a_t    = actor([h | s_post], det.)        actor.py:47-52  La x (Linear -> LayerNorm -> ELU) -> mu -> tanh squash
//   state' = prior(a_t, post)                 rssm.py:53-68   [s_post | a] -> Linear -> ELU -> GRUCell -> Linear -> ELU
//                                                             -> Linear -> sample
// The reference runs this at B = 1 as ~45 ATen launches per environment step (launch-latency-bound: a few hundred
// microseconds); here a thread-block CLUSTER of 8 CTAs walks the 9 + La dependent layers of one row: every layer is a
// GEMV whose output neurons are dealt to the 128 warps of the cluster (fp32 weights streamed from L2 with all of a
// warp's loads in flight, fp32 FMA, warp-shuffle reduction), each result is written into ALL 8 CTAs' shared-memory
// copy of the activation vector through distributed shared memory, and one cluster barrier separates two layers.
// LayerNorm statistics, the activation functions and both latent samples are computed redundantly by every CTA from
// its own full copy (no further exchange).  fp32 throughout (same arithmetic class as the reference; summation order
// differs).  One cluster per batch row (grid = 8 x B).
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace wmrl {

constexpr int kActCluster = 8, kActThreads = 512, kActWarps = kActThreads / 32;

struct ActArgs {
  wmrl_rssm_params w;
  wmrl_actor_params aw;
  const float* obs_embed; const float* deter; const float* noise_post; const float* noise_pri;
  float* action; float* post_stoch; float* post_a; float* post_b;
  float* deter_next; float* pri_stoch; float* pri_a; float* pri_b;
  wmrl_act_sampling smp;   // std_w == NULL: deterministic action
  int variant, B, D, S, C, A, Hd, E, Ha, La, S_in, P, F;
  float inv_action_scale;
  long long ldo;   // row stride (floats) of every output tensor
  // shared-memory layout (float offsets)
  int o_he, o_hs, o_sa, o_y0, o_y1, o_hn, o_out, o_red, total;
};

__device__ __forceinline__ float act_warp_sum(float v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}
__device__ __forceinline__ float act_warp_max(float v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, s));
  return v;
}

// NR dot products of weight rows (global, K floats each) with one shared-memory vector; every lane returns the sums
template <int NR>
__device__ __forceinline__ void warp_dots(const float* const (&rows)[NR], const float* __restrict__ x, int K, float (&acc)[NR]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int r = 0; r < NR; ++r) acc[r] = 0.f;
  bool vec = (K & 3) == 0;
#pragma unroll
  for (int r = 0; r < NR; ++r) vec = vec && ((reinterpret_cast<uintptr_t>(rows[r]) & 15) == 0);
  if (vec) {
    const float4* x4 = reinterpret_cast<const float4*>(x);
#pragma unroll 4
    for (int k = lane; k < K / 4; k += 32) {
      float4 wv[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r) wv[r] = __ldg(reinterpret_cast<const float4*>(rows[r]) + k);
      const float4 xv = x4[k];
#pragma unroll
      for (int r = 0; r < NR; ++r)
        acc[r] = fmaf(wv[r].x, xv.x, fmaf(wv[r].y, xv.y, fmaf(wv[r].z, xv.z, fmaf(wv[r].w, xv.w, acc[r]))));
    }
  } else {
#pragma unroll 8
    for (int k = lane; k < K; k += 32) {
      const float xv = x[k];
#pragma unroll
      for (int r = 0; r < NR; ++r) acc[r] = fmaf(__ldg(rows[r] + k), xv, acc[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < NR; ++r) acc[r] = act_warp_sum(acc[r]);
}

struct ActCtx {
  cg::cluster_group cluster;
  float* smem;
  int rank, gwarp, lane;
};
// write one value into every CTA's copy of the activation vector
__device__ __forceinline__ void bcast_store(const ActCtx& c, int off, float v) {
#pragma unroll
  for (int r = 0; r < kActCluster; ++r) *c.cluster.map_shared_rank(c.smem + off, r) = v;
}

enum { kNone = 0, kElu = 1 };
// out[j] = act(W[j] . in + b[j]), j in [0, N): neurons dealt to the cluster's warps, 4 per pass
template <int kAct>
__device__ __forceinline__ void layer(const ActCtx& c, const float* __restrict__ W, const float* __restrict__ b, int N, int K,
                                      int in_off, int out_off) {
  const float* x = c.smem + in_off;
  constexpr int G = kActCluster * kActWarps;
  for (int j0 = c.gwarp * 4; j0 < N; j0 += G * 4) {
    const float* rows[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) rows[r] = W + (size_t)min(j0 + r, N - 1) * K;
    float acc[4];
    warp_dots<4>(rows, x, K, acc);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (j0 + r < N && c.lane == r) {
        float v = acc[r] + __ldg(b + j0 + r);
        if (kAct == kElu) v = v > 0.f ? v : expm1f(v);
        bcast_store(c, out_off + j0 + r, v);
      }
    }
  }
}

// LayerNorm (eps 1e-5, affine, two-pass variance) + ELU over smem[off, off + W), by the whole CTA, in place
__device__ __forceinline__ void ln_elu_inplace(const ActCtx& c, const ActArgs& p, int off, int W, const float* __restrict__ g,
                                               const float* __restrict__ be) {
  float* v = c.smem + off;
  float* red = c.smem + p.o_red;
  const int warp = threadIdx.x >> 5;
  float s = 0.f;
  for (int i = threadIdx.x; i < W; i += kActThreads) s += v[i];
  s = act_warp_sum(s);
  if (c.lane == 0) red[warp] = s;
  __syncthreads();
  float mean = 0.f;
  for (int i = 0; i < kActWarps; ++i) mean += red[i];
  mean /= (float)W;
  __syncthreads();
  float q = 0.f;
  for (int i = threadIdx.x; i < W; i += kActThreads) { const float d = v[i] - mean; q = fmaf(d, d, q); }
  q = act_warp_sum(q);
  if (c.lane == 0) red[warp] = q;
  __syncthreads();
  float var = 0.f;
  for (int i = 0; i < kActWarps; ++i) var += red[i];
  const float rstd = rsqrtf(var / (float)W + 1e-5f);
  for (int i = threadIdx.x; i < W; i += kActThreads) {
    const float u = (v[i] - mean) * rstd * __ldg(g + i) + __ldg(be + i);
    v[i] = u > 0.f ? u : expm1f(u);
  }
  __syncthreads();
}

// head output (P floats at smem[o_out]) -> sampled latent at smem[dst_off, +S_in) of this CTA; rank 0 writes the
// distribution parameters and the sample to global memory.  Same arithmetic as sample_cont_kernel / sample_disc_kernel.
__device__ __forceinline__ void sample_state(const ActCtx& c, const ActArgs& p, int row, const float* __restrict__ noise,
                                             int dst_off, int dst_off2, float* g_stoch, float* g_a, float* g_b) {
  const float* out = c.smem + p.o_out;
  const bool wr = c.rank == 0;
  if (p.variant == WMRL_CONTINUOUS) {
    for (int j = threadIdx.x; j < p.S; j += kActThreads) {
      const float m = out[j], raw = out[p.S + j];
      const float sp = raw > 20.f ? raw : log1pf(expf(raw));
      const float sd = sp + 0.1f;
      const float s = m + sd * __ldg(noise + (size_t)row * p.S + j);
      c.smem[dst_off + j] = s;
      if (dst_off2 >= 0) c.smem[dst_off2 + j] = s;
      if (wr) { g_a[row * p.ldo + j] = m; g_b[row * p.ldo + j] = sd; g_stoch[row * p.ldo + j] = s; }
    }
  } else {
    const int warp = threadIdx.x >> 5, S = p.S;
    for (int g = warp; g < p.C; g += kActWarps) {
      const float* l = out + g * S;
      const float* qq = noise + (size_t)row * p.S_in + g * S;
      float mx = -INFINITY;
      for (int j = c.lane; j < S; j += 32) mx = fmaxf(mx, l[j]);
      mx = act_warp_max(mx);
      float se = 0.f;
      for (int j = c.lane; j < S; j += 32) se += expf(l[j] - mx);
      const float lse = mx + logf(act_warp_sum(se));
      float mx2 = -INFINITY;
      for (int j = c.lane; j < S; j += 32) mx2 = fmaxf(mx2, l[j] - lse);
      mx2 = act_warp_max(mx2);
      float se2 = 0.f;
      for (int j = c.lane; j < S; j += 32) se2 += expf((l[j] - lse) - mx2);
      se2 = act_warp_sum(se2);
      float best = -INFINITY;
      int bi = 0x7fffffff;
      for (int j = c.lane; j < S; j += 32) {
        const float pj = expf((l[j] - lse) - mx2) / se2;
        const float v = pj / __ldg(qq + j);
        if (v > best) { best = v; bi = j; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      for (int j = c.lane; j < S; j += 32) {
        const float oh = (j == bi) ? 1.f : 0.f;
        c.smem[dst_off + g * S + j] = oh;
        if (dst_off2 >= 0) c.smem[dst_off2 + g * S + j] = oh;
        if (wr) { g_a[row * p.ldo + g * S + j] = l[j]; g_stoch[row * p.ldo + g * S + j] = oh; }
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kActThreads, 1) act_step_kernel(const __grid_constant__ ActArgs p) {
  extern __shared__ __align__(16) float act_smem[];
  ActCtx c{cg::this_cluster(), act_smem, 0, 0, (int)(threadIdx.x & 31)};
  c.rank = (int)c.cluster.block_rank();
  c.gwarp = c.rank * kActWarps + (int)(threadIdx.x >> 5);
  const int row = blockIdx.y;
  const wmrl_rssm_params& w = p.w;
  const wmrl_actor_params& aw = p.aw;
  // ---- inputs -> this CTA's copies: [h | e], [h | .], zero padding nowhere needed (exact widths)
  for (int i = threadIdx.x; i < p.D; i += kActThreads) {
    const float h = __ldg(p.deter + (size_t)row * p.D + i);
    c.smem[p.o_he + i] = h;
    c.smem[p.o_hs + i] = h;
  }
  for (int i = threadIdx.x; i < p.E; i += kActThreads) c.smem[p.o_he + p.D + i] = __ldg(p.obs_embed + (size_t)row * p.E + i);
  __syncthreads();
  c.cluster.sync();   // every CTA of the cluster is running before the first remote store
  // ---- posterior (rssm.py:70-82)
  layer<kElu>(c, w.post0_w, w.post0_b, p.Hd, p.D + p.E, p.o_he, p.o_y0);
  c.cluster.sync();
  layer<kNone>(c, w.post2_w, w.post2_b, p.P, p.Hd, p.o_y0, p.o_out);
  c.cluster.sync();
  sample_state(c, p, row, p.noise_post, p.o_hs + p.D, p.o_sa, p.post_stoch, p.post_a, p.post_b);
  // ---- actor (actor.py:47-52, deterministic)
  int in_off = p.o_hs, K = p.F, cur = p.o_y0;
  for (int l = 0; l < p.La; ++l) {
    // the last readers of `cur` in any CTA finished before an earlier cluster barrier
    layer<kNone>(c, aw.lin_w[l], aw.lin_b[l], p.Ha, K, in_off, cur);
    c.cluster.sync();
    ln_elu_inplace(c, p, cur, p.Ha, aw.ln_g[l], aw.ln_b[l]);
    in_off = cur; K = p.Ha;
    cur = (cur == p.o_y0) ? p.o_y1 : p.o_y0;
  }
  {
    // mu head + tanh squash -> action (every CTA's [s | a] buffer; global by the computing warp).  With the sampling
    // block (actor.py:54-59 and its use-site: rsample, clamp to the bound, entropy) warps A..2A-1 evaluate the stddev
    // head; the pair (mu_j, std_j) meets in the warp that owns mu_j through this CTA's own scratch copy.
    const float* x = c.smem + in_off;
    const bool smp = p.smp.std_w != nullptr;
    const int nwarp = kActCluster * kActWarps;
    if (smp) {
      // std_j first, broadcast into every CTA's scratch (o_out is free here), then one barrier
      for (int j = c.gwarp; j < p.A; j += nwarp) {
        const float* rows[1] = {p.smp.std_w + (size_t)j * p.Ha};
        float acc[1];
        warp_dots<1>(rows, x, p.Ha, acc);
        if (c.lane == 0) {
          const float z = acc[0] + __ldg(p.smp.std_b + j);
          bcast_store(c, p.o_out + j, 1.f / (1.f + expf(-z)) + 0.01f);
        }
      }
      c.cluster.sync();
    }
    for (int j = c.gwarp; j < p.A; j += nwarp) {
      const float* rows[1] = {aw.mu_w + (size_t)j * p.Ha};
      float acc[1];
      warp_dots<1>(rows, x, p.Ha, acc);
      if (c.lane == 0) {
        const float bd = __ldg(aw.bound + j);
        float a = tanhf((acc[0] + __ldg(aw.mu_b + j)) * p.inv_action_scale) * bd;
        if (smp) {
          const float sd = c.smem[p.o_out + j];
          p.smp.act_mean[row * p.ldo + j] = a;
          p.smp.act_std[row * p.ldo + j] = sd;
          a = fmaxf(fminf(a + sd * __ldg(p.smp.noise_action + (size_t)row * p.A + j), bd), -bd);
        }
        bcast_store(c, p.o_sa + p.S_in + j, a);
        p.action[row * p.ldo + j] = a;
      }
    }
    if (smp && c.gwarp == 0 && p.smp.entropy) {
      // H = sum_j 0.5 + 0.5 log(2 pi) + log std_j  (Independent(Normal, 1).entropy())
      float h = 0.f;
      for (int j = c.lane; j < p.A; j += 32) h += 1.4189385332046727f + logf(c.smem[p.o_out + j]);
      h = act_warp_sum(h);
      if (c.lane == 0) p.smp.entropy[row * p.ldo] = h;
    }
  }
  c.cluster.sync();
  // ---- prior (rssm.py:53-68): embed -> GRUCell -> hidden -> head -> sample
  layer<kElu>(c, w.embed_w, w.embed_b, p.D, p.S_in + p.A, p.o_sa, cur);
  c.cluster.sync();
  {
    const float* x = c.smem + cur;
    const float* h = c.smem + p.o_he;
    const int D = p.D;
    for (int j = c.gwarp; j < D; j += kActCluster * kActWarps) {
      const float* ri[3] = {w.w_ih + (size_t)j * D, w.w_ih + (size_t)(D + j) * D, w.w_ih + (size_t)(2 * D + j) * D};
      const float* rh[3] = {w.w_hh + (size_t)j * D, w.w_hh + (size_t)(D + j) * D, w.w_hh + (size_t)(2 * D + j) * D};
      float ai[3], ah[3];
      warp_dots<3>(ri, x, D, ai);
      warp_dots<3>(rh, h, D, ah);
      if (c.lane == 0) {
        const float r = 1.f / (1.f + expf(-(ai[0] + __ldg(w.b_ih + j) + ah[0] + __ldg(w.b_hh + j))));
        const float z = 1.f / (1.f + expf(-(ai[1] + __ldg(w.b_ih + D + j) + ah[1] + __ldg(w.b_hh + D + j))));
        const float n = tanhf(ai[2] + __ldg(w.b_ih + 2 * D + j) + r * (ah[2] + __ldg(w.b_hh + 2 * D + j)));
        const float hn = (1.f - z) * n + z * h[j];
        bcast_store(c, p.o_hn + j, hn);
        p.deter_next[row * p.ldo + j] = hn;
      }
    }
  }
  c.cluster.sync();
  const int hid = (cur == p.o_y0) ? p.o_y1 : p.o_y0;
  layer<kElu>(c, w.prior0_w, w.prior0_b, p.Hd, p.D, p.o_hn, hid);
  c.cluster.sync();
  layer<kNone>(c, w.prior2_w, w.prior2_b, p.P, p.Hd, hid, p.o_out);
  c.cluster.sync();
  sample_state(c, p, row, p.noise_pri, p.o_hs + p.D, -1, p.pri_stoch, p.pri_a, p.pri_b);
  c.cluster.sync();   // no CTA exits while a peer may still write into its shared memory
}

}  // namespace wmrl

using namespace wmrl;

extern "C" {

int wmrl_act_step(const wmrl_dims* d, const wmrl_rssm_params* w, const wmrl_actor_params* aw, const float* obs_embed,
                  const float* deter, const float* noise_post, const float* noise_prior, float* action,
                  float* post_stoch, float* post_dist_a, float* post_dist_b, float* deter_next, float* prior_stoch,
                  float* prior_dist_a, float* prior_dist_b, int64_t ld_out, const wmrl_act_sampling* sampling, void* stream) {
  if (int r = check_dims(d, 0)) return r;
  WMRL_REQUIRE(w && aw, "act_step: NULL weights");
  Dims dd(*d);
  if (dd.B == 0) return 0;
  const bool disc = dd.variant == WMRL_DISCRETE;
  WMRL_REQUIRE(obs_embed && deter && noise_post && noise_prior && action && post_stoch && post_dist_a && deter_next &&
                   prior_stoch && prior_dist_a && (disc || (post_dist_b && prior_dist_b)), "act_step: NULL tensor");
  WMRL_REQUIRE(dd.E >= 1 && dd.B <= 65535, "act_step: bad dims");
  WMRL_REQUIRE(ld_out >= std::max(std::max(dd.D, dd.S_in), std::max(dd.P, dd.A)), "act_step: ld_out smaller than an output row");
  WMRL_REQUIRE(w->post0_w && w->post0_b && w->post2_w && w->post2_b && w->embed_w && w->w_ih && w->w_hh && w->prior0_w &&
                   w->prior2_w && aw->mu_w && aw->mu_b && aw->bound, "act_step: NULL parameter");
  ActArgs a{};
  a.w = *w; a.aw = *aw;
  a.obs_embed = obs_embed; a.deter = deter; a.noise_post = noise_post; a.noise_pri = noise_prior;
  a.action = action; a.post_stoch = post_stoch; a.post_a = post_dist_a; a.post_b = post_dist_b;
  a.deter_next = deter_next; a.pri_stoch = prior_stoch; a.pri_a = prior_dist_a; a.pri_b = prior_dist_b;
  a.variant = dd.variant; a.B = dd.B; a.D = dd.D; a.S = dd.S; a.C = dd.C; a.A = dd.A; a.Hd = dd.Hd; a.E = dd.E;
  a.Ha = dd.Ha; a.La = dd.La; a.S_in = dd.S_in; a.P = dd.P; a.F = dd.F;
  a.inv_action_scale = 1.f / dd.action_scale;
  a.ldo = ld_out;
  if (sampling && sampling->std_w) {
    WMRL_REQUIRE(sampling->std_b && sampling->noise_action && sampling->act_mean && sampling->act_std, "act_step: incomplete sampling block");
    WMRL_REQUIRE(dd.A <= dd.P, "act_step: sampling needs A <= P (scratch)");
    a.smp = *sampling;
  }
  auto al4 = [](int x) { return (x + 3) / 4 * 4; };
  int o = 0;
  a.o_he = o; o += al4(dd.D + dd.E);
  a.o_hs = o; o += al4(dd.D + dd.S_in);
  a.o_sa = o; o += al4(dd.S_in + dd.A);
  const int wmax = std::max(std::max(dd.Ha, dd.Hd), dd.D);
  a.o_y0 = o; o += al4(wmax);
  a.o_y1 = o; o += al4(wmax);
  a.o_hn = o; o += al4(dd.D);
  a.o_out = o; o += al4(dd.P);
  a.o_red = o; o += 32;
  a.total = o;
  const size_t smem = (size_t)o * sizeof(float);
  WMRL_REQUIRE(smem <= 200 * 1024, "act_step: activation vectors do not fit shared memory");
  { int dev = -1; WMRL_CUDA_OK(cudaGetDevice(&dev)); }
  WMRL_CUDA_OK(cudaFuncSetAttribute(act_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(kActCluster, (unsigned)dd.B, 1);
  cfg.blockDim = dim3(kActThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kActCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  WMRL_CUDA_OK(cudaLaunchKernelEx(&cfg, act_step_kernel, a));
  return 0;
}

}  // extern "C"
