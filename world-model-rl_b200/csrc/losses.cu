// World-model loss side of posterior filtering (WorldModel.update, world_model.py:128-141): the KL term between the
// prior and posterior sequences that observe_rollout returns, with the free-nats clamp, as ONE fused pass over the
// distribution parameters (the reference builds two torch.distributions objects, D.kl_divergence, torch.max, two
// means: ~20 elementwise launches over [B,T,.] tensors).  HBM-bound: each parameter is read once forward, once
// backward.
//   continuous: KL(N(m1,s1) || N(m2,s2)) summed over the S latent dims (D.Independent(Normal, 1), state/continuous.py:79-84)
//   discrete:   KL(Cat(l1) || Cat(l2)) summed over the C groups (Independent(OneHotCategoricalStraightThrough, 1),
//               state/discrete.py:77-79; the KL of one-hot categoricals is the categorical KL)
// Index 0 of the sequences (the initial state) is excluded, as get_dist() does.
#include <algorithm>

#include "common.cuh"

namespace wmrl {

__device__ __forceinline__ float kl_warp_sum(float v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}
__device__ __forceinline__ float kl_warp_max(float v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, s));
  return v;
}

// one warp per (b, t >= 1).  kBwd: write the parameter gradients of  scale * sum_rows max(kl_row, rho)
template <bool kDisc, bool kBwd>
__global__ void kl_kernel(const float* __restrict__ pa, const float* __restrict__ pb, const float* __restrict__ qa,
                          const float* __restrict__ qb, long long B, int T, int S, int C, float rho, float scale,
                          float* __restrict__ kl_out, float* __restrict__ sums, float* g_pa, float* g_pb, float* g_qa,
                          float* g_qb) {
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B * (T - 1)) return;
  const long long b = warp / (T - 1);
  const int t = (int)(warp % (T - 1)) + 1;
  const int W = kDisc ? C * S : S;
  const long long o = (b * T + t) * W;
  float kl = 0.f;
  if (!kDisc) {
    for (int j = lane; j < S; j += 32) {
      const float m1 = pa[o + j], s1 = pb[o + j], m2 = qa[o + j], s2 = qb[o + j];
      const float dm = m1 - m2, r = s1 / s2;
      // torch's kl_normal_normal: 0.5 * (var_ratio + t1 - 1 - log(var_ratio)), var_ratio = (s1/s2)^2, t1 = ((m1-m2)/s2)^2
      const float vr = r * r, t1 = (dm / s2) * (dm / s2);
      kl += 0.5f * (vr + t1 - 1.f - logf(vr));
    }
    kl = kl_warp_sum(kl);
  } else {
    for (int c = 0; c < C; ++c) {
      float part = 0.f;
      for (int j0 = 0; j0 < S; j0 += 32) {   // S <= 32 in every config; the loop keeps larger S correct when one chunk
        const int j = j0 + lane;             // holds a whole group (S <= 32) - asserted on the host
        const float l1 = j < S ? pa[o + c * S + j] : -INFINITY, l2 = j < S ? qa[o + c * S + j] : -INFINITY;
        const float mx1 = kl_warp_max(l1), mx2 = kl_warp_max(l2);
        const float e1 = j < S ? expf(l1 - mx1) : 0.f, e2 = j < S ? expf(l2 - mx2) : 0.f;
        const float z1 = kl_warp_sum(e1), z2 = kl_warp_sum(e2);
        const float lp1 = l1 - mx1 - logf(z1), lp2 = l2 - mx2 - logf(z2);
        part += j < S ? (e1 / z1) * (lp1 - lp2) : 0.f;
      }
      kl += kl_warp_sum(part);
    }
  }
  if (!kBwd) {
    if (lane == 0) {
      kl_out[b * (T - 1) + t - 1] = kl;
      atomicAdd(sums, fmaxf(kl, rho));
      atomicAdd(sums + 1, kl);
    }
    return;
  }
  const float g = (kl > rho) ? scale : 0.f;   // d max(kl, rho) / d kl
  if (!kDisc) {
    for (int j = lane; j < S; j += 32) {
      const float m1 = pa[o + j], s1 = pb[o + j], m2 = qa[o + j], s2 = qb[o + j];
      const float dm = m1 - m2, i2 = 1.f / (s2 * s2);
      g_pa[o + j] = g * dm * i2;
      g_qa[o + j] = -g * dm * i2;
      g_pb[o + j] = g * (s1 * i2 - 1.f / s1);
      g_qb[o + j] = g * (1.f / s2 - (s1 * s1 + dm * dm) * i2 / s2);
    }
  } else {
    for (int c = 0; c < C; ++c) {
      const int j = lane;
      const float l1 = j < S ? pa[o + c * S + j] : -INFINITY, l2 = j < S ? qa[o + c * S + j] : -INFINITY;
      const float mx1 = kl_warp_max(l1), mx2 = kl_warp_max(l2);
      const float e1 = j < S ? expf(l1 - mx1) : 0.f, e2 = j < S ? expf(l2 - mx2) : 0.f;
      const float z1 = kl_warp_sum(e1), z2 = kl_warp_sum(e2);
      const float p1 = e1 / z1, p2 = e2 / z2;
      const float d = (l1 - mx1 - logf(z1)) - (l2 - mx2 - logf(z2));
      const float klc = kl_warp_sum(j < S ? p1 * d : 0.f);
      if (j < S) {
        g_pa[o + c * S + j] = g * p1 * (d - klc);     // d KL / d l1_j = p_j (log p_j - log q_j - KL)
        g_qa[o + c * S + j] = g * (p2 - p1);          // d KL / d l2_j = q_j - p_j
      }
    }
  }
}

// Reward and done losses of WorldModel.update (world_model.py:130, 143-144) in one pass over the two heads' outputs:
//   reward: -log N(r; p, 1) = 0.5 (r - p)^2 + 0.5 log(2 pi)   (get_norm_dist: Independent(Normal(p, 1), 1))
//   done:   binary_cross_entropy_with_logits(l, y) = max(l, 0) - l y + log(1 + exp(-|l|))
// sums[0] += sum of reward terms, sums[1] += sum of done terms; the unscaled gradients d/dp = p - r and
// d/dl = sigmoid(l) - y are written in the same pass (the caller scales them by the upstream gradient / count).
__global__ void head_losses_kernel(const float* __restrict__ pr, const float* __restrict__ tr, long long nr,
                                   const float* __restrict__ ld, const float* __restrict__ td, long long nd,
                                   float* __restrict__ sums, float* __restrict__ g_pr, float* __restrict__ g_ld) {
  float a = 0.f, b = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nr; i += stride) {
    const float d = pr[i] - tr[i];
    a += 0.5f * d * d + 0.9189385332046727f;
    if (g_pr) g_pr[i] = d;
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nd; i += stride) {
    const float l = ld[i], y = td[i];
    const float e = expf(-fabsf(l));
    b += fmaxf(l, 0.f) - l * y + log1pf(e);
    if (g_ld) g_ld[i] = (l >= 0.f ? 1.f / (1.f + e) : e / (1.f + e)) - y;
  }
  a = kl_warp_sum(a); b = kl_warp_sum(b);
  __shared__ float sa[32], sb[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sa[warp] = a; sb[warp] = b; }
  __syncthreads();
  if (warp == 0) {
    a = lane < (int)(blockDim.x >> 5) ? sa[lane] : 0.f;
    b = lane < (int)(blockDim.x >> 5) ? sb[lane] : 0.f;
    a = kl_warp_sum(a); b = kl_warp_sum(b);
    if (lane == 0) { atomicAdd(sums, a); atomicAdd(sums + 1, b); }
  }
}

// get_features() of a state sequence (state/continuous.py:73-77, state/discrete.py): cat(deter[:, 1:], stoch[:, 1:], -1) as ONE
// streaming pass (torch.cat runs this strided gather at about a third of HBM bandwidth), and its backward: the gradient of the
// feature tensor scattered back into [B][T][D] / [B][T][S_in] gradients whose index 0 is zero.
template <int V, bool kBwd>
__global__ void features_kernel(float* __restrict__ deter, float* __restrict__ stoch, float* __restrict__ feats, long long B,
                                int T, int D, int S) {
  const int F = D + S;
  const long long nv = B * (long long)(T - 1) * (F / V);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
    const long long row = i / (F / V);
    const int col = (int)(i % (F / V)) * V;
    const long long b = row / (T - 1);
    const int t = (int)(row % (T - 1)) + 1;
    float* src = col < D ? deter + (b * T + t) * D + col : stoch + (b * T + t) * S + (col - D);
    float* dst = feats + row * F + col;
    if (V == 4) { if (kBwd) *reinterpret_cast<float4*>(src) = *reinterpret_cast<const float4*>(dst); else *reinterpret_cast<float4*>(dst) = *reinterpret_cast<const float4*>(src); }
    else if (V == 2) { if (kBwd) *reinterpret_cast<float2*>(src) = *reinterpret_cast<const float2*>(dst); else *reinterpret_cast<float2*>(dst) = *reinterpret_cast<const float2*>(src); }
    else { if (kBwd) *src = *dst; else *dst = *src; }
  }
  if (kBwd) {   // index 0 of both gradients
    const long long n0 = B * (long long)F;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n0; i += stride) {
      const long long b = i / F;
      const int col = (int)(i % F);
      if (col < D) deter[(b * T) * D + col] = 0.f; else stoch[(b * T) * S + (col - D)] = 0.f;
    }
  }
}

}  // namespace wmrl

using namespace wmrl;

extern "C" {

int wmrl_kl_loss_fwd(const wmrl_dims* d, const float* pri_a, const float* pri_b, const float* post_a, const float* post_b,
                     float rho, float* kl, float* sums, void* stream) {
  if (int r = check_dims(d, 1)) return r;
  Dims dd(*d);
  if (dd.B == 0 || dd.T <= 1) return 0;
  const bool disc = dd.variant == WMRL_DISCRETE;
  WMRL_REQUIRE(pri_a && post_a && kl && sums && (disc || (pri_b && post_b)), "kl_loss_fwd: NULL tensor");
  WMRL_REQUIRE(!disc || dd.S <= 32, "kl_loss: more than 32 classes per categorical group is not supported");
  const long long rows = (long long)dd.B * (dd.T - 1);
  const unsigned blocks = (unsigned)((rows * 32 + 255) / 256);
  if (disc) kl_kernel<true, false><<<blocks, 256, 0, (cudaStream_t)stream>>>(pri_a, nullptr, post_a, nullptr, dd.B, dd.T, dd.S, dd.C, rho, 0.f, kl, sums, nullptr, nullptr, nullptr, nullptr);
  else kl_kernel<false, false><<<blocks, 256, 0, (cudaStream_t)stream>>>(pri_a, pri_b, post_a, post_b, dd.B, dd.T, dd.S, 1, rho, 0.f, kl, sums, nullptr, nullptr, nullptr, nullptr);
  WMRL_CUDA_OK(cudaGetLastError());
  return 0;
}

int wmrl_kl_loss_bwd(const wmrl_dims* d, const float* pri_a, const float* pri_b, const float* post_a, const float* post_b,
                     float rho, float scale, float* g_pri_a, float* g_pri_b, float* g_post_a, float* g_post_b, void* stream) {
  if (int r = check_dims(d, 1)) return r;
  Dims dd(*d);
  if (dd.B == 0 || dd.T <= 1) return 0;
  const bool disc = dd.variant == WMRL_DISCRETE;
  WMRL_REQUIRE(pri_a && post_a && g_pri_a && g_post_a && (disc || (pri_b && post_b && g_pri_b && g_post_b)), "kl_loss_bwd: NULL tensor");
  WMRL_REQUIRE(!disc || dd.S <= 32, "kl_loss: more than 32 classes per categorical group is not supported");
  const long long rows = (long long)dd.B * (dd.T - 1);
  const unsigned blocks = (unsigned)((rows * 32 + 255) / 256);
  if (disc) kl_kernel<true, true><<<blocks, 256, 0, (cudaStream_t)stream>>>(pri_a, nullptr, post_a, nullptr, dd.B, dd.T, dd.S, dd.C, rho, scale, nullptr, nullptr, g_pri_a, nullptr, g_post_a, nullptr);
  else kl_kernel<false, true><<<blocks, 256, 0, (cudaStream_t)stream>>>(pri_a, pri_b, post_a, post_b, dd.B, dd.T, dd.S, 1, rho, scale, nullptr, nullptr, g_pri_a, g_pri_b, g_post_a, g_post_b);
  WMRL_CUDA_OK(cudaGetLastError());
  return 0;
}

int wmrl_head_losses(const float* pred_r, const float* tgt_r, int64_t n_reward, const float* logit_d, const float* tgt_d,
                     int64_t n_done, float* sums, float* g_pred_r, float* g_logit_d, void* stream) {
  WMRL_REQUIRE(n_reward >= 0 && n_done >= 0 && sums, "head_losses: bad arguments");
  WMRL_REQUIRE((n_reward == 0 || (pred_r && tgt_r)) && (n_done == 0 || (logit_d && tgt_d)), "head_losses: NULL tensor");
  const long long n = n_reward > n_done ? n_reward : n_done;
  if (n == 0) return 0;
  const unsigned blocks = (unsigned)std::min<long long>((n + 255) / 256, 148 * 8);
  head_losses_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(pred_r, tgt_r, n_reward, logit_d, tgt_d, n_done, sums, g_pred_r,
                                                                g_logit_d);
  WMRL_CUDA_OK(cudaGetLastError());
  return 0;
}

static int features_launch(float* deter, float* stoch, float* feats, int64_t B, int32_t T, int32_t D, int32_t S, bool bwd,
                           cudaStream_t st) {
  WMRL_REQUIRE(B >= 0 && T >= 1 && D >= 1 && S >= 1, "features: bad dims");
  if (B == 0 || T == 1) return 0;
  WMRL_REQUIRE(deter && stoch && feats, "features: NULL tensor");
  const int V = (D % 4 == 0 && S % 4 == 0) ? 4 : ((D % 2 == 0 && S % 2 == 0) ? 2 : 1);
  const long long nv = (long long)B * (T - 1) * ((D + S) / V);
  const unsigned blocks = (unsigned)std::min<long long>((nv + 255) / 256, 148 * 16);
#define WMRL_FEAT(VV)                                                                                   \
  do {                                                                                                  \
    if (bwd) features_kernel<VV, true><<<blocks, 256, 0, st>>>(deter, stoch, feats, B, T, D, S);        \
    else features_kernel<VV, false><<<blocks, 256, 0, st>>>(deter, stoch, feats, B, T, D, S);           \
  } while (0)
  if (V == 4) WMRL_FEAT(4); else if (V == 2) WMRL_FEAT(2); else WMRL_FEAT(1);
#undef WMRL_FEAT
  WMRL_CUDA_OK(cudaGetLastError());
  return 0;
}
int wmrl_features_fwd(const float* deter_seq, const float* stoch_seq, int64_t B, int32_t T, int32_t D, int32_t S_in,
                      float* features, void* stream) {
  return features_launch(const_cast<float*>(deter_seq), const_cast<float*>(stoch_seq), features, B, T, D, S_in, false,
                         (cudaStream_t)stream);
}
int wmrl_features_bwd(const float* g_features, int64_t B, int32_t T, int32_t D, int32_t S_in, float* g_deter_seq,
                      float* g_stoch_seq, void* stream) {
  return features_launch(g_deter_seq, g_stoch_seq, const_cast<float*>(g_features), B, T, D, S_in, true, (cudaStream_t)stream);
}

}  // extern "C"
