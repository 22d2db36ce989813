/*
 * wmrl.h - C ABI of the B200-native RSSM latent-imagination path.
 *
 * One shared object (libwmrl.so, sm_100a) replaces the ATen op chains that the
 * reference (mauicv/world-model-rl, package `reflect`) executes in
 *   - RSSMBase.imagine_rollout   src/reflect/components/rssm_world_model/rssm.py:123-140
 *   - RSSMBase.observe_rollout   src/reflect/components/rssm_world_model/rssm.py:93-121
 *   - Actor.forward(det.)        src/reflect/components/models/actor.py:47-52
 *   - ValueGradTrainer lambda-return loops
 *                                src/reflect/components/trainers/value/value_trainer.py:62-134
 * and, either side of the path,
 *   - the MLP heads over the features src/reflect/components/rssm_world_model/models.py:3-37,
 *                                src/reflect/components/trainers/value/critic.py:4-29
 *   - WorldModel.update's losses src/reflect/components/rssm_world_model/world_model.py:128-150
 *   - get_features()             src/reflect/components/rssm_world_model/state/continuous.py:73-77
 *   - WorldModelActor.__call__   src/reflect/components/rssm_world_model/memory_actor.py:34-57
 * The reference has no FFI of its own (pure Python over torch); the binding a
 * maintainer adds is the ctypes stub shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain C types only; every pointer is a BORROWED CUDA device pointer that
 *     outlives the call; nothing is allocated or freed inside;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*);
 *   - return 0 on success, non-zero on error with the text in wmrl_last_error();
 *   - all tensors are contiguous fp32 unless stated; weights are in torch
 *     nn.Linear / nn.GRUCell layout, i.e. weight[out][in] row-major, GRU gate
 *     order r,z,n (rssm.py:28-42);
 *   - symbols: B rows, D deter_size, S stoch_size, C num_categories (1 for the
 *     continuous variant), S_in = S (continuous) or C*S (discrete), P = 2S or
 *     C*S, Hd hidden_size, E obs_embed_size, A action dim, Ha/La actor hidden
 *     width / number of [Linear,LayerNorm,ELU] blocks, F = D + S_in.
 *   - one host thread per device context; not re-entrant.
 */
#ifndef WMRL_H_
#define WMRL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WMRL_ABI_VERSION 3
#if defined(__GNUC__)
#define WMRL_API __attribute__((visibility("default")))
#else
#define WMRL_API
#endif
#define WMRL_MAX_ACTOR_LAYERS 8

enum { WMRL_CONTINUOUS = 0, WMRL_DISCRETE = 1 };
enum { WMRL_FP32 = 0, WMRL_BF16 = 1 };
/* dims.flags */
enum {
  WMRL_FLAG_WM_WGRAD = 1,   /* imagine_bwd: also produce RSSM weight grads (WM not frozen) */
  WMRL_FLAG_SAVE_FOR_BWD = 2, /* imagine_fwd / observe_fwd: fill `saved` for a later *_bwd  */
  WMRL_FLAG_TRACE = 4,        /* imagine_fwd (bf16): write a cycle timeline into the workspace (tools/trace_tc.py) */
  WMRL_FLAG_BACKWARD = 8      /* wmrl_workspace_bytes / imagine_bwd: size the workspace for the BACKWARD call (the
                                 bf16 BPTT keeps its gradient stash there; the forward does not need it)       */
};

typedef struct wmrl_dims {
  int32_t variant;      /* WMRL_CONTINUOUS | WMRL_DISCRETE                               */
  int32_t precision;    /* WMRL_FP32 (SIMT validation mode) | WMRL_BF16 (tcgen05 path)    */
  int32_t B;            /* rows: start states (imagine) or sequences (observe)            */
  int32_t T;            /* imagine: n_steps H.  observe: sequence length T (T-1 steps)    */
  int32_t D, S, C, A, Hd, E;
  int32_t Ha, La;       /* actor hidden width / blocks (ignored by observe)               */
  float   action_scale; /* Actor.action_scale, actor.py:14                                */
  int32_t flags;
} wmrl_dims;

/* RSSMBase.__init__ parameters, rssm.py:27-42.  state_dict names in comments. */
typedef struct wmrl_rssm_params {
  const float* embed_w;   /* fc_state_action_embed.weight [D, S_in+A] */
  const float* embed_b;   /* fc_state_action_embed.bias   [D]         */
  const float* w_ih;      /* rnn.weight_ih [3D, D]                    */
  const float* w_hh;      /* rnn.weight_hh [3D, D]                    */
  const float* b_ih;      /* rnn.bias_ih   [3D]                       */
  const float* b_hh;      /* rnn.bias_hh   [3D]                       */
  const float* prior0_w;  /* state_prior.0.weight [Hd, D]             */
  const float* prior0_b;  /* state_prior.0.bias   [Hd]                */
  const float* prior2_w;  /* state_prior.2.weight [P, Hd]             */
  const float* prior2_b;  /* state_prior.2.bias   [P]                 */
  const float* post0_w;   /* state_posterior.0.weight [Hd, D+E]       */
  const float* post0_b;   /* state_posterior.0.bias   [Hd]            */
  const float* post2_w;   /* state_posterior.2.weight [P, Hd]         */
  const float* post2_b;   /* state_posterior.2.bias   [P]             */
} wmrl_rssm_params;

/* Same fields, writable: gradients are ACCUMULATED (+=) into them; the caller
 * zero-fills.  Any pointer may be NULL (that gradient is skipped). */
typedef struct wmrl_rssm_grads {
  float *embed_w, *embed_b, *w_ih, *w_hh, *b_ih, *b_hh, *prior0_w, *prior0_b, *prior2_w,
      *prior2_b, *post0_w, *post0_b, *post2_w, *post2_b;
} wmrl_rssm_grads;

/* Actor parameters, actor.py:22-39.  lin_* / ln_* index l is block l:
 * layers.{3l}.weight [Ha, F or Ha], layers.{3l}.bias, layers.{3l+1}.weight/bias. */
typedef struct wmrl_actor_params {
  const float* lin_w[WMRL_MAX_ACTOR_LAYERS];
  const float* lin_b[WMRL_MAX_ACTOR_LAYERS];
  const float* ln_g[WMRL_MAX_ACTOR_LAYERS];
  const float* ln_b[WMRL_MAX_ACTOR_LAYERS];
  const float* mu_w;    /* mu.weight [A, Ha] */
  const float* mu_b;    /* mu.bias   [A]     */
  const float* bound;   /* Actor.bound broadcast to [A] (actor.py:20) */
} wmrl_actor_params;

typedef struct wmrl_actor_grads {
  float* lin_w[WMRL_MAX_ACTOR_LAYERS];
  float* lin_b[WMRL_MAX_ACTOR_LAYERS];
  float* ln_g[WMRL_MAX_ACTOR_LAYERS];
  float* ln_b[WMRL_MAX_ACTOR_LAYERS];
  float* mu_w;
  float* mu_b;
} wmrl_actor_grads;

WMRL_API int wmrl_abi_version(void);
WMRL_API const char* wmrl_last_error(void);

/* 1 if the running device can execute the sm_100a tcgen05 path, else 0. */
WMRL_API int wmrl_device_supports_bf16(void);

/* 1 if WMRL_BF16 can also TRAIN at this shape: wmrl_imagine_fwd with WMRL_FLAG_SAVE_FOR_BWD records a bf16
 * activation stash from inside the persistent rollout kernel and wmrl_imagine_bwd(precision WMRL_BF16) runs the
 * persistent tcgen05 BPTT kernel + weight-gradient GEMMs over it (autograd of rssm.py:123-140, actor gradients
 * only: the world model is frozen by the caller, world_model.py:172).  0: record with WMRL_FP32. */
WMRL_API int wmrl_bf16_train_supported(const wmrl_dims* d);

/* Buffer sizes the caller must provide.  op: 0 imagine, 1 observe.  They depend on d->precision and d->flags
 * (the bf16 training path keeps a bf16 stash in `saved` and a gradient stash in the workspace). */
WMRL_API size_t wmrl_saved_bytes(const wmrl_dims* d, int op);
WMRL_API size_t wmrl_workspace_bytes(const wmrl_dims* d, int op);
/* Bytes of the packed bf16 weight image used by the WMRL_BF16 path (0 if the
 * shape is not supported by the tcgen05 kernel). */
WMRL_API size_t wmrl_packed_bytes(const wmrl_dims* d);
/* fp32 torch-layout params -> bf16 MMA-panel image (refreshed by the host
 * wrapper whenever a parameter's version counter changes). */
WMRL_API int wmrl_pack_weights(const wmrl_dims* d, const wmrl_rssm_params* w, const wmrl_actor_params* aw,
                      void* packed, void* stream);

/*
 * RSSMBase.imagine_rollout (rssm.py:123-140): H steps of
 *   a_t = Actor(cat(h_t, s_t).detach(), deterministic=True)     actor.py:47-52
 *   h_{t+1}, out = prior(a_t, (h_t, s_t))                       rssm.py:64-67
 *   s_{t+1} = sample(out, noise[t])                             rssm.py:179-181 / 223-227
 * noise: continuous eps[H][B][S] ~ N(0,1); discrete q[H][B][C][S] ~ Exp(1)
 *        (index = argmax(p/q), torch.multinomial's single-draw rule).
 * Outputs (index 0 of deter_seq/stoch_seq is filled from deter0/stoch0; index 0
 * of dist_a/dist_b is left to the caller, which owns the initial state):
 *   deter_seq[B][H+1][D], stoch_seq[B][H+1][S_in],
 *   dist_a = means[B][H+1][S] | logits[B][H+1][C][S], dist_b = stds[B][H+1][S] | NULL,
 *   actions[B][H][A], idx[B][H][C] int32 (discrete, else NULL).
 */
WMRL_API int wmrl_imagine_fwd(const wmrl_dims* d, const wmrl_rssm_params* w, const wmrl_actor_params* aw,
                     const void* packed, const float* deter0, const float* stoch0,
                     const float* noise, float* deter_seq, float* stoch_seq, float* dist_a,
                     float* dist_b, float* actions, int32_t* idx, void* saved, size_t saved_bytes,
                     void* ws, size_t ws_bytes, void* stream);

/*
 * Hand-written BPTT of the above.  d->precision must be the precision the forward recorded with (`saved` is the
 * fp32 activation record of the fp32 chain or the bf16 stash of the tcgen05 kernel); WMRL_BF16 needs `packed`.  Actor input is detached (rssm.py:135): actor
 * weights get gradient only through a_t -> dynamics.  g_* inputs are the
 * upstream gradients of the forward outputs (same shapes; index 0 slots of
 * g_deter_seq / g_stoch_seq flow straight to g_deter0 / g_stoch0).  Any g_*
 * input may be NULL (treated as zero).  gw may be NULL unless
 * WMRL_FLAG_WM_WGRAD is set.
 */
WMRL_API int wmrl_imagine_bwd(const wmrl_dims* d, const wmrl_rssm_params* w, const wmrl_actor_params* aw,
                     const void* packed, const float* noise, const float* deter_seq,
                     const float* stoch_seq, const float* dist_a, const float* dist_b,
                     const float* actions, const void* saved, const float* g_deter_seq,
                     const float* g_stoch_seq, const float* g_dist_a, const float* g_dist_b,
                     const wmrl_actor_grads* gaw, const wmrl_rssm_grads* gw, float* g_deter0,
                     float* g_stoch0, void* ws, size_t ws_bytes, void* stream);

/*
 * RSSMBase.observe_rollout (rssm.py:93-121) from the zero initial state:
 * for t in 0..T-2: prior_{t+1} = prior(a_t, post_t); post_{t+1} =
 * posterior(e_{t+1}, prior_{t+1}) (rssm.py:84-91).  obs_embeds[B][T][E],
 * actions[B][T][A]; noise_prior / noise_post [T-1][B][S] or [T-1][B][C][S].
 * Outputs are the two sequences, fields [B][T][...]; index 0 is zero-filled
 * for deter/stoch and left to the caller for the dist fields
 * (rssm.py:166-172 / 215-220: discrete initial logits are torch.randn).
 */
WMRL_API int wmrl_observe_fwd(const wmrl_dims* d, const wmrl_rssm_params* w, const float* obs_embeds,
                     const float* actions, const float* noise_prior, const float* noise_post,
                     float* pri_deter, float* pri_stoch, float* pri_dist_a, float* pri_dist_b,
                     float* post_deter, float* post_stoch, float* post_dist_a, float* post_dist_b,
                     void* saved, size_t saved_bytes, void* ws, size_t ws_bytes, void* stream);

/*
 * The same forward on the bf16 layer-wise tcgen05 chain.  The RSSM weights (prior AND posterior heads, plus their
 * transposes for the backward) are packed once per version with wmrl_pack_observe_weights into a buffer of
 * wmrl_observe_packed_bytes (0: shape not supported); d->precision = WMRL_BF16; buffer sizes from
 * wmrl_workspace_bytes(d, 1) / wmrl_saved_bytes(d, 1) with the flags of the call.  With WMRL_FLAG_SAVE_FOR_BWD the
 * per-step activation panels are kept in `saved` for wmrl_observe_bwd_bf16.  Same tensors and index conventions as
 * wmrl_observe_fwd.
 */
WMRL_API size_t wmrl_observe_packed_bytes(const wmrl_dims* d);
WMRL_API int wmrl_pack_observe_weights(const wmrl_dims* d, const wmrl_rssm_params* w, void* packed, void* stream);
WMRL_API int wmrl_observe_fwd_bf16(const wmrl_dims* d, const void* packed, const float* obs_embeds, const float* actions,
                          const float* noise_prior, const float* noise_post, float* pri_deter, float* pri_stoch,
                          float* pri_dist_a, float* pri_dist_b, float* post_deter, float* post_stoch, float* post_dist_a,
                          float* post_dist_b, void* saved, size_t saved_bytes, void* ws, size_t ws_bytes, void* stream);
/*
 * Full backward of the above (world-model training, world_model.py:120-163) on the same chain: dgrad GEMMs with fused
 * epilogues backwards in time, RSSM weight gradients as big-K tcgen05 GEMMs over the recorded panels, d obs_embeds and
 * d actions.  d->flags must carry WMRL_FLAG_BACKWARD (workspace sizing).  g_deter = dL/d pri_deter + dL/d post_deter
 * (both sequences hold the same deterministic states); any other g_* may be NULL.  Gradients are ACCUMULATED into gw
 * (caller zero-fills); g_obs_embeds [B][T][E] / g_actions [B][T][A] are written for the entries the scan reads
 * (obs 1..T-1, actions 0..T-2; the caller zero-fills the rest); either may be NULL.
 */
WMRL_API int wmrl_observe_bwd_bf16(const wmrl_dims* d, const void* packed, const float* noise_prior, const float* noise_post,
                          const float* pri_deter, const float* pri_dist_a, const float* pri_dist_b, const float* post_dist_a,
                          const float* post_dist_b, const void* saved, const float* g_deter, const float* g_pri_stoch,
                          const float* g_pri_dist_a, const float* g_pri_dist_b, const float* g_post_stoch,
                          const float* g_post_dist_a, const float* g_post_dist_b, const wmrl_rssm_grads* gw,
                          float* g_obs_embeds, float* g_actions, void* ws, size_t ws_bytes, void* stream);

/* Full backward (weight grads + d obs_embeds + d actions) of observe_fwd, as
 * needed by WorldModel.update (world_model.py:120-163). */
WMRL_API int wmrl_observe_bwd(const wmrl_dims* d, const wmrl_rssm_params* w, const float* obs_embeds,
                     const float* actions, const float* noise_prior, const float* noise_post,
                     const float* pri_deter, const float* pri_stoch, const float* pri_dist_a,
                     const float* pri_dist_b, const float* post_deter, const float* post_stoch,
                     const float* post_dist_a, const float* post_dist_b, const void* saved,
                     const float* g_pri_deter, const float* g_pri_stoch, const float* g_pri_dist_a,
                     const float* g_pri_dist_b, const float* g_post_deter, const float* g_post_stoch,
                     const float* g_post_dist_a, const float* g_post_dist_b,
                     const wmrl_rssm_grads* gw, float* g_obs_embeds, float* g_actions, void* ws,
                     size_t ws_bytes, void* stream);

/*
 * KL(prior || posterior) of the two sequences wmrl_observe_fwd returns, with the free-nats clamp of WorldModel.update
 * (world_model.py:131-141): kl[B][T-1] per (row, t >= 1) summed over the latent (Independent(Normal, 1) /
 * Independent(OneHotCategoricalStraightThrough, 1), state/continuous.py:79-84, state/discrete.py:77-79);
 * sums[0] += sum max(kl, rho), sums[1] += sum kl (the caller zero-fills sums and divides by B (T-1)).
 * *_a = means | logits, *_b = stds | NULL, all [B][T][.] as produced by the observe calls (index 0 is skipped).
 * bwd: gradients of  scale * sum_rows max(kl_row, rho)  w.r.t. the four parameter tensors (index 0 untouched: the
 * caller zero-fills); scale = upstream gradient / (B (T-1)).
 */
WMRL_API int wmrl_kl_loss_fwd(const wmrl_dims* d, const float* pri_a, const float* pri_b, const float* post_a,
                     const float* post_b, float rho, float* kl, float* sums, void* stream);
WMRL_API int wmrl_kl_loss_bwd(const wmrl_dims* d, const float* pri_a, const float* pri_b, const float* post_a,
                     const float* post_b, float rho, float scale, float* g_pri_a, float* g_pri_b, float* g_post_a,
                     float* g_post_b, void* stream);

/*
 * One online acting step of the stateful policy (rssm_world_model/memory_actor.py:34-57) as ONE kernel, fp32:
 *   post = posterior(obs_embed, state) (rssm.py:70-82), action = actor([h | s_post], deterministic) (actor.py:47-52),
 *   state' = prior(action, post) (rssm.py:53-68).
 * d.B rows (B = 1 in the reference; one 8-CTA cluster per row), d.T ignored.  obs_embed [B][E], deter [B][D] (the
 * carried state's deterministic part: its stochastic part is not read by the posterior), noise_post / noise_prior
 * [B][S] N(0,1) draws (continuous) or [B][C][S] Exp(1) draws (discrete).  Outputs: action [B][A]; the posterior
 * state's sample [B][S_in] and distribution (dist_a = mean | logits, dist_b = std | NULL); the next carried state:
 * deter_next [B][D], prior sample and distribution.  Every output is read as rows of stride ld_out floats (>= its
 * width), so a caller may hand slices of ONE allocation.  The posterior sample is drawn before the action, the
 * prior sample after it, as in the reference.
 */
/* Optional action sampling of the step (the Normal head, actor.py:54-59, with the rsample / clamp / entropy of its
 * use-site): std = sigmoid(stddev(hid)) + 0.01, action = clamp(mean + std * noise_action, -bound, bound),
 * entropy = sum_j 0.5 + 0.5 log(2 pi) + log std_j.  sampling == NULL or std_w == NULL: deterministic action. */
typedef struct wmrl_act_sampling {
  const float* std_w;          /* stddev.weight [A][Ha] */
  const float* std_b;          /* stddev.bias   [A]     */
  const float* noise_action;   /* [B][A] N(0,1) draws   */
  float* act_mean;             /* [B][A] (row stride ld_out): tanh-squashed mean */
  float* act_std;              /* [B][A] (row stride ld_out) */
  float* entropy;              /* [B] (row stride ld_out, like every output), may be NULL */
} wmrl_act_sampling;
WMRL_API int wmrl_act_step(const wmrl_dims* d, const wmrl_rssm_params* w, const wmrl_actor_params* aw,
                  const float* obs_embed, const float* deter, const float* noise_post, const float* noise_prior,
                  float* action, float* post_stoch, float* post_dist_a, float* post_dist_b, float* deter_next,
                  float* prior_stoch, float* prior_dist_a, float* prior_dist_b, int64_t ld_out,
                  const wmrl_act_sampling* sampling, void* stream);

/*
 * get_features() of a state sequence (state/continuous.py:73-77, state/discrete.py get_features): features[B][T-1][D + S_in] =
 * cat(deter_seq[:, 1:], stoch_seq[:, 1:], -1) as one streaming pass, and its backward (g_deter_seq [B][T][D] / g_stoch_seq
 * [B][T][S_in] are fully written: index 0 is zero).
 */
WMRL_API int wmrl_features_fwd(const float* deter_seq, const float* stoch_seq, int64_t B, int32_t T, int32_t D, int32_t S_in,
                      float* features, void* stream);
WMRL_API int wmrl_features_bwd(const float* g_features, int64_t B, int32_t T, int32_t D, int32_t S_in, float* g_deter_seq,
                      float* g_stoch_seq, void* stream);

/*
 * Reward and done losses of WorldModel.update (world_model.py:130, 143-144) in one pass over the two heads' outputs:
 * sums[0] += sum_i 0.5 (tgt_r - pred_r)^2 + 0.5 log(2 pi)  (unit-variance Normal NLL, get_norm_dist),
 * sums[1] += sum_i binary_cross_entropy_with_logits(logit_d, tgt_d); the caller zero-fills sums and divides by the
 * row counts.  g_pred_r / g_logit_d (either may be NULL) receive the UNSCALED gradients pred_r - tgt_r and
 * sigmoid(logit_d) - tgt_d in the same pass.
 */
WMRL_API int wmrl_head_losses(const float* pred_r, const float* tgt_r, int64_t n_reward, const float* logit_d,
                     const float* tgt_d, int64_t n_done, float* sums, float* g_pred_r, float* g_logit_d, void* stream);

/*
 * lambda-return targets of ValueGradTrainer.value_loss (value_trainer.py:126-134
 * over :62-113) as one O(H) scan per row; V, r, d, out: [B][H] (the trailing
 * singleton of the reference's [B,H,1] dropped).  d is the pre-processed 0/1
 * done mask (value_trainer.py:148-149, see wmrl_done_mask).
 */
WMRL_API int wmrl_lambda_return_fwd(const float* V, const float* r, const float* d, float gamma, float lam,
                           int32_t B, int32_t H, float* out, void* stream);
WMRL_API int wmrl_lambda_return_bwd(const float* g_out, const float* d, float gamma, float lam, int32_t B,
                           int32_t H, float* g_V, float* g_r, void* stream);
/* shift right by one, threshold 0.5, running max (value_trainer.py:148-149). */
WMRL_API int wmrl_done_mask(const float* dones, int32_t B, int32_t H, float* out, void* stream);

/*
 * The Linear primitive every layer of the path is built from (torch.nn.Linear as used at
 * rssm.py:28-42, actor.py:23-40 and by the heads next to the path, models.py:3-37), exposed so the
 * callers either side can run their MLPs on the same kernels and so the GEMM can be tested directly.
 * fp32 in / fp32 out; more than 64 rows run on the tensor cores as 3xTF32 (fp32-accurate), fewer on
 * the cluster split-K SIMT kernel.  W is torch layout [N][K] row-major; ld* are row strides in elements.
 *   fwd:   Y[M,N]  = act(X[M,K] W^T + bias)          act: 0 none, 1 ELU; bias may be NULL
 *   dgrad: dX[M,K] (+)= dY[M,N] W                    accumulate != 0 adds into dX
 *   wgrad: dW[N,K] += dY[M,N]^T X[M,K]               always accumulates (zero dW first)
 */
WMRL_API int wmrl_linear_fwd(int32_t M, int32_t N, int32_t K, const float* X, int64_t ldx, const float* W,
                    const float* bias, float* Y, int64_t ldy, int32_t act, void* stream);
WMRL_API int wmrl_linear_dgrad(int32_t M, int32_t N, int32_t K, const float* dY, int64_t lddy, const float* W,
                      float* dX, int64_t lddx, int32_t accumulate, void* stream);
WMRL_API int wmrl_linear_wgrad(int32_t M, int32_t N, int32_t K, const float* dY, int64_t lddy, const float* X,
                      int64_t ldx, float* dW, void* stream);

/*
 * Fused LayerNorm (eps 1e-5, affine) + activation over rows, the block that follows every hidden Linear
 * of the actor (actor.py:27-36, ELU) and of the heads next to the path (models.py:3-37 SiLU,
 * trainers/value/critic.py:4-29 ReLU).  act: 0 ELU, 1 SiLU, 2 ReLU.  x, xhat, y, dy, dx: [M][W] contiguous;
 * rstd: [M].  fwd stores xhat and rstd for the backward; bwd ADDS into dgamma / dbeta (either may be NULL).
 */
WMRL_API int wmrl_ln_act_fwd(int32_t M, int32_t W, const float* x, const float* gamma, const float* beta, int32_t act,
                    float* xhat, float* rstd, float* y, void* stream);
WMRL_API int wmrl_ln_act_bwd(int32_t M, int32_t W, const float* dy, const float* xhat, const float* rstd,
                    const float* gamma, const float* beta, int32_t act, float* dx, float* dgamma, float* dbeta,
                    void* stream);

/*
 * The MLP heads either side of the path on the bf16 tcgen05 chain: the reward / done models
 * (rssm_world_model/models.py:3-37, [Linear, LayerNorm, SiLU] x depth -> Linear, called on the imagined features at
 * world_model.py:178-180) and the value critic / target critic (trainers/value/critic.py:4-29, ReLU; called at
 * value_trainer.py:121-123).  rows = B * H feature rows; hidden % 32 == 0, hidden <= 512, out_dim <= 16.
 *   wmrl_mlp_to_panel   fp32 x[rows][in_dim] -> the bf16 MMA panel every head over the same features reads
 *   wmrl_mlp_fwd_bf16   out[rows][out_dim] (pre-output-activation); with WMRL_MLP_SAVE records y / x-hat / rstd of
 *                       every block into `saved` for the backward
 *   wmrl_mlp_bwd_bf16   g_x[rows][in_dim] (NULL: skipped; WMRL_MLP_XGRAD_ACCUMULATE adds onto it) and, with
 *                       WMRL_MLP_WGRAD, the parameter gradients (ADDED into gw's tensors: zero them first; NULL
 *                       members are skipped).  dims.flags must carry the same WMRL_MLP_WGRAD bit in
 *                       wmrl_mlp_workspace_bytes(d, 1) and in the call.
 * Weights travel as a packed bf16 image (wmrl_mlp_pack) holding the forward and the transposed backward operands.
 */
#define WMRL_MAX_MLP_LAYERS 8
enum { WMRL_ACT_ELU = 0, WMRL_ACT_SILU = 1, WMRL_ACT_RELU = 2 };
enum { WMRL_MLP_SAVE = 1, WMRL_MLP_WGRAD = 2, WMRL_MLP_XGRAD_ACCUMULATE = 4 };
typedef struct wmrl_mlp_dims {
  int64_t rows;
  int32_t in_dim, hidden, depth, out_dim;
  int32_t act;     /* WMRL_ACT_*  */
  int32_t flags;   /* WMRL_MLP_*  */
} wmrl_mlp_dims;
typedef struct wmrl_mlp_params {
  const float* lin_w[WMRL_MAX_MLP_LAYERS];   /* [hidden][in_dim] (block 0) or [hidden][hidden] */
  const float* lin_b[WMRL_MAX_MLP_LAYERS];
  const float* ln_g[WMRL_MAX_MLP_LAYERS];
  const float* ln_b[WMRL_MAX_MLP_LAYERS];
  const float* out_w;                        /* [out_dim][hidden] */
  const float* out_b;
} wmrl_mlp_params;
typedef struct wmrl_mlp_grads {
  float* lin_w[WMRL_MAX_MLP_LAYERS];
  float* lin_b[WMRL_MAX_MLP_LAYERS];
  float* ln_g[WMRL_MAX_MLP_LAYERS];
  float* ln_b[WMRL_MAX_MLP_LAYERS];
  float* out_w;
  float* out_b;
} wmrl_mlp_grads;
WMRL_API size_t wmrl_mlp_packed_bytes(const wmrl_mlp_dims* d);      /* 0: shape not supported */
WMRL_API size_t wmrl_mlp_panel_bytes(const wmrl_mlp_dims* d);
WMRL_API size_t wmrl_mlp_saved_bytes(const wmrl_mlp_dims* d);
WMRL_API size_t wmrl_mlp_workspace_bytes(const wmrl_mlp_dims* d, int32_t backward);
WMRL_API int wmrl_mlp_pack(const wmrl_mlp_dims* d, const wmrl_mlp_params* w, void* packed, void* stream);
WMRL_API int wmrl_mlp_to_panel(const wmrl_mlp_dims* d, const float* x, void* panel, void* stream);
WMRL_API int wmrl_mlp_fwd_bf16(const wmrl_mlp_dims* d, const void* packed, const void* x_panel, float* out, void* saved,
                      size_t saved_bytes, void* ws, size_t ws_bytes, void* stream);
WMRL_API int wmrl_mlp_bwd_bf16(const wmrl_mlp_dims* d, const void* packed, const void* x_panel, const void* saved,
                      const float* g_out, const wmrl_mlp_grads* gw, float* g_x, void* ws, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WMRL_H_ */
