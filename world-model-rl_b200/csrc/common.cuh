// Shared helpers for libwmrl (sm_100a).  Internal - the public surface is include/wmrl.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "wmrl.h"

namespace wmrl {

void set_error(const std::string& msg);

#define WMRL_CUDA_OK(expr)                                                                  \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      ::wmrl::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " at " +        \
                        __FILE__ + ":" + std::to_string(__LINE__));                         \
      return 1;                                                                             \
    }                                                                                       \
  } while (0)

#define WMRL_REQUIRE(cond, msg)                                  \
  do {                                                           \
    if (!(cond)) {                                               \
      ::wmrl::set_error(std::string("wmrl: ") + (msg));          \
      return 2;                                                  \
    }                                                            \
  } while (0)

// Derived sizes of a wmrl_dims.
struct Dims {
  int variant, B, T, D, S, C, A, Hd, E, Ha, La, flags;
  int S_in, P, F;
  float action_scale;
  explicit Dims(const wmrl_dims& d)
      : variant(d.variant), B(d.B), T(d.T), D(d.D), S(d.S), C(d.variant == WMRL_DISCRETE ? d.C : 1),
        A(d.A), Hd(d.Hd), E(d.E), Ha(d.Ha), La(d.La), flags(d.flags), action_scale(d.action_scale) {
    S_in = S * C;
    P = (variant == WMRL_DISCRETE) ? S * C : 2 * S;
    F = D + S_in;
  }
};

int check_dims(const wmrl_dims* d, int op);

// fp32 SIMT path (f32_path.cu)
size_t f32_saved_bytes(const Dims& d, int op);
size_t f32_workspace_bytes(const Dims& d, int op);
int f32_imagine_fwd(const Dims& d, const wmrl_rssm_params* w, const wmrl_actor_params* aw,
                    const float* deter0, const float* stoch0, const float* noise, float* deter_seq,
                    float* stoch_seq, float* dist_a, float* dist_b, float* actions, int32_t* idx,
                    float* saved, float* ws, cudaStream_t st);
int f32_imagine_bwd(const Dims& d, const wmrl_rssm_params* w, const wmrl_actor_params* aw,
                    const float* noise, const float* deter_seq, const float* stoch_seq,
                    const float* dist_a, const float* dist_b, const float* actions,
                    const float* saved, const float* g_deter_seq, const float* g_stoch_seq,
                    const float* g_dist_a, const float* g_dist_b, const wmrl_actor_grads* gaw,
                    const wmrl_rssm_grads* gw, float* g_deter0, float* g_stoch0, float* ws,
                    cudaStream_t st);
int f32_observe_fwd(const Dims& d, const wmrl_rssm_params* w, const float* obs, const float* act,
                    const float* noise_pri, const float* noise_post, float* pri_deter,
                    float* pri_stoch, float* pri_da, float* pri_db, float* post_deter,
                    float* post_stoch, float* post_da, float* post_db, float* saved, float* ws,
                    cudaStream_t st);
int f32_observe_bwd(const Dims& d, const wmrl_rssm_params* w, const float* obs, const float* act,
                    const float* noise_pri, const float* noise_post, const float* pri_deter,
                    const float* pri_stoch, const float* pri_da, const float* pri_db,
                    const float* post_deter, const float* post_stoch, const float* post_da,
                    const float* post_db, const float* saved, const float* g_pri_deter,
                    const float* g_pri_stoch, const float* g_pri_da, const float* g_pri_db,
                    const float* g_post_deter, const float* g_post_stoch, const float* g_post_da,
                    const float* g_post_db, const wmrl_rssm_grads* gw, float* g_obs, float* g_act,
                    float* ws, cudaStream_t st);

// bf16 tcgen05 path (tc_imagine.cu)
bool tc_supported(const Dims& d);
size_t tc_packed_bytes(const Dims& d);
size_t tc_workspace_bytes(const Dims& d);
int tc_pack_weights(const Dims& d, const wmrl_rssm_params* w, const wmrl_actor_params* aw,
                    void* packed, cudaStream_t st);
int tc_imagine_fwd(const Dims& d, const wmrl_rssm_params* w, const wmrl_actor_params* aw,
                   const void* packed, const float* deter0, const float* stoch0,
                   const float* noise, float* deter_seq, float* stoch_seq, float* dist_a,
                   float* dist_b, float* actions, int32_t* idx, void* saved, void* ws, cudaStream_t st);
// bf16 BPTT (tc_bptt.cu): persistent reverse-time kernel + weight-gradient GEMMs over the activation stash
bool tc_train_supported(const Dims& d);
size_t tc_saved_bytes(const Dims& d);
size_t tc_bwd_workspace_bytes(const Dims& d);
int tc_imagine_bwd(const Dims& d, const void* packed, const float* noise, const float* deter_seq, const float* dist_b,
                   const float* actions, const void* saved, const float* g_deter_seq, const float* g_stoch_seq,
                   const float* g_dist_a, const float* g_dist_b, const wmrl_actor_grads* gaw, float* g_deter0,
                   float* g_stoch0, void* ws, cudaStream_t st);

// layer-wise bf16 tcgen05 chain (tc_layer.cu): the discrete RSSM and the continuous shapes outside the persistent kernel
bool tcl_supported(const Dims& d);
size_t tcl_packed_bytes(const Dims& d);
size_t tcl_workspace_bytes(const Dims& d);
int tcl_pack_weights(const Dims& d, const wmrl_rssm_params* w, const wmrl_actor_params* aw, void* packed,
                     cudaStream_t st);
int tcl_imagine_fwd(const Dims& d, const void* packed, const float* deter0, const float* stoch0, const float* noise,
                    float* deter_seq, float* stoch_seq, float* dist_a, float* dist_b, float* actions, int32_t* idx,
                    void* saved, void* ws, cudaStream_t st);
bool tcl_train_supported(const Dims& d);
size_t tcl_saved_bytes(const Dims& d);
int tcl_imagine_bwd(const Dims& d, const void* packed, const float* noise, const float* deter_seq, const float* dist_a,
                    const float* dist_b, const float* actions, const void* saved, const float* g_deter_seq,
                    const float* g_stoch_seq, const float* g_dist_a, const float* g_dist_b, const wmrl_actor_grads* gaw,
                    float* g_deter0, float* g_stoch0, void* ws, cudaStream_t st);
// posterior filtering (observe_rollout forward) on the layer-wise chain
bool tcl_observe_supported(const Dims& d);
size_t tcl_observe_packed_bytes(const Dims& d);
size_t tcl_observe_workspace_bytes(const Dims& d);
int tcl_pack_observe_weights(const Dims& d, const wmrl_rssm_params* w, void* packed, cudaStream_t st);
int tcl_observe_fwd(const Dims& d, const void* packed, const float* obs, const float* act, const float* noise_pri,
                    const float* noise_post, float* pri_deter, float* pri_stoch, float* pri_da, float* pri_db,
                    float* post_deter, float* post_stoch, float* post_da, float* post_db, void* saved, void* ws,
                    cudaStream_t st);
size_t tcl_observe_saved_bytes(const Dims& d);
int tcl_observe_bwd(const Dims& d, const void* packed, const float* noise_pri, const float* noise_post, const float* pri_deter,
                    const float* pri_da, const float* pri_db, const float* post_da, const float* post_db, const void* saved,
                    const float* g_deter, const float* g_pri_stoch, const float* g_pri_da, const float* g_pri_db,
                    const float* g_post_stoch, const float* g_post_da, const float* g_post_db, const wmrl_rssm_grads* gw,
                    float* g_obs, float* g_act, void* ws, cudaStream_t st);
// MLP heads (reward / done models, critic) on the layer-wise chain
bool tcm_supported(const wmrl_mlp_dims* d);
size_t tcm_packed_bytes(const wmrl_mlp_dims* d);
size_t tcm_panel_bytes(const wmrl_mlp_dims* d);
size_t tcm_saved_bytes(const wmrl_mlp_dims* d);
size_t tcm_workspace_bytes(const wmrl_mlp_dims* d, int backward);
int tcm_pack(const wmrl_mlp_dims* d, const wmrl_mlp_params* w, void* packed, cudaStream_t st);
int tcm_to_panel(const wmrl_mlp_dims* d, const float* x, void* panel, cudaStream_t st);
int tcm_fwd(const wmrl_mlp_dims* d, const void* packed, const void* x_panel, float* out, void* saved, void* ws,
            cudaStream_t st);
int tcm_bwd(const wmrl_mlp_dims* d, const void* packed, const void* x_panel, const void* saved, const float* g_out,
            const wmrl_mlp_grads* gw, float* g_x, void* ws, cudaStream_t st);
// which bf16 forward serves this shape: 1 persistent folded-pair / single-CTA kernel, 2 layer-wise chain, 0 none
int bf16_path(const Dims& d);

}  // namespace wmrl
