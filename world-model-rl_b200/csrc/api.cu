// C ABI of libwmrl (include/wmrl.h): argument checking and dispatch between the
// fp32 SIMT path (f32_path.cu) and the bf16 tcgen05 path (tc_imagine.cu).
#include <stdlib.h>

#include <mutex>

#include "common.cuh"

namespace wmrl {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }

int check_dims(const wmrl_dims* d, int op) {
  WMRL_REQUIRE(d != nullptr, "dims is NULL");
  WMRL_REQUIRE(d->variant == WMRL_CONTINUOUS || d->variant == WMRL_DISCRETE, "bad variant");
  WMRL_REQUIRE(d->precision == WMRL_FP32 || d->precision == WMRL_BF16, "bad precision");
  WMRL_REQUIRE(d->B >= 0 && d->T >= 0, "negative B/T");
  WMRL_REQUIRE(d->D > 0 && d->S > 0 && d->A > 0 && d->Hd > 0, "D,S,A,Hd must be positive");
  if (d->variant == WMRL_DISCRETE) WMRL_REQUIRE(d->C > 0, "discrete variant needs C > 0");
  if (op == 0) {
    WMRL_REQUIRE(d->Ha > 0 && d->La > 0 && d->La <= WMRL_MAX_ACTOR_LAYERS, "bad actor shape");
    WMRL_REQUIRE(d->action_scale != 0.f, "action_scale must be non-zero");
  } else {
    WMRL_REQUIRE(d->E > 0, "observe needs E > 0");
  }
  return 0;
}

// Which bf16 forward serves a shape.  The persistent kernels (tc_imagine.cu) are preferred; the layer-wise
// chain (tc_layer.cu) covers the discrete variant and everything else.  WMRL_BF16_PATH=layer forces the chain
// (A/B comparisons, tests).
int bf16_path(const Dims& d) {
  static int force_layer = -1;
  if (force_layer < 0) {
    const char* e = getenv("WMRL_BF16_PATH");
    force_layer = (e && e[0] == 'l') ? 1 : 0;
  }
  if (!force_layer && tc_supported(d)) return 1;
  if (tcl_supported(d)) return 2;
  return (force_layer && tc_supported(d)) ? 1 : 0;
}

static int require_device() {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error(std::string("wmrl: no CUDA device: ") + cudaGetErrorString(e));
    return 3;
  }
  return 0;
}

}  // namespace wmrl

using namespace wmrl;

extern "C" {

int wmrl_abi_version(void) { return WMRL_ABI_VERSION; }
const char* wmrl_last_error(void) { return g_last_error.c_str(); }

int wmrl_device_supports_bf16(void) {
  // cudaGetDeviceProperties costs milliseconds; the attribute query is cheap and cached per device
  static int cached[64];
  static bool have[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (dev >= 0 && dev < 64 && have[dev]) return cached[dev];
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  const int ok = major == 10 ? 1 : 0;
  if (dev >= 0 && dev < 64) { cached[dev] = ok; have[dev] = true; }
  return ok;
}

int wmrl_bf16_train_supported(const wmrl_dims* d) {
  if (check_dims(d, 0)) return 0;
  Dims dd(*d);
  const int path = bf16_path(dd);
  return ((path == 1 && tc_train_supported(dd)) || (path == 2 && tcl_train_supported(dd))) ? 1 : 0;
}

size_t wmrl_saved_bytes(const wmrl_dims* d, int op) {
  if (check_dims(d, op)) return 0;
  Dims dd(*d);
  if (op == 0 && d->precision == WMRL_BF16) {
    const int path = bf16_path(dd);
    if (path == 1 && tc_train_supported(dd)) return tc_saved_bytes(dd);
    if (path == 2 && tcl_train_supported(dd)) return tcl_saved_bytes(dd);
  }
  if (op == 1 && d->precision == WMRL_BF16 && tcl_observe_supported(dd)) return tcl_observe_saved_bytes(dd);
  return f32_saved_bytes(dd, op);
}

size_t wmrl_workspace_bytes(const wmrl_dims* d, int op) {
  if (check_dims(d, op)) return 0;
  Dims dd(*d);
  size_t a = f32_workspace_bytes(dd, op);
  if (d->precision == WMRL_BF16 && op == 0) {
    const int path = bf16_path(dd);
    size_t b = path == 2 ? tcl_workspace_bytes(dd) : tc_workspace_bytes(dd);
    if (b > a) a = b;
    if ((d->flags & WMRL_FLAG_BACKWARD) && path == 1 && tc_train_supported(dd)) {   // gradient stash of the bf16 BPTT
      b = tc_bwd_workspace_bytes(dd);
      if (b > a) a = b;
    }
  }
  if (d->precision == WMRL_BF16 && op == 1) {
    const size_t b = tcl_observe_workspace_bytes(dd);
    if (b > a) a = b;
  }
  return a;
}

size_t wmrl_observe_packed_bytes(const wmrl_dims* d) {
  if (check_dims(d, 1)) return 0;
  Dims dd(*d);
  return tcl_observe_packed_bytes(dd);
}

int wmrl_pack_observe_weights(const wmrl_dims* d, const wmrl_rssm_params* w, void* packed, void* stream) {
  if (int r = check_dims(d, 1)) return r;
  WMRL_REQUIRE(w && packed, "pack_observe_weights: NULL argument");
  Dims dd(*d);
  WMRL_REQUIRE(tcl_observe_supported(dd), "pack_observe_weights: shape not supported by the tcgen05 path");
  if (int r = require_device()) return r;
  return tcl_pack_observe_weights(dd, w, packed, (cudaStream_t)stream);
}

int wmrl_observe_fwd_bf16(const wmrl_dims* d, const void* packed, const float* obs_embeds, const float* actions,
                          const float* noise_prior, const float* noise_post, float* pri_deter, float* pri_stoch,
                          float* pri_dist_a, float* pri_dist_b, float* post_deter, float* post_stoch, float* post_dist_a,
                          float* post_dist_b, void* saved, size_t saved_bytes, void* ws, size_t ws_bytes, void* stream) {
  if (int r = check_dims(d, 1)) return r;
  Dims dd(*d);
  if (dd.B == 0 || dd.T == 0) return 0;
  WMRL_REQUIRE(d->precision == WMRL_BF16, "observe_fwd_bf16: dims.precision must be WMRL_BF16");
  WMRL_REQUIRE(packed, "observe_fwd_bf16: needs the packed weight image (wmrl_pack_observe_weights)");
  WMRL_REQUIRE(tcl_observe_supported(dd), "observe_fwd_bf16: shape not supported by the tcgen05 path");
  WMRL_REQUIRE(pri_deter && pri_stoch && pri_dist_a && post_deter && post_stoch && post_dist_a, "observe_fwd_bf16: NULL output");
  WMRL_REQUIRE(dd.T == 1 || (obs_embeds && actions && noise_prior && noise_post), "observe_fwd_bf16: NULL input");
  WMRL_REQUIRE(dd.variant == WMRL_DISCRETE || (pri_dist_b && post_dist_b), "observe_fwd_bf16: continuous needs stds");
  WMRL_REQUIRE(ws && ws_bytes >= tcl_observe_workspace_bytes(dd), "observe_fwd_bf16: workspace too small");
  const bool save = (d->flags & WMRL_FLAG_SAVE_FOR_BWD) != 0;
  if (save) WMRL_REQUIRE(saved && saved_bytes >= tcl_observe_saved_bytes(dd), "observe_fwd_bf16: saved buffer too small");
  WMRL_REQUIRE(wmrl_device_supports_bf16(), "observe_fwd_bf16: needs an sm_100 device");
  if (int r = require_device()) return r;
  return tcl_observe_fwd(dd, packed, obs_embeds, actions, noise_prior, noise_post, pri_deter, pri_stoch, pri_dist_a,
                         pri_dist_b, post_deter, post_stoch, post_dist_a, post_dist_b, save ? saved : nullptr, ws,
                         (cudaStream_t)stream);
}

int wmrl_observe_bwd_bf16(const wmrl_dims* d, const void* packed, const float* noise_prior, const float* noise_post,
                          const float* pri_deter, const float* pri_dist_a, const float* pri_dist_b, const float* post_dist_a,
                          const float* post_dist_b, const void* saved, const float* g_deter, const float* g_pri_stoch,
                          const float* g_pri_dist_a, const float* g_pri_dist_b, const float* g_post_stoch,
                          const float* g_post_dist_a, const float* g_post_dist_b, const wmrl_rssm_grads* gw,
                          float* g_obs_embeds, float* g_actions, void* ws, size_t ws_bytes, void* stream) {
  if (int r = check_dims(d, 1)) return r;
  Dims dd(*d);
  if (dd.B == 0 || dd.T <= 1) return 0;
  WMRL_REQUIRE(d->precision == WMRL_BF16 && (d->flags & WMRL_FLAG_BACKWARD), "observe_bwd_bf16: dims need WMRL_BF16 and WMRL_FLAG_BACKWARD");
  WMRL_REQUIRE(packed && saved && gw, "observe_bwd_bf16: NULL packed / saved / grads");
  WMRL_REQUIRE(tcl_observe_supported(dd), "observe_bwd_bf16: shape not supported by the tcgen05 path");
  WMRL_REQUIRE(noise_prior && noise_post && pri_deter && pri_dist_a && post_dist_a, "observe_bwd_bf16: NULL tensor");
  WMRL_REQUIRE(dd.variant == WMRL_DISCRETE || (pri_dist_b && post_dist_b), "observe_bwd_bf16: continuous needs stds");
  WMRL_REQUIRE(ws && ws_bytes >= tcl_observe_workspace_bytes(dd), "observe_bwd_bf16: workspace too small");
  WMRL_REQUIRE(wmrl_device_supports_bf16(), "observe_bwd_bf16: needs an sm_100 device");
  if (int r = require_device()) return r;
  return tcl_observe_bwd(dd, packed, noise_prior, noise_post, pri_deter, pri_dist_a, pri_dist_b, post_dist_a, post_dist_b,
                         saved, g_deter, g_pri_stoch, g_pri_dist_a, g_pri_dist_b, g_post_stoch, g_post_dist_a, g_post_dist_b,
                         gw, g_obs_embeds, g_actions, ws, (cudaStream_t)stream);
}

size_t wmrl_packed_bytes(const wmrl_dims* d) {
  if (check_dims(d, 0)) return 0;
  Dims dd(*d);
  const int path = bf16_path(dd);
  return path == 1 ? tc_packed_bytes(dd) : (path == 2 ? tcl_packed_bytes(dd) : 0);
}

int wmrl_pack_weights(const wmrl_dims* d, const wmrl_rssm_params* w, const wmrl_actor_params* aw,
                      void* packed, void* stream) {
  if (int r = check_dims(d, 0)) return r;
  WMRL_REQUIRE(w && aw && packed, "pack_weights: NULL argument");
  Dims dd(*d);
  const int path = bf16_path(dd);
  WMRL_REQUIRE(path != 0, "pack_weights: shape not supported by the tcgen05 path");
  if (int r = require_device()) return r;
  if (path == 2) return tcl_pack_weights(dd, w, aw, packed, (cudaStream_t)stream);
  return tc_pack_weights(dd, w, aw, packed, (cudaStream_t)stream);
}

int wmrl_imagine_fwd(const wmrl_dims* d, const wmrl_rssm_params* w, const wmrl_actor_params* aw,
                     const void* packed, const float* deter0, const float* stoch0,
                     const float* noise, float* deter_seq, float* stoch_seq, float* dist_a,
                     float* dist_b, float* actions, int32_t* idx, void* saved, size_t saved_bytes,
                     void* ws, size_t ws_bytes, void* stream) {
  if (int r = check_dims(d, 0)) return r;
  WMRL_REQUIRE(w && aw, "imagine_fwd: NULL weights");
  Dims dd(*d);
  if (dd.B == 0 || dd.T == 0) return 0;
  WMRL_REQUIRE(deter0 && stoch0 && noise && deter_seq && stoch_seq && dist_a && actions,
               "imagine_fwd: NULL tensor");
  WMRL_REQUIRE(dd.variant == WMRL_DISCRETE || dist_b, "imagine_fwd: continuous needs dist_b (stds)");
  WMRL_REQUIRE(ws && ws_bytes >= wmrl_workspace_bytes(d, 0), "imagine_fwd: workspace too small");
  if (int r = require_device()) return r;
  const bool save = (d->flags & WMRL_FLAG_SAVE_FOR_BWD) != 0;
  if (save)
    WMRL_REQUIRE(saved && saved_bytes >= wmrl_saved_bytes(d, 0), "imagine_fwd: saved buffer too small");
  if (d->precision == WMRL_BF16) {
    const int path = bf16_path(dd);
    WMRL_REQUIRE(path != 0, "imagine_fwd: shape not supported by the bf16 tcgen05 path");
    WMRL_REQUIRE(packed, "imagine_fwd: bf16 path needs the packed weight image");
    WMRL_REQUIRE(wmrl_device_supports_bf16(), "imagine_fwd: bf16 path needs an sm_100 device");
    if (path == 2) {
      WMRL_REQUIRE(dd.variant == WMRL_CONTINUOUS || idx, "imagine_fwd: discrete needs idx");
      return tcl_imagine_fwd(dd, packed, deter0, stoch0, noise, deter_seq, stoch_seq, dist_a, dist_b, actions, idx,
                             save ? saved : nullptr, ws, (cudaStream_t)stream);
    }
    WMRL_REQUIRE(!save || tc_train_supported(dd), "imagine_fwd: this shape has no bf16 training path (activation "
                                                  "stash + BPTT kernel); record with the fp32 forward");
    return tc_imagine_fwd(dd, w, aw, packed, deter0, stoch0, noise, deter_seq, stoch_seq, dist_a,
                          dist_b, actions, idx, save ? saved : nullptr, ws, (cudaStream_t)stream);
  }
  return f32_imagine_fwd(dd, w, aw, deter0, stoch0, noise, deter_seq, stoch_seq, dist_a, dist_b,
                         actions, idx, save ? (float*)saved : nullptr, (float*)ws,
                         (cudaStream_t)stream);
}

int wmrl_imagine_bwd(const wmrl_dims* d, const wmrl_rssm_params* w, const wmrl_actor_params* aw,
                     const void* packed, const float* noise, const float* deter_seq,
                     const float* stoch_seq, const float* dist_a, const float* dist_b,
                     const float* actions, const void* saved, const float* g_deter_seq,
                     const float* g_stoch_seq, const float* g_dist_a, const float* g_dist_b,
                     const wmrl_actor_grads* gaw, const wmrl_rssm_grads* gw, float* g_deter0,
                     float* g_stoch0, void* ws, size_t ws_bytes, void* stream) {
  if (int r = check_dims(d, 0)) return r;
  WMRL_REQUIRE(w && aw && gaw, "imagine_bwd: NULL weights / grads");
  Dims dd(*d);
  if (dd.B == 0 || dd.T == 0) return 0;
  WMRL_REQUIRE(noise && deter_seq && stoch_seq && dist_a && actions && saved, "imagine_bwd: NULL tensor");
  WMRL_REQUIRE(!(d->flags & WMRL_FLAG_WM_WGRAD) || gw, "imagine_bwd: WM_WGRAD needs gw");
  WMRL_REQUIRE(ws && ws_bytes >= wmrl_workspace_bytes(d, 0), "imagine_bwd: workspace too small");
  if (int r = require_device()) return r;
  if (d->precision == WMRL_BF16) {
    // `saved` is the bf16 activation stash written by wmrl_imagine_fwd(precision BF16, SAVE_FOR_BWD)
    const int path = bf16_path(dd);
    WMRL_REQUIRE((path == 1 && tc_train_supported(dd)) || (path == 2 && tcl_train_supported(dd)),
                 "imagine_bwd: this shape has no bf16 BPTT kernel");
    WMRL_REQUIRE(packed, "imagine_bwd: bf16 path needs the packed weight image");
    WMRL_REQUIRE(!(d->flags & WMRL_FLAG_WM_WGRAD), "imagine_bwd: the bf16 BPTT produces actor gradients only "
                                                   "(world model frozen, world_model.py:172); use the fp32 path");
    WMRL_REQUIRE(g_deter0 && g_stoch0 && (dd.variant == WMRL_DISCRETE || dist_b), "imagine_bwd: NULL tensor");
    WMRL_REQUIRE(wmrl_device_supports_bf16(), "imagine_bwd: bf16 path needs an sm_100 device");
    if (path == 2) {
      return tcl_imagine_bwd(dd, packed, noise, deter_seq, dist_a, dist_b, actions, saved, g_deter_seq, g_stoch_seq,
                             g_dist_a, g_dist_b, gaw, g_deter0, g_stoch0, ws, (cudaStream_t)stream);
    }
    return tc_imagine_bwd(dd, packed, noise, deter_seq, dist_b, actions, saved, g_deter_seq, g_stoch_seq, g_dist_a,
                          g_dist_b, gaw, g_deter0, g_stoch0, ws, (cudaStream_t)stream);
  }
  return f32_imagine_bwd(dd, w, aw, noise, deter_seq, stoch_seq, dist_a, dist_b, actions,
                         (const float*)saved, g_deter_seq, g_stoch_seq, g_dist_a, g_dist_b, gaw, gw,
                         g_deter0, g_stoch0, (float*)ws, (cudaStream_t)stream);
}

int wmrl_observe_fwd(const wmrl_dims* d, const wmrl_rssm_params* w, const float* obs_embeds,
                     const float* actions, const float* noise_prior, const float* noise_post,
                     float* pri_deter, float* pri_stoch, float* pri_dist_a, float* pri_dist_b,
                     float* post_deter, float* post_stoch, float* post_dist_a, float* post_dist_b,
                     void* saved, size_t saved_bytes, void* ws, size_t ws_bytes, void* stream) {
  if (int r = check_dims(d, 1)) return r;
  WMRL_REQUIRE(w, "observe_fwd: NULL weights");
  Dims dd(*d);
  if (dd.B == 0 || dd.T == 0) return 0;
  WMRL_REQUIRE(pri_deter && pri_stoch && pri_dist_a && post_deter && post_stoch && post_dist_a,
               "observe_fwd: NULL output");
  WMRL_REQUIRE(dd.T == 1 || (obs_embeds && actions && noise_prior && noise_post), "observe_fwd: NULL input");
  WMRL_REQUIRE(dd.variant == WMRL_DISCRETE || (pri_dist_b && post_dist_b), "observe_fwd: continuous needs stds");
  WMRL_REQUIRE(ws && ws_bytes >= wmrl_workspace_bytes(d, 1), "observe_fwd: workspace too small");
  const bool save = (d->flags & WMRL_FLAG_SAVE_FOR_BWD) != 0;
  if (save && dd.T > 1)
    WMRL_REQUIRE(saved && saved_bytes >= f32_saved_bytes(dd, 1), "observe_fwd: saved buffer too small");
  if (int r = require_device()) return r;
  return f32_observe_fwd(dd, w, obs_embeds, actions, noise_prior, noise_post, pri_deter, pri_stoch,
                         pri_dist_a, pri_dist_b, post_deter, post_stoch, post_dist_a, post_dist_b,
                         save ? (float*)saved : nullptr, (float*)ws, (cudaStream_t)stream);
}

int wmrl_observe_bwd(const wmrl_dims* d, const wmrl_rssm_params* w, const float* obs_embeds,
                     const float* actions, const float* noise_prior, const float* noise_post,
                     const float* pri_deter, const float* pri_stoch, const float* pri_dist_a,
                     const float* pri_dist_b, const float* post_deter, const float* post_stoch,
                     const float* post_dist_a, const float* post_dist_b, const void* saved,
                     const float* g_pri_deter, const float* g_pri_stoch, const float* g_pri_dist_a,
                     const float* g_pri_dist_b, const float* g_post_deter, const float* g_post_stoch,
                     const float* g_post_dist_a, const float* g_post_dist_b,
                     const wmrl_rssm_grads* gw, float* g_obs_embeds, float* g_actions, void* ws,
                     size_t ws_bytes, void* stream) {
  if (int r = check_dims(d, 1)) return r;
  WMRL_REQUIRE(w && gw, "observe_bwd: NULL weights / grads");
  Dims dd(*d);
  if (dd.B == 0 || dd.T <= 1) return 0;
  WMRL_REQUIRE(obs_embeds && actions && noise_prior && noise_post && pri_deter && pri_dist_a &&
                   post_deter && post_stoch && post_dist_a && saved,
               "observe_bwd: NULL tensor");
  WMRL_REQUIRE(ws && ws_bytes >= wmrl_workspace_bytes(d, 1), "observe_bwd: workspace too small");
  if (int r = require_device()) return r;
  return f32_observe_bwd(dd, w, obs_embeds, actions, noise_prior, noise_post, pri_deter, pri_stoch,
                         pri_dist_a, pri_dist_b, post_deter, post_stoch, post_dist_a, post_dist_b,
                         (const float*)saved, g_pri_deter, g_pri_stoch, g_pri_dist_a, g_pri_dist_b,
                         g_post_deter, g_post_stoch, g_post_dist_a, g_post_dist_b, gw, g_obs_embeds,
                         g_actions, (float*)ws, (cudaStream_t)stream);
}

size_t wmrl_mlp_packed_bytes(const wmrl_mlp_dims* d) { return tcm_packed_bytes(d); }
size_t wmrl_mlp_panel_bytes(const wmrl_mlp_dims* d) { return tcm_panel_bytes(d); }
size_t wmrl_mlp_saved_bytes(const wmrl_mlp_dims* d) { return tcm_saved_bytes(d); }
size_t wmrl_mlp_workspace_bytes(const wmrl_mlp_dims* d, int32_t backward) { return tcm_workspace_bytes(d, backward); }

int wmrl_mlp_pack(const wmrl_mlp_dims* d, const wmrl_mlp_params* w, void* packed, void* stream) {
  WMRL_REQUIRE(d && w && packed, "mlp_pack: NULL argument");
  WMRL_REQUIRE(tcm_supported(d), "mlp_pack: shape not supported (hidden % 32 == 0, hidden <= 512, out_dim <= 16)");
  for (int l = 0; l < d->depth; ++l)
    WMRL_REQUIRE(w->lin_w[l] && w->lin_b[l] && w->ln_g[l] && w->ln_b[l], "mlp_pack: NULL block parameter");
  WMRL_REQUIRE(w->out_w && w->out_b, "mlp_pack: NULL output Linear");
  if (int r = require_device()) return r;
  return tcm_pack(d, w, packed, (cudaStream_t)stream);
}

int wmrl_mlp_to_panel(const wmrl_mlp_dims* d, const float* x, void* panel, void* stream) {
  WMRL_REQUIRE(d && tcm_supported(d), "mlp_to_panel: shape not supported");
  if (d->rows == 0) return 0;
  WMRL_REQUIRE(x && panel, "mlp_to_panel: NULL tensor");
  if (int r = require_device()) return r;
  return tcm_to_panel(d, x, panel, (cudaStream_t)stream);
}

int wmrl_mlp_fwd_bf16(const wmrl_mlp_dims* d, const void* packed, const void* x_panel, float* out, void* saved,
                      size_t saved_bytes, void* ws, size_t ws_bytes, void* stream) {
  WMRL_REQUIRE(d && tcm_supported(d), "mlp_fwd: shape not supported");
  if (d->rows == 0) return 0;
  WMRL_REQUIRE(packed && x_panel && out, "mlp_fwd: NULL tensor");
  WMRL_REQUIRE(wmrl_device_supports_bf16(), "mlp_fwd: needs an sm_100 device");
  const bool save = (d->flags & WMRL_MLP_SAVE) != 0;
  WMRL_REQUIRE(!save || (saved && saved_bytes >= tcm_saved_bytes(d)), "mlp_fwd: saved buffer too small");
  WMRL_REQUIRE(ws && ws_bytes >= tcm_workspace_bytes(d, 0), "mlp_fwd: workspace too small");
  if (int r = require_device()) return r;
  return tcm_fwd(d, packed, x_panel, out, save ? saved : nullptr, ws, (cudaStream_t)stream);
}

int wmrl_mlp_bwd_bf16(const wmrl_mlp_dims* d, const void* packed, const void* x_panel, const void* saved,
                      const float* g_out, const wmrl_mlp_grads* gw, float* g_x, void* ws, size_t ws_bytes, void* stream) {
  WMRL_REQUIRE(d && tcm_supported(d), "mlp_bwd: shape not supported");
  if (d->rows == 0) return 0;
  WMRL_REQUIRE(packed && x_panel && saved && g_out, "mlp_bwd: NULL tensor");
  WMRL_REQUIRE(wmrl_device_supports_bf16(), "mlp_bwd: needs an sm_100 device");
  WMRL_REQUIRE(ws && ws_bytes >= tcm_workspace_bytes(d, 1), "mlp_bwd: workspace too small");
  if (int r = require_device()) return r;
  return tcm_bwd(d, packed, x_panel, saved, g_out, gw, g_x, ws, (cudaStream_t)stream);
}

}  // extern "C"
