"""ctypes binding of libwmrl.so (include/wmrl.h).

The library is the product: there is no Python / CPU fallback.  Loading
failures and non-zero return codes raise ``RuntimeError``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwmrl.so")
MAX_ACTOR_LAYERS = 8
ABI_VERSION = 3

CONTINUOUS, DISCRETE = 0, 1
FP32, BF16 = 0, 1
FLAG_WM_WGRAD, FLAG_SAVE_FOR_BWD, FLAG_TRACE, FLAG_BACKWARD = 1, 2, 4, 8
OP_IMAGINE, OP_OBSERVE = 0, 1

_f32p = C.c_void_p  # device pointers travel as integers


class Dims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("variant", "precision", "B", "T", "D", "S", "C", "A", "Hd", "E",
                                          "Ha", "La")] + [("action_scale", C.c_float), ("flags", C.c_int32)]


RSSM_FIELDS = ("embed_w", "embed_b", "w_ih", "w_hh", "b_ih", "b_hh", "prior0_w", "prior0_b", "prior2_w",
               "prior2_b", "post0_w", "post0_b", "post2_w", "post2_b")
# wmrl_rssm_params field -> reference state_dict key (rssm.py:27-42)
RSSM_KEYS = {
    "embed_w": "fc_state_action_embed.weight", "embed_b": "fc_state_action_embed.bias",
    "w_ih": "rnn.weight_ih", "w_hh": "rnn.weight_hh", "b_ih": "rnn.bias_ih", "b_hh": "rnn.bias_hh",
    "prior0_w": "state_prior.0.weight", "prior0_b": "state_prior.0.bias",
    "prior2_w": "state_prior.2.weight", "prior2_b": "state_prior.2.bias",
    "post0_w": "state_posterior.0.weight", "post0_b": "state_posterior.0.bias",
    "post2_w": "state_posterior.2.weight", "post2_b": "state_posterior.2.bias",
}


class RssmPtrs(C.Structure):
    """wmrl_rssm_params and wmrl_rssm_grads share this layout."""
    _fields_ = [(n, _f32p) for n in RSSM_FIELDS]


class ActorPtrs(C.Structure):
    """wmrl_actor_params (with bound) - wmrl_actor_grads is the same minus `bound`."""
    _fields_ = [("lin_w", _f32p * MAX_ACTOR_LAYERS), ("lin_b", _f32p * MAX_ACTOR_LAYERS),
                ("ln_g", _f32p * MAX_ACTOR_LAYERS), ("ln_b", _f32p * MAX_ACTOR_LAYERS),
                ("mu_w", _f32p), ("mu_b", _f32p), ("bound", _f32p)]


class ActorGradPtrs(C.Structure):
    _fields_ = [("lin_w", _f32p * MAX_ACTOR_LAYERS), ("lin_b", _f32p * MAX_ACTOR_LAYERS),
                ("ln_g", _f32p * MAX_ACTOR_LAYERS), ("ln_b", _f32p * MAX_ACTOR_LAYERS),
                ("mu_w", _f32p), ("mu_b", _f32p)]


MAX_MLP_LAYERS = 8
MLP_SAVE, MLP_WGRAD, MLP_XGRAD_ACCUMULATE = 1, 2, 4


class MlpDims(C.Structure):
    """wmrl_mlp_dims"""
    _fields_ = [("rows", C.c_int64)] + [(n, C.c_int32) for n in ("in_dim", "hidden", "depth", "out_dim", "act", "flags")]


class MlpPtrs(C.Structure):
    """wmrl_mlp_params and wmrl_mlp_grads share this layout."""
    _fields_ = [("lin_w", _f32p * MAX_MLP_LAYERS), ("lin_b", _f32p * MAX_MLP_LAYERS),
                ("ln_g", _f32p * MAX_MLP_LAYERS), ("ln_b", _f32p * MAX_MLP_LAYERS),
                ("out_w", _f32p), ("out_b", _f32p)]


class ActSampling(C.Structure):
    """wmrl_act_sampling"""
    _fields_ = [(n, _f32p) for n in ("std_w", "std_b", "noise_action", "act_mean", "act_std", "entropy")]


# every symbol include/wmrl.h declares, with its ctypes signature
_P = C.c_void_p
_SIGNATURES = {
    "wmrl_abi_version": (C.c_int, []),
    "wmrl_last_error": (C.c_char_p, []),
    "wmrl_device_supports_bf16": (C.c_int, []),
    "wmrl_bf16_train_supported": (C.c_int, [C.POINTER(Dims)]),
    "wmrl_saved_bytes": (C.c_size_t, [C.POINTER(Dims), C.c_int]),
    "wmrl_workspace_bytes": (C.c_size_t, [C.POINTER(Dims), C.c_int]),
    "wmrl_packed_bytes": (C.c_size_t, [C.POINTER(Dims)]),
    "wmrl_pack_weights": (C.c_int, [C.POINTER(Dims), C.POINTER(RssmPtrs), C.POINTER(ActorPtrs), _P, _P]),
    "wmrl_imagine_fwd": (C.c_int, [C.POINTER(Dims), C.POINTER(RssmPtrs), C.POINTER(ActorPtrs)] + [_P] * 11 +
                         [C.c_size_t, _P, C.c_size_t, _P]),
    "wmrl_imagine_bwd": (C.c_int, [C.POINTER(Dims), C.POINTER(RssmPtrs), C.POINTER(ActorPtrs)] + [_P] * 12 +
                         [C.POINTER(ActorGradPtrs), C.POINTER(RssmPtrs), _P, _P, _P, C.c_size_t, _P]),
    "wmrl_observe_fwd": (C.c_int, [C.POINTER(Dims), C.POINTER(RssmPtrs)] + [_P] * 13 +
                         [C.c_size_t, _P, C.c_size_t, _P]),
    "wmrl_observe_packed_bytes": (C.c_size_t, [C.POINTER(Dims)]),
    "wmrl_pack_observe_weights": (C.c_int, [C.POINTER(Dims), C.POINTER(RssmPtrs), _P, _P]),
    "wmrl_observe_fwd_bf16": (C.c_int, [C.POINTER(Dims)] + [_P] * 14 + [C.c_size_t, _P, C.c_size_t, _P]),
    "wmrl_observe_bwd_bf16": (C.c_int, [C.POINTER(Dims)] + [_P] * 16 + [C.POINTER(RssmPtrs), _P, _P, _P, C.c_size_t, _P]),
    "wmrl_observe_bwd": (C.c_int, [C.POINTER(Dims), C.POINTER(RssmPtrs)] + [_P] * 21 +
                         [C.POINTER(RssmPtrs), _P, _P, _P, C.c_size_t, _P]),
    "wmrl_kl_loss_fwd": (C.c_int, [C.POINTER(Dims), _P, _P, _P, _P, C.c_float, _P, _P, _P]),
    "wmrl_kl_loss_bwd": (C.c_int, [C.POINTER(Dims), _P, _P, _P, _P, C.c_float, C.c_float, _P, _P, _P, _P, _P]),
    "wmrl_act_step": (C.c_int, [C.POINTER(Dims), C.POINTER(RssmPtrs), C.POINTER(ActorPtrs)] + [_P] * 12 + [C.c_int64, C.POINTER(ActSampling), _P]),
    "wmrl_features_fwd": (C.c_int, [_P, _P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "wmrl_features_bwd": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P]),
    "wmrl_head_losses": (C.c_int, [_P, _P, C.c_int64, _P, _P, C.c_int64, _P, _P, _P, _P]),
    "wmrl_lambda_return_fwd": (C.c_int, [_P, _P, _P, C.c_float, C.c_float, C.c_int32, C.c_int32, _P, _P]),
    "wmrl_lambda_return_bwd": (C.c_int, [_P, _P, C.c_float, C.c_float, C.c_int32, C.c_int32, _P, _P, _P]),
    "wmrl_done_mask": (C.c_int, [_P, C.c_int32, C.c_int32, _P, _P]),
    "wmrl_linear_fwd": (C.c_int, [C.c_int32] * 3 + [_P, C.c_int64, _P, _P, _P, C.c_int64, C.c_int32, _P]),
    "wmrl_linear_dgrad": (C.c_int, [C.c_int32] * 3 + [_P, C.c_int64, _P, _P, C.c_int64, C.c_int32, _P]),
    "wmrl_linear_wgrad": (C.c_int, [C.c_int32] * 3 + [_P, C.c_int64, _P, C.c_int64, _P, _P]),
    "wmrl_ln_act_fwd": (C.c_int, [C.c_int32, C.c_int32, _P, _P, _P, C.c_int32, _P, _P, _P, _P]),
    "wmrl_ln_act_bwd": (C.c_int, [C.c_int32, C.c_int32, _P, _P, _P, _P, _P, C.c_int32, _P, _P, _P, _P]),
    "wmrl_mlp_packed_bytes": (C.c_size_t, [C.POINTER(MlpDims)]),
    "wmrl_mlp_panel_bytes": (C.c_size_t, [C.POINTER(MlpDims)]),
    "wmrl_mlp_saved_bytes": (C.c_size_t, [C.POINTER(MlpDims)]),
    "wmrl_mlp_workspace_bytes": (C.c_size_t, [C.POINTER(MlpDims), C.c_int32]),
    "wmrl_mlp_pack": (C.c_int, [C.POINTER(MlpDims), C.POINTER(MlpPtrs), _P, _P]),
    "wmrl_mlp_to_panel": (C.c_int, [C.POINTER(MlpDims), _P, _P, _P]),
    "wmrl_mlp_fwd_bf16": (C.c_int, [C.POINTER(MlpDims), _P, _P, _P, _P, C.c_size_t, _P, C.c_size_t, _P]),
    "wmrl_mlp_bwd_bf16": (C.c_int, [C.POINTER(MlpDims), _P, _P, _P, _P, C.POINTER(MlpPtrs), _P, _P, C.c_size_t, _P]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load libwmrl.so (built in-tree by build.py).  Raises if it is missing -
    there is deliberately no fallback implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"wmrl: native library {LIB_PATH} is missing. Build it with "
            "`python world-model-rl_b200/build.py` (needs nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    ver = lib.wmrl_abi_version()
    if ver != ABI_VERSION:
        raise RuntimeError(f"wmrl: libwmrl.so ABI {ver} != binding ABI {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().wmrl_last_error()
        raise RuntimeError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")


def ptr(t) -> Optional[int]:
    """device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()
