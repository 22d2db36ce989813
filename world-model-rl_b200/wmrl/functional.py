"""torch.autograd bindings of the libwmrl C ABI.

Three custom Functions - imagine rollout (forward + hand-written actor BPTT),
observe rollout (forward + full backward) and the lambda-return scan - each a
thin marshalling layer: allocate outputs / workspace with torch, pass raw device
pointers and the current CUDA stream to the C ABI, never compute on the host.
There is no fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from . import _lib

# parameter order used for the flat lists handed to the autograd Functions
RSSM_ORDER = _lib.RSSM_FIELDS


@dataclass(frozen=True)
class PathConfig:
    """Static shape/config of one call (mirrors wmrl_dims)."""
    variant: int
    D: int
    S: int
    C: int
    A: int
    Hd: int
    E: int
    Ha: int = 0
    La: int = 0
    action_scale: float = 5.0
    precision: int = _lib.FP32
    record: bool = True   # autograd is recording (torch.is_grad_enabled() at the call site)

    @property
    def S_in(self) -> int:
        return self.S * (self.C if self.variant == _lib.DISCRETE else 1)

    @property
    def dist_width(self) -> int:
        return self.S_in if self.variant == _lib.DISCRETE else self.S

    def dims(self, B: int, T: int, flags: int = 0, precision: Optional[int] = None) -> _lib.Dims:
        return _lib.Dims(variant=self.variant, precision=self.precision if precision is None else precision,
                         B=B, T=T, D=self.D, S=self.S, C=self.C if self.variant == _lib.DISCRETE else 1,
                         A=self.A, Hd=self.Hd, E=self.E, Ha=self.Ha, La=self.La,
                         action_scale=self.action_scale, flags=flags)


def _require_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("wmrl: the RSSM path runs only on CUDA tensors (no CPU fallback); "
                               "move the modules and inputs to a B200 device")
        dev = dev or t.device
        if t.device != dev:
            raise RuntimeError(f"wmrl: tensors on different devices ({t.device} vs {dev})")
    return dev


_WARNED = set()


def _warn_once(key: str, msg: str) -> None:
    if key not in _WARNED:
        _WARNED.add(key)
        import warnings
        warnings.warn(msg, RuntimeWarning, stacklevel=3)


def effective_imagine_precision(cfg: "PathConfig", B: int, H: int, need_actor_grad: bool, need_rssm_grad: bool) -> str:
    """'bf16' or 'fp32': which kernels an ``imagine_rollout`` call with this configuration runs."""
    if cfg.precision != _lib.BF16:
        return "fp32"
    if not (need_actor_grad or need_rssm_grad):
        return "bf16"
    dq = cfg.dims(B, H, _lib.FLAG_SAVE_FOR_BWD, _lib.BF16)
    return "bf16" if (not need_rssm_grad and _lib.load().wmrl_bf16_train_supported(C.byref(dq))) else "fp32"


def _on(dev):
    """libwmrl launches kernels, uploads constants and sets function attributes on the CUDA *current*
    device; every native call therefore runs with the tensors' device made current (the modules may live
    on a device other than torch's current one, as with the reference)."""
    return torch.cuda.device(dev)


def _native(dev, fn, *args) -> None:
    """one C-ABI call with ``dev`` current; non-zero status raises with wmrl_last_error()."""
    with _on(dev):
        _lib.check(fn(*args), fn.__name__)


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream_ptr(dev) -> int:
    """cudaStream_t of torch's current stream on ``dev`` (the raw accessor is ~10x cheaper than building a Stream
    object; the acting step is host-latency-bound)."""
    if _raw_stream is not None and dev.index is not None:
        return _raw_stream(dev.index)
    return torch.cuda.current_stream(dev).cuda_stream


def _rssm_ptrs(params: Sequence[Optional[torch.Tensor]]) -> _lib.RssmPtrs:
    s = _lib.RssmPtrs()
    for name, t in zip(RSSM_ORDER, params):
        setattr(s, name, _lib.ptr(t))
    return s


def _actor_layout(La: int) -> List[str]:
    """flat order of actor tensors: per block (lin_w, lin_b, ln_g, ln_b), then mu_w, mu_b."""
    names = []
    for l in range(La):
        names += [f"lin_w{l}", f"lin_b{l}", f"ln_g{l}", f"ln_b{l}"]
    return names + ["mu_w", "mu_b"]


def _actor_ptrs(params: Sequence[Optional[torch.Tensor]], La: int, bound: Optional[torch.Tensor], grads=False):
    s = _lib.ActorGradPtrs() if grads else _lib.ActorPtrs()
    for l in range(La):
        s.lin_w[l] = _lib.ptr(params[4 * l + 0])
        s.lin_b[l] = _lib.ptr(params[4 * l + 1])
        s.ln_g[l] = _lib.ptr(params[4 * l + 2])
        s.ln_b[l] = _lib.ptr(params[4 * l + 3])
    s.mu_w = _lib.ptr(params[4 * La])
    s.mu_b = _lib.ptr(params[4 * La + 1])
    if not grads:
        s.bound = _lib.ptr(bound)
    return s


def _bytes(n: int, dev) -> torch.Tensor:
    return torch.empty(max(int(n), 16), dtype=torch.uint8, device=dev)


# ----------------------------------------------------------------------------
# packed bf16 weight image cache (refreshed when a parameter's version changes)
# ----------------------------------------------------------------------------
class PackedWeights:
    def __init__(self):
        self.key = None
        self.buf: Optional[torch.Tensor] = None

    def get(self, cfg: PathConfig, rssm_p: Sequence[torch.Tensor], actor_p: Sequence[torch.Tensor],
            bound: torch.Tensor) -> torch.Tensor:
        lib = _lib.load()
        key = (dataclasses.replace(cfg, record=True), bound.data_ptr(), bound._version,
               tuple((p.data_ptr(), p._version) for p in list(rssm_p) + list(actor_p)))
        if self.key == key and self.buf is not None:
            return self.buf
        dev = rssm_p[0].device
        d = cfg.dims(0, 0, precision=_lib.BF16)
        n = lib.wmrl_packed_bytes(C.byref(d))
        if n == 0:
            raise RuntimeError("wmrl: this shape is not supported by the bf16 tcgen05 path; use precision='fp32'")
        buf = _bytes(n, dev)
        r32 = [_f32c(p) for p in rssm_p]
        a32 = [_f32c(p) for p in actor_p]
        w, aw = _rssm_ptrs(r32), _actor_ptrs(a32, cfg.La, bound)
        _native(dev, lib.wmrl_pack_weights, C.byref(d), C.byref(w), C.byref(aw), buf.data_ptr(), _stream_ptr(dev))
        self.key, self.buf = key, buf
        return buf


class PackedObserveWeights:
    """bf16 image of the RSSM weights (prior and posterior heads) for the filtering forward on the tcgen05 chain."""

    def __init__(self):
        self.key = None
        self.buf: Optional[torch.Tensor] = None

    def get(self, cfg: PathConfig, rssm_p: Sequence[torch.Tensor]) -> Optional[torch.Tensor]:
        lib = _lib.load()
        key = (dataclasses.replace(cfg, record=True), tuple((p.data_ptr(), p._version) for p in rssm_p))
        if self.key == key and self.buf is not None:
            return self.buf
        dev = rssm_p[0].device
        d = cfg.dims(0, 0, precision=_lib.BF16)
        n = lib.wmrl_observe_packed_bytes(C.byref(d))
        if n == 0:
            return None
        buf = _bytes(n, dev)
        r32 = [_f32c(p) for p in rssm_p]
        w = _rssm_ptrs(r32)
        _native(dev, lib.wmrl_pack_observe_weights, C.byref(d), C.byref(w), buf.data_ptr(), _stream_ptr(dev))
        self.key, self.buf = key, buf
        return buf


class ObserveRolloutBF16(torch.autograd.Function):
    """RSSMBase.observe_rollout on the bf16 layer-wise tcgen05 chain, forward (wmrl_observe_fwd_bf16) and full backward
    (wmrl_observe_bwd_bf16: RSSM weight gradients, d obs_embeds, d actions).  Same inputs / outputs as
    ``ObserveRollout`` plus the packed weight image."""

    @staticmethod
    def forward(ctx, cfg: PathConfig, packed, obs, act, noise_pri, noise_post, init_pri_da, init_pri_db, init_post_da,
                init_post_db, *rssm_p):
        lib = _lib.load()
        dev = _require_cuda(obs, act, noise_pri, noise_post, *rssm_p)
        B, T = obs.shape[0], obs.shape[1]
        disc = cfg.variant == _lib.DISCRETE
        need_bwd = cfg.record and any(ctx.needs_input_grad)
        ob, ac, n1, n2 = _f32c(obs), _f32c(act), _f32c(noise_pri), _f32c(noise_post)

        def seq_bufs():
            de = torch.empty(B, T, cfg.D, device=dev)
            st = torch.empty(B, T, cfg.S_in, device=dev)
            if disc:
                return de, st, torch.empty(B, T, cfg.C, cfg.S, device=dev), None
            return de, st, torch.empty(B, T, cfg.S, device=dev), torch.empty(B, T, cfg.S, device=dev)

        pri, post = seq_bufs(), seq_bufs()
        flags = _lib.FLAG_SAVE_FOR_BWD if need_bwd else 0
        d = cfg.dims(B, T, flags, _lib.BF16)
        saved_n = lib.wmrl_saved_bytes(C.byref(d), _lib.OP_OBSERVE) if need_bwd else 0
        saved = _bytes(saved_n, dev) if need_bwd else None
        ws_n = lib.wmrl_workspace_bytes(C.byref(d), _lib.OP_OBSERVE)
        ws = _bytes(ws_n, dev)
        if B > 0 and T > 0:
            _native(dev, lib.wmrl_observe_fwd_bf16, C.byref(d), packed.data_ptr(), ob.data_ptr(), ac.data_ptr(), n1.data_ptr(),
                    n2.data_ptr(), pri[0].data_ptr(), pri[1].data_ptr(), pri[2].data_ptr(), _lib.ptr(pri[3]),
                    post[0].data_ptr(), post[1].data_ptr(), post[2].data_ptr(), _lib.ptr(post[3]), _lib.ptr(saved), saved_n,
                    ws.data_ptr(), ws_n, _stream_ptr(dev))
        if T > 0:
            pri[2][:, 0] = init_pri_da.detach()
            post[2][:, 0] = init_post_da.detach()
            if not disc:
                pri[3][:, 0] = init_pri_db.detach()
                post[3][:, 0] = init_post_db.detach()
        ctx.cfg, ctx.disc, ctx.packed = cfg, disc, packed
        ctx.shape = (B, T, obs.shape[2], act.shape[2])
        ctx.n_params = len(rssm_p)
        if need_bwd:
            empty = torch.empty(0, device=dev)
            ctx.save_for_backward(n1, n2, pri[0], pri[2], pri[3] if pri[3] is not None else empty, post[2],
                                  post[3] if post[3] is not None else empty, saved, *[p.detach() for p in rssm_p])
        if disc:
            return pri[0], pri[1], pri[2], post[0], post[1], post[2]
        return pri + post

    @staticmethod
    def backward(ctx, *gouts):
        lib = _lib.load()
        cfg, disc = ctx.cfg, ctx.disc
        n1, n2, p_de, p_da, p_db, q_da, q_db, saved, *params = ctx.saved_tensors
        dev = p_de.device
        B, T, E, A = ctx.shape

        def prep(g):
            return None if g is None else g.to(torch.float32).contiguous()

        g = [prep(x) for x in gouts]
        if disc:
            gp_de, gp_st, gp_da, gp_db, gq_de, gq_st, gq_da, gq_db = g[0], g[1], g[2], None, g[3], g[4], g[5], None
        else:
            gp_de, gp_st, gp_da, gp_db, gq_de, gq_st, gq_da, gq_db = g
        g_de = gp_de if gq_de is None else (gq_de if gp_de is None else gp_de + gq_de)
        d = cfg.dims(B, T, _lib.FLAG_SAVE_FOR_BWD | _lib.FLAG_BACKWARD, _lib.BF16)
        gr = [torch.zeros(p.shape, device=dev, dtype=torch.float32) for p in params]
        g_obs = torch.zeros(B, T, E, device=dev)
        g_act = torch.zeros(B, T, A, device=dev)
        ws_n = lib.wmrl_workspace_bytes(C.byref(d), _lib.OP_OBSERVE)
        ws = _bytes(ws_n, dev)
        gw = _rssm_ptrs(gr)
        if B > 0 and T > 1:
            _native(dev, lib.wmrl_observe_bwd_bf16, C.byref(d), ctx.packed.data_ptr(), n1.data_ptr(), n2.data_ptr(),
                    p_de.data_ptr(), p_da.data_ptr(), None if disc else p_db.data_ptr(), q_da.data_ptr(),
                    None if disc else q_db.data_ptr(), saved.data_ptr(), _lib.ptr(g_de), _lib.ptr(gp_st), _lib.ptr(gp_da),
                    _lib.ptr(gp_db), _lib.ptr(gq_st), _lib.ptr(gq_da), _lib.ptr(gq_db), C.byref(gw), g_obs.data_ptr(),
                    g_act.data_ptr(), ws.data_ptr(), ws_n, _stream_ptr(dev))
        need = ctx.needs_input_grad

        def first(gseq):
            return None if gseq is None else gseq[:, 0]

        grads = [None, None, g_obs if need[2] else None, g_act if need[3] else None, None, None,
                 first(gp_da) if need[6] else None, first(gp_db) if need[7] else None,
                 first(gq_da) if need[8] else None, first(gq_db) if need[9] else None]
        for i, gt in enumerate(gr):
            grads.append(gt if need[10 + i] else None)
        return tuple(grads)


# ----------------------------------------------------------------------------
# imagine
# ----------------------------------------------------------------------------
class ImagineRollout(torch.autograd.Function):
    """outputs: deter_seq[B,H+1,D], stoch_seq[B,H+1,S_in], dist_a, dist_b, actions[B,H,A], idx.

    inputs after the static ones: deter0, stoch0, init_da, init_db (initial
    mean/std or logits -> index 0 of the dist outputs), noise, bound, then the
    actor tensors (``_actor_layout``) and the 14 RSSM tensors (``RSSM_ORDER``).
    """

    @staticmethod
    def forward(ctx, cfg: PathConfig, H: int, packed, deter0, stoch0, init_da, init_db, noise, bound, *params):
        lib = _lib.load()
        n_actor = 4 * cfg.La + 2
        actor_p, rssm_p = params[:n_actor], params[n_actor:]
        dev = _require_cuda(deter0, stoch0, noise, bound, *params)
        B = deter0.shape[0]
        disc = cfg.variant == _lib.DISCRETE
        need_bwd = cfg.record and any(ctx.needs_input_grad)
        a32 = [_f32c(p) for p in actor_p]
        r32 = [_f32c(p) for p in rssm_p]
        d0, s0, nz, bd = _f32c(deter0), _f32c(stoch0), _f32c(noise), _f32c(bound)
        T1 = H + 1
        deter_seq = torch.empty(B, T1, cfg.D, device=dev)
        stoch_seq = torch.empty(B, T1, cfg.S_in, device=dev)
        if disc:
            dist_a = torch.empty(B, T1, cfg.C, cfg.S, device=dev)
            dist_b = None
            idx = torch.empty(B, H, cfg.C, dtype=torch.int32, device=dev)
        else:
            dist_a = torch.empty(B, T1, cfg.S, device=dev)
            dist_b = torch.empty(B, T1, cfg.S, device=dev)
            idx = None
        actions = torch.empty(B, H, cfg.A, device=dev)
        # bf16 + autograd: the persistent rollout kernel records a bf16 activation stash and the backward is the
        # persistent tcgen05 BPTT kernel (continuous variant, world model frozen as in world_model.py:172).
        # Shapes / cases it does not cover (RSSM weight gradients requested, discrete latent) record with the
        # fp32 chain instead - loudly, once per process.
        precision = cfg.precision
        if need_bwd and precision == _lib.BF16:
            need_rssm = any(ctx.needs_input_grad[9 + n_actor:])
            dq = cfg.dims(B, H, _lib.FLAG_SAVE_FOR_BWD, _lib.BF16)
            if need_rssm or not lib.wmrl_bf16_train_supported(C.byref(dq)):
                _warn_once("bf16-train-fallback",
                           "wmrl: precision='bf16' with autograd recording runs the fp32 tensor-core chain for this call "
                           + ("(RSSM weight gradients requested: the bf16 BPTT kernel covers the frozen-world-model case)"
                              if need_rssm else "(this shape / variant has no bf16 BPTT kernel)"))
                precision = _lib.FP32
        flags = _lib.FLAG_SAVE_FOR_BWD if need_bwd else 0
        d = cfg.dims(B, H, flags, precision)
        saved_n = lib.wmrl_saved_bytes(C.byref(d), _lib.OP_IMAGINE) if need_bwd else 0
        saved = _bytes(saved_n, dev) if need_bwd else None
        ws_n = lib.wmrl_workspace_bytes(C.byref(d), _lib.OP_IMAGINE)
        ws = _bytes(ws_n, dev)
        w, aw = _rssm_ptrs(r32), _actor_ptrs(a32, cfg.La, bd)
        if B > 0 and H > 0:
            _native(dev, lib.wmrl_imagine_fwd,
                C.byref(d), C.byref(w), C.byref(aw), _lib.ptr(packed) if precision == _lib.BF16 else None,
                d0.data_ptr(), s0.data_ptr(), nz.data_ptr(), deter_seq.data_ptr(), stoch_seq.data_ptr(),
                dist_a.data_ptr(), _lib.ptr(dist_b), actions.data_ptr(), _lib.ptr(idx), _lib.ptr(saved),
                saved_n, ws.data_ptr(), ws_n, _stream_ptr(dev))
        else:
            deter_seq[:, 0], stoch_seq[:, 0] = d0, s0
        dist_a[:, 0] = init_da.detach()
        if dist_b is not None:
            dist_b[:, 0] = init_db.detach()
        ctx.cfg, ctx.H, ctx.n_actor = cfg, H, n_actor
        ctx.has_b = dist_b is not None
        ctx.precision = precision
        ctx.packed = packed if precision == _lib.BF16 else None
        if need_bwd:
            ctx.save_for_backward(nz, bd, deter_seq, stoch_seq, dist_a,
                                  dist_b if dist_b is not None else torch.empty(0, device=dev),
                                  actions, saved, *a32, *r32)
        outs = (deter_seq, stoch_seq, dist_a)
        if dist_b is not None:
            outs += (dist_b,)
        ctx.mark_non_differentiable(actions)
        if idx is not None:
            ctx.mark_non_differentiable(idx)
            return outs + (actions, idx)
        return outs + (actions,)

    @staticmethod
    def backward(ctx, *gouts):
        lib = _lib.load()
        cfg, H, n_actor = ctx.cfg, ctx.H, ctx.n_actor
        nz, bd, deter_seq, stoch_seq, dist_a, dist_b, actions, saved, *rest = ctx.saved_tensors
        a32, r32 = rest[:n_actor], rest[n_actor:]
        dev = deter_seq.device
        B = deter_seq.shape[0]
        g_deter, g_stoch, g_da = gouts[0], gouts[1], gouts[2]
        g_db = gouts[3] if ctx.has_b else None

        def prep(g):
            return None if g is None else g.to(torch.float32).contiguous()

        g_deter, g_stoch, g_da, g_db = prep(g_deter), prep(g_stoch), prep(g_da), prep(g_db)
        # which inputs need gradients (inputs: cfg,H,packed,deter0,stoch0,init_da,init_db,noise,bound,*params)
        need = ctx.needs_input_grad
        base = 9
        need_rssm = any(need[base + n_actor:])
        flags = _lib.FLAG_SAVE_FOR_BWD | _lib.FLAG_BACKWARD | (_lib.FLAG_WM_WGRAD if need_rssm else 0)
        d = cfg.dims(B, H, flags, ctx.precision)
        ga = [torch.zeros_like(p) for p in a32]
        gr = [torch.zeros_like(p) for p in r32] if need_rssm else [None] * len(r32)
        g_d0 = torch.zeros(B, cfg.D, device=dev)
        g_s0 = torch.zeros(B, cfg.S_in, device=dev)
        ws_n = lib.wmrl_workspace_bytes(C.byref(d), _lib.OP_IMAGINE)
        ws = _bytes(ws_n, dev)
        w, aw = _rssm_ptrs(r32), _actor_ptrs(a32, cfg.La, bd)
        gaw = _actor_ptrs(ga, cfg.La, None, grads=True)
        gw = _rssm_ptrs(gr)
        if B > 0 and H > 0:
            _native(dev, lib.wmrl_imagine_bwd,
                C.byref(d), C.byref(w), C.byref(aw), _lib.ptr(ctx.packed), nz.data_ptr(), deter_seq.data_ptr(),
                stoch_seq.data_ptr(), dist_a.data_ptr(), _lib.ptr(dist_b if ctx.has_b else None),
                actions.data_ptr(), saved.data_ptr(), _lib.ptr(g_deter), _lib.ptr(g_stoch), _lib.ptr(g_da),
                _lib.ptr(g_db), C.byref(gaw), C.byref(gw) if need_rssm else None, g_d0.data_ptr(),
                g_s0.data_ptr(), ws.data_ptr(), ws_n, _stream_ptr(dev))
        else:
            if g_deter is not None:
                g_d0 = g_deter[:, 0].clone()
            if g_stoch is not None:
                g_s0 = g_stoch[:, 0].clone()
        g_init_da = g_da[:, 0] if (g_da is not None and need[5]) else None
        g_init_db = g_db[:, 0] if (g_db is not None and need[6]) else None
        grads = [None, None, None, g_d0 if need[3] else None, g_s0 if need[4] else None, g_init_da, g_init_db,
                 None, None]
        for i, g in enumerate(ga):
            grads.append(g if need[base + i] else None)
        for i, g in enumerate(gr):
            grads.append(g if (g is not None and need[base + n_actor + i]) else None)
        return tuple(grads)


# ----------------------------------------------------------------------------
# observe
# ----------------------------------------------------------------------------
class ObserveRollout(torch.autograd.Function):
    """inputs: cfg, obs_embeds[B,T,E], actions[B,T,A], noise_prior, noise_post, init_da, init_db
    (index-0 dist slots of both sequences), 14 RSSM tensors.
    outputs: (pri_deter, pri_stoch, pri_da[, pri_db], post_deter, post_stoch, post_da[, post_db])."""

    @staticmethod
    def forward(ctx, cfg: PathConfig, obs, act, noise_pri, noise_post, init_pri_da, init_pri_db, init_post_da,
                init_post_db, *rssm_p):
        lib = _lib.load()
        dev = _require_cuda(obs, act, *rssm_p)
        B, T = obs.shape[0], obs.shape[1]
        disc = cfg.variant == _lib.DISCRETE
        need_bwd = cfg.record and any(ctx.needs_input_grad)
        r32 = [_f32c(p) for p in rssm_p]
        ob, ac = _f32c(obs), _f32c(act)
        n1 = _f32c(noise_pri) if noise_pri is not None else None
        n2 = _f32c(noise_post) if noise_post is not None else None

        def seq_bufs():
            de = torch.empty(B, T, cfg.D, device=dev)
            st = torch.empty(B, T, cfg.S_in, device=dev)
            if disc:
                return de, st, torch.empty(B, T, cfg.C, cfg.S, device=dev), None
            return de, st, torch.empty(B, T, cfg.S, device=dev), torch.empty(B, T, cfg.S, device=dev)

        pri, post = seq_bufs(), seq_bufs()
        flags = _lib.FLAG_SAVE_FOR_BWD if need_bwd else 0
        d = cfg.dims(B, T, flags, _lib.FP32)
        saved_n = lib.wmrl_saved_bytes(C.byref(d), _lib.OP_OBSERVE) if need_bwd else 0
        saved = _bytes(saved_n, dev) if need_bwd else None
        ws_n = lib.wmrl_workspace_bytes(C.byref(d), _lib.OP_OBSERVE)
        ws = _bytes(ws_n, dev)
        w = _rssm_ptrs(r32)
        if B > 0 and T > 0:
            _native(dev, lib.wmrl_observe_fwd,
                C.byref(d), C.byref(w), ob.data_ptr(), ac.data_ptr(), _lib.ptr(n1), _lib.ptr(n2),
                pri[0].data_ptr(), pri[1].data_ptr(), pri[2].data_ptr(), _lib.ptr(pri[3]),
                post[0].data_ptr(), post[1].data_ptr(), post[2].data_ptr(), _lib.ptr(post[3]),
                _lib.ptr(saved), saved_n, ws.data_ptr(), ws_n, _stream_ptr(dev))
        if T > 0:
            pri[2][:, 0] = init_pri_da.detach()
            post[2][:, 0] = init_post_da.detach()
            if not disc:
                pri[3][:, 0] = init_pri_db.detach()
                post[3][:, 0] = init_post_db.detach()
        ctx.cfg, ctx.disc = cfg, disc
        if need_bwd:
            empty = torch.empty(0, device=dev)
            ctx.save_for_backward(ob, ac, n1 if n1 is not None else empty, n2 if n2 is not None else empty,
                                  pri[0], pri[1], pri[2], pri[3] if pri[3] is not None else empty,
                                  post[0], post[1], post[2], post[3] if post[3] is not None else empty,
                                  saved, *r32)
        if disc:
            return pri[0], pri[1], pri[2], post[0], post[1], post[2]
        return pri + post

    @staticmethod
    def backward(ctx, *gouts):
        lib = _lib.load()
        cfg, disc = ctx.cfg, ctx.disc
        (ob, ac, n1, n2, p_de, p_st, p_da, p_db, q_de, q_st, q_da, q_db, saved, *r32) = ctx.saved_tensors
        dev = ob.device
        B, T = ob.shape[0], ob.shape[1]

        def prep(g):
            return None if g is None else g.to(torch.float32).contiguous()

        if disc:
            g = [prep(x) for x in gouts]
            gp = (g[0], g[1], g[2], None)
            gq = (g[3], g[4], g[5], None)
        else:
            g = [prep(x) for x in gouts]
            gp, gq = tuple(g[0:4]), tuple(g[4:8])
        d = cfg.dims(B, T, _lib.FLAG_SAVE_FOR_BWD, _lib.FP32)
        gr = [torch.zeros_like(p) for p in r32]
        g_obs = torch.zeros_like(ob)
        g_act = torch.zeros_like(ac)
        ws_n = lib.wmrl_workspace_bytes(C.byref(d), _lib.OP_OBSERVE)
        ws = _bytes(ws_n, dev)
        w, gw = _rssm_ptrs(r32), _rssm_ptrs(gr)
        if B > 0 and T > 1:
            _native(dev, lib.wmrl_observe_bwd,
                C.byref(d), C.byref(w), ob.data_ptr(), ac.data_ptr(), n1.data_ptr(), n2.data_ptr(),
                p_de.data_ptr(), p_st.data_ptr(), p_da.data_ptr(), None if disc else p_db.data_ptr(),
                q_de.data_ptr(), q_st.data_ptr(), q_da.data_ptr(), None if disc else q_db.data_ptr(),
                saved.data_ptr(),
                _lib.ptr(gp[0]), _lib.ptr(gp[1]), _lib.ptr(gp[2]), _lib.ptr(gp[3]),
                _lib.ptr(gq[0]), _lib.ptr(gq[1]), _lib.ptr(gq[2]), _lib.ptr(gq[3]),
                C.byref(gw), g_obs.data_ptr(), g_act.data_ptr(), ws.data_ptr(), ws_n, _stream_ptr(dev))
        need = ctx.needs_input_grad

        def first(gseq):
            return None if gseq is None else gseq[:, 0]

        grads = [None, g_obs if need[1] else None, g_act if need[2] else None, None, None,
                 first(gp[2]) if need[5] else None, first(gp[3]) if need[6] else None,
                 first(gq[2]) if need[7] else None, first(gq[3]) if need[8] else None]
        for i, gt in enumerate(gr):
            grads.append(gt if need[9 + i] else None)
        return tuple(grads)


# ----------------------------------------------------------------------------
# lambda-return
# ----------------------------------------------------------------------------
class LambdaReturn(torch.autograd.Function):
    """targets[B,H] from V, r, d [B,H] (value_trainer.py:62-134 as one scan).
    Linear in (V, r) for a given mask d, so the backward is the transposed scan;
    no gradient flows through d (value_trainer.py:149 thresholds it)."""

    @staticmethod
    def forward(ctx, V, r, d, gamma: float, lam: float):
        lib = _lib.load()
        dev = _require_cuda(V, r, d)
        B, H = V.shape
        v32, r32, d32 = _f32c(V), _f32c(r), _f32c(d)
        out = torch.empty(B, H, device=dev)
        if B > 0:
            _native(dev, lib.wmrl_lambda_return_fwd, v32.data_ptr(), r32.data_ptr(), d32.data_ptr(), gamma, lam,
                                                  B, H, out.data_ptr(), _stream_ptr(dev))
        ctx.save_for_backward(d32)
        ctx.gamma, ctx.lam = gamma, lam
        return out

    @staticmethod
    def backward(ctx, g_out):
        lib = _lib.load()
        (d32,) = ctx.saved_tensors
        B, H = d32.shape
        g = g_out.to(torch.float32).contiguous()
        g_V, g_r = torch.empty_like(g), torch.empty_like(g)
        if B > 0:
            _native(g.device, lib.wmrl_lambda_return_bwd, g.data_ptr(), d32.data_ptr(), ctx.gamma, ctx.lam, B, H,
                                                  g_V.data_ptr(), g_r.data_ptr(), _stream_ptr(g.device))
        return g_V, g_r, None, None, None


def lambda_return(V: torch.Tensor, r: torch.Tensor, d: torch.Tensor, gamma: float, lam: float) -> torch.Tensor:
    """V, r, d: [B,H,1] (or [B,H]) -> targets of the same shape."""
    shape = V.shape
    if V.dim() == 3:
        V, r, d = V[..., 0], r[..., 0], d[..., 0]
    if V.shape[1] < 1:
        raise ValueError("lambda_return needs H >= 1")
    return LambdaReturn.apply(V, r, d, float(gamma), float(lam)).reshape(shape)


class KLLoss(torch.autograd.Function):
    """(mean max(KL, rho), mean KL) of prior || posterior over [B, T-1] (world_model.py:131-141) in one fused pass
    each way.  Inputs are the sequences' distribution tensors [B,T,.]; b tensors are None for the discrete variant."""

    @staticmethod
    def forward(ctx, cfg: PathConfig, rho: float, pa, pb, qa, qb):
        lib = _lib.load()
        dev = _require_cuda(pa, qa)
        B, T = pa.shape[0], pa.shape[1]
        pa32, qa32 = _f32c(pa), _f32c(qa)
        pb32 = None if pb is None else _f32c(pb)
        qb32 = None if qb is None else _f32c(qb)
        kl = torch.empty(B, max(T - 1, 0), device=dev)
        sums = torch.zeros(2, device=dev)
        d = cfg.dims(B, T, 0, _lib.FP32)
        if B > 0 and T > 1:
            _native(dev, lib.wmrl_kl_loss_fwd, C.byref(d), pa32.data_ptr(), _lib.ptr(pb32), qa32.data_ptr(), _lib.ptr(qb32),
                    float(rho), kl.data_ptr(), sums.data_ptr(), _stream_ptr(dev))
        ctx.cfg, ctx.rho, ctx.has_b = cfg, float(rho), pb is not None
        empty = torch.empty(0, device=dev)
        ctx.save_for_backward(pa32, pb32 if pb32 is not None else empty, qa32, qb32 if qb32 is not None else empty)
        n = max(B * (T - 1), 1)
        ctx.mark_non_differentiable(kl)
        return sums[0] / n, sums[1] / n, kl

    @staticmethod
    def backward(ctx, g_clamped, g_mean, _g_kl):
        lib = _lib.load()
        pa, pb, qa, qb = ctx.saved_tensors
        if g_mean is not None and bool((g_mean != 0).any()):
            raise RuntimeError("wmrl: KLLoss differentiates the clamped mean only (the reference reports the plain mean as a float)")
        B, T = pa.shape[0], pa.shape[1]
        dev = pa.device
        d = ctx.cfg.dims(B, T, 0, _lib.FP32)
        g_pa, g_qa = torch.zeros_like(pa), torch.zeros_like(qa)
        g_pb = torch.zeros_like(pb) if ctx.has_b else None
        g_qb = torch.zeros_like(qb) if ctx.has_b else None
        if B > 0 and T > 1:
            scale = float(g_clamped) / (B * (T - 1))
            _native(dev, lib.wmrl_kl_loss_bwd, C.byref(d), pa.data_ptr(), _lib.ptr(pb if ctx.has_b else None), qa.data_ptr(),
                    _lib.ptr(qb if ctx.has_b else None), ctx.rho, scale, g_pa.data_ptr(), _lib.ptr(g_pb), g_qa.data_ptr(),
                    _lib.ptr(g_qb), _stream_ptr(dev))
        return None, None, g_pa, g_pb, g_qa, g_qb


def kl_loss(cfg: PathConfig, rho: float, prior_a, prior_b, post_a, post_b):
    """-> (mean over [B,T-1] of max(KL, rho), mean KL, kl[B,T-1])."""
    return KLLoss.apply(cfg, float(rho), prior_a, prior_b, post_a, post_b)


class ActStepPlan:
    """Everything of an online acting step that does not change between environment steps: the validated parameter
    lists, the pointer structs and the dims struct per batch size.  At B = 1 the step is host-latency-bound (the kernel
    takes ~30 us), so the per-call Python work is kept to pointer comparisons, one allocation and the ctypes call.
    ``valid()`` is false once any parameter moved (``.to()``, ``param.data = ...`` as WorldModelActor.perturb_actor does)."""

    def __init__(self, cfg: PathConfig, rssm_p: Sequence[torch.Tensor], actor_p: Sequence[torch.Tensor], bound: torch.Tensor):
        self.cfg = cfg
        self.dev = _require_cuda(*rssm_p, *actor_p, bound)
        self.tensors = list(rssm_p) + list(actor_p) + [bound]
        self.ptrs = [t.data_ptr() for t in self.tensors]
        # fp32 contiguous parameters are passed by reference (optimiser steps are seen); anything else would need a
        # converted copy, which must not be cached
        self.cacheable = all(t.dtype == torch.float32 and t.is_contiguous() for t in self.tensors)
        self.keep = [_f32c(p) for p in rssm_p], [_f32c(p) for p in actor_p]
        self.w = _rssm_ptrs(self.keep[0])
        self.aw = _actor_ptrs(self.keep[1], cfg.La, bound)
        self.dims = {}
        disc = cfg.variant == _lib.DISCRETE
        W = cfg.dist_width
        self.widths = [cfg.A, cfg.S_in, W, cfg.D, cfg.S_in, W] + ([] if disc else [W, W])
        self.nlat = cfg.S_in if disc else cfg.S

    def valid(self) -> bool:
        return self.cacheable and [t.data_ptr() for t in self.tensors] == self.ptrs

    def run(self, obs_embed, deter, noise_post, noise_prior, sampling=None):
        lib = _lib.load()
        cfg, dev = self.cfg, self.dev
        if deter.device != dev or obs_embed.device != dev:
            raise RuntimeError(f"wmrl: act_step inputs must live on {dev} (no CPU fallback)")
        B = deter.shape[0]
        fast = lambda t: t if (t.dtype == torch.float32 and t.is_contiguous() and not t.requires_grad) else _f32c(t)  # noqa: E731
        e, h = fast(obs_embed), fast(deter)
        if e.shape != (B, cfg.E) or h.shape != (B, cfg.D):
            raise ValueError(f"act_step: obs_embed {tuple(e.shape)} / deter {tuple(h.shape)} do not match E={cfg.E}, D={cfg.D}")
        n1, n2 = fast(noise_post), fast(noise_prior)
        if n1.numel() != B * self.nlat or n2.numel() != B * self.nlat:
            raise ValueError("act_step: noise tensors must hold one draw per latent element")
        disc = cfg.variant == _lib.DISCRETE
        widths = self.widths + ([cfg.A, cfg.A, 1] if sampling else [])
        flat = torch.empty(B, sum(widths), device=dev)      # one allocation for all outputs
        parts, o = [], 0
        for wd in widths:
            parts.append(flat[:, o:o + wd])
            o += wd
        action, post_s, post_a, deter_n, pri_s, pri_a = parts[:6]
        post_b, pri_b = (None, None) if disc else (parts[6], parts[7])
        d = self.dims.get(B)
        if d is None:
            d = self.dims[B] = cfg.dims(B, 1, 0, _lib.FP32)
        smp = None
        if sampling:
            sw, sb, na = (fast(t) for t in sampling)
            if sw.shape != (cfg.A, cfg.Ha) or na.numel() != B * cfg.A:
                raise ValueError("act_step: sampling = (stddev.weight [A,Ha], stddev.bias [A], noise_action [B,A])")
            smp = _lib.ActSampling(std_w=sw.data_ptr(), std_b=sb.data_ptr(), noise_action=na.data_ptr(),
                                   act_mean=parts[-3].data_ptr(), act_std=parts[-2].data_ptr(), entropy=parts[-1].data_ptr())
        if B > 0:
            base, ld = flat.data_ptr(), flat.shape[1]
            offs, o = [], 0
            for wd in widths:
                offs.append(base + 4 * o)
                o += wd
            args = (C.byref(d), C.byref(self.w), C.byref(self.aw), e.data_ptr(), h.data_ptr(), n1.data_ptr(), n2.data_ptr(),
                    offs[0], offs[1], offs[2], None if disc else offs[6], offs[3], offs[4], offs[5], None if disc else offs[7],
                    ld, C.byref(smp) if smp is not None else None, _stream_ptr(dev))
            if torch.cuda.current_device() == dev.index:
                rc = lib.wmrl_act_step(*args)
            else:
                with torch.cuda.device(dev):
                    rc = lib.wmrl_act_step(*args)
            if rc:
                _lib.check(rc, "wmrl_act_step")
        out = (action, post_s, post_a, post_b, deter_n, pri_s, pri_a, pri_b)
        if sampling:
            out += (parts[-3], parts[-2], parts[-1].reshape(B))
        return out


def act_step(cfg: PathConfig, rssm_p: Sequence[torch.Tensor], actor_p: Sequence[torch.Tensor], bound: torch.Tensor,
             obs_embed: torch.Tensor, deter: torch.Tensor, noise_post: torch.Tensor, noise_prior: torch.Tensor,
             sampling: Optional[tuple] = None):
    """One online acting step (memory_actor.py:34-57: posterior -> deterministic actor -> prior) as ONE cluster kernel
    (wmrl_act_step), fp32, no gradients.  -> (action, post_stoch, post_a, post_b, deter_next, prior_stoch, prior_a,
    prior_b); the *_b tensors are None for the discrete variant.  ``sampling`` = (stddev.weight, stddev.bias,
    noise_action[B,A]) switches the action to clamp(mean + std * noise, +-bound) (actor.py:54-59) and appends
    (act_mean, act_std, entropy) to the result.  (``RSSM.act_step`` keeps an ``ActStepPlan`` across calls instead.)"""
    return ActStepPlan(cfg, rssm_p, actor_p, bound).run(obs_embed, deter, noise_post, noise_prior, sampling)


class SequenceFeatures(torch.autograd.Function):
    """``cat(deter_states[:, 1:], stoch_states[:, 1:], -1)`` (state/continuous.py:73-77) as one streaming pass each way
    (wmrl_features_fwd / _bwd) - torch.cat and its backward's two gathers run at about a third of HBM bandwidth."""

    @staticmethod
    def forward(ctx, deter_seq, stoch_seq):
        lib = _lib.load()
        dev = _require_cuda(deter_seq, stoch_seq)
        d32, s32 = _f32c(deter_seq), _f32c(stoch_seq)
        B, T, D = d32.shape
        S = s32.shape[2]
        out = torch.empty(B, max(T - 1, 0), D + S, device=dev)
        if B > 0 and T > 1:
            _native(dev, lib.wmrl_features_fwd, d32.data_ptr(), s32.data_ptr(), B, T, D, S, out.data_ptr(), _stream_ptr(dev))
        ctx.dims = (B, T, D, S)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        B, T, D, S = ctx.dims
        g32 = _f32c(g)
        dev = g32.device
        gd, gs = torch.empty(B, T, D, device=dev), torch.empty(B, T, S, device=dev)
        if B > 0 and T > 1:
            _native(dev, lib.wmrl_features_bwd, g32.data_ptr(), B, T, D, S, gd.data_ptr(), gs.data_ptr(), _stream_ptr(dev))
        else:
            gd.zero_(); gs.zero_()
        return gd, gs


def sequence_features(deter_seq: torch.Tensor, stoch_seq: torch.Tensor) -> torch.Tensor:
    """features[B, T-1, D + S_in] of a [B, T, .] state sequence (index 0, the initial state, excluded)."""
    return SequenceFeatures.apply(deter_seq, stoch_seq)


class HeadLosses(torch.autograd.Function):
    """(reward NLL, done BCE-with-logits) of WorldModel.update (world_model.py:130, 143-144), values and both
    gradients in ONE pass (wmrl_head_losses).  reward loss = mean over rows of the summed unit-variance Normal NLL
    over the last dim; done loss = mean over all elements."""

    @staticmethod
    def forward(ctx, pred_r, tgt_r, logit_d, tgt_d):
        lib = _lib.load()
        dev = _require_cuda(pred_r, logit_d)
        pr, tr = _f32c(pred_r), _f32c(tgt_r.expand_as(pred_r))
        ld, td = _f32c(logit_d), _f32c(tgt_d.expand_as(logit_d))
        sums = torch.zeros(2, device=dev)
        need = pred_r.requires_grad or logit_d.requires_grad
        g_pr = torch.empty_like(pr) if need else None
        g_ld = torch.empty_like(ld) if need else None
        _native(dev, lib.wmrl_head_losses, pr.data_ptr(), tr.data_ptr(), pr.numel(), ld.data_ptr(), td.data_ptr(),
                ld.numel(), sums.data_ptr(), _lib.ptr(g_pr), _lib.ptr(g_ld), _stream_ptr(dev))
        rows_r = max(pr.numel() // max(pr.shape[-1], 1), 1)
        n_d = max(ld.numel(), 1)
        ctx.rows_r, ctx.n_d = rows_r, n_d
        ctx.shapes = (pred_r.shape, logit_d.shape)
        if need:
            ctx.save_for_backward(g_pr, g_ld)
        return sums[0] / rows_r, sums[1] / n_d

    @staticmethod
    def backward(ctx, g_r, g_d):
        g_pr, g_ld = ctx.saved_tensors
        out_r = (g_pr * (g_r / ctx.rows_r)).reshape(ctx.shapes[0]) if ctx.needs_input_grad[0] else None
        out_d = (g_ld * (g_d / ctx.n_d)).reshape(ctx.shapes[1]) if ctx.needs_input_grad[2] else None
        return out_r, None, out_d, None


def head_losses(pred_reward, reward, done_logits, done):
    """-> (reward_loss, done_loss) as WorldModel.update computes them (world_model.py:143-144)."""
    return HeadLosses.apply(pred_reward, reward, done_logits, done)


def done_mask(dones: torch.Tensor) -> torch.Tensor:
    """shift right, threshold at 0.5, running max (value_trainer.py:148-149)."""
    lib = _lib.load()
    dev = _require_cuda(dones)
    shape = dones.shape
    d2 = _f32c(dones.reshape(shape[0], shape[1]))
    out = torch.empty_like(d2)
    if d2.shape[0] > 0:
        _native(dev, lib.wmrl_done_mask, d2.data_ptr(), d2.shape[0], d2.shape[1], out.data_ptr(), _stream_ptr(dev))
    return out.reshape(shape)


# ----------------------------------------------------------------------------
# Linear primitive (the GEMM under every layer of the path)
# ----------------------------------------------------------------------------
class LinearFn(torch.autograd.Function):
    """y = act(x W^T + b) on libwmrl's GEMM kernels (3xTF32 tcgen05 for more than 64 rows, cluster
    split-K SIMT below), with dgrad / wgrad through the same kernels.  `act`: 0 none, 1 ELU."""

    @staticmethod
    def forward(ctx, x, weight, bias, act: int):
        lib = _lib.load()
        dev = _require_cuda(x, weight)
        x2 = _f32c(x.reshape(-1, x.shape[-1]))
        w = _f32c(weight)
        b = None if bias is None else _f32c(bias)
        M, K = x2.shape
        N = w.shape[0]
        if w.shape[1] != K:
            raise ValueError(f"linear: x has {K} features, weight expects {w.shape[1]}")
        y = torch.empty(M, N, device=dev)
        if M > 0:
            _native(dev, lib.wmrl_linear_fwd, M, N, K, x2.data_ptr(), K, w.data_ptr(), _lib.ptr(b), y.data_ptr(), N,
                                           int(act), _stream_ptr(dev))
        ctx.save_for_backward(x2, w, y if act else None)
        ctx.act, ctx.has_bias, ctx.x_shape = int(act), bias is not None, x.shape
        return y.reshape(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, g_y):
        lib = _lib.load()
        x2, w, y = ctx.saved_tensors
        M, K = x2.shape
        N = w.shape[0]
        g = _f32c(g_y.reshape(M, N))
        if ctx.act == 1:   # ELU'(z) = 1 for z > 0 else y + 1
            g = g * torch.where(y > 0, torch.ones_like(y), y + 1.0)
        dev = g.device
        g_x = g_w = g_b = None
        if M > 0 and ctx.needs_input_grad[0]:
            g_x = torch.empty(M, K, device=dev)
            _native(dev, lib.wmrl_linear_dgrad, M, N, K, g.data_ptr(), N, w.data_ptr(), g_x.data_ptr(), K, 0,
                                             _stream_ptr(dev))
            g_x = g_x.reshape(ctx.x_shape)
        elif ctx.needs_input_grad[0]:
            g_x = torch.zeros(ctx.x_shape, device=dev)
        if ctx.needs_input_grad[1]:
            g_w = torch.zeros(N, K, device=dev)
            if M > 0:
                _native(dev, lib.wmrl_linear_wgrad, M, N, K, g.data_ptr(), N, x2.data_ptr(), K, g_w.data_ptr(),
                                                 _stream_ptr(dev))
        if ctx.has_bias and ctx.needs_input_grad[2]:
            g_b = g.sum(0)
        return g_x, g_w, g_b, None


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, act: int = 0) -> torch.Tensor:
    """Drop-in for ``torch.nn.functional.linear`` (optionally fused ELU) on libwmrl's kernels."""
    return LinearFn.apply(x, weight, bias, act)


# ----------------------------------------------------------------------------
# fused LayerNorm + activation (the block after every hidden Linear of the actor and the MLP heads)
# ----------------------------------------------------------------------------
ACT_ELU, ACT_SILU, ACT_RELU = 0, 1, 2


class LayerNormAct(torch.autograd.Function):
    """y = act(LayerNorm(x) * gamma + beta), eps 1e-5, over the last dim; act: 0 ELU, 1 SiLU, 2 ReLU
    (actor.py:27-36, models.py:3-37, trainers/value/critic.py:4-29)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, act: int):
        lib = _lib.load()
        dev = _require_cuda(x, gamma, beta)
        x2 = _f32c(x.reshape(-1, x.shape[-1]))
        g, b = _f32c(gamma), _f32c(beta)
        M, W = x2.shape
        xhat, y = torch.empty_like(x2), torch.empty_like(x2)
        rstd = torch.empty(M, device=dev)
        if M > 0:
            _native(dev, lib.wmrl_ln_act_fwd, M, W, x2.data_ptr(), g.data_ptr(), b.data_ptr(), int(act), xhat.data_ptr(),
                                           rstd.data_ptr(), y.data_ptr(), _stream_ptr(dev))
        ctx.save_for_backward(xhat, rstd, g, b)
        ctx.act, ctx.x_shape = int(act), x.shape
        return y.reshape(x.shape)

    @staticmethod
    def backward(ctx, g_y):
        lib = _lib.load()
        xhat, rstd, g, b = ctx.saved_tensors
        M, W = xhat.shape
        dy = _f32c(g_y.reshape(M, W))
        dev = dy.device
        dx = torch.empty_like(xhat)
        need_p = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        dg = torch.zeros(W, device=dev) if need_p else None
        db = torch.zeros(W, device=dev) if need_p else None
        if M > 0:
            _native(dev, lib.wmrl_ln_act_bwd, M, W, dy.data_ptr(), xhat.data_ptr(), rstd.data_ptr(), g.data_ptr(),
                                           b.data_ptr(), ctx.act, dx.data_ptr(), _lib.ptr(dg), _lib.ptr(db),
                                           _stream_ptr(dev))
        return (dx.reshape(ctx.x_shape) if ctx.needs_input_grad[0] else None,
                dg if ctx.needs_input_grad[1] else None, db if ctx.needs_input_grad[2] else None, None)


def layer_norm_act(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, act: int) -> torch.Tensor:
    return LayerNormAct.apply(x, gamma, beta, act)


# ----------------------------------------------------------------------------
# MLP heads on the bf16 tcgen05 chain (wmrl_mlp_*): reward / done models, critic, target critic
# ----------------------------------------------------------------------------
def _mlp_pattern(seq: "torch.nn.Sequential"):
    """(blocks [(Linear, LayerNorm)], act id, output Linear) when ``seq`` is [Linear, LayerNorm, act] x depth -> Linear
    with one activation type, else None."""
    nn = torch.nn
    acts = {nn.ELU: ACT_ELU, nn.SiLU: ACT_SILU, nn.ReLU: ACT_RELU}
    mods = list(seq)
    if len(mods) < 4 or (len(mods) - 1) % 3 != 0 or not isinstance(mods[-1], nn.Linear):
        return None
    blocks, act = [], None
    for i in range(0, len(mods) - 1, 3):
        lin, ln, a = mods[i], mods[i + 1], mods[i + 2]
        if not (isinstance(lin, nn.Linear) and lin.bias is not None and isinstance(ln, nn.LayerNorm)
                and ln.elementwise_affine and ln.bias is not None and abs(ln.eps - 1e-5) < 1e-12
                and len(ln.normalized_shape) == 1 and type(a) in acts):
            return None
        if isinstance(a, nn.ELU) and a.alpha != 1.0:
            return None
        if act is not None and acts[type(a)] != act:
            return None
        act = acts[type(a)]
        blocks.append((lin, ln))
    if mods[-1].bias is None:
        return None
    hidden = blocks[0][0].out_features
    if any(b[0].out_features != hidden for b in blocks) or mods[-1].in_features != hidden:
        return None
    return blocks, act, mods[-1]


def _mlp_dims(rows: int, in_dim: int, hidden: int, depth: int, out_dim: int, act: int, flags: int = 0) -> _lib.MlpDims:
    return _lib.MlpDims(rows=rows, in_dim=in_dim, hidden=hidden, depth=depth, out_dim=out_dim, act=act, flags=flags)


def _mlp_ptrs(params: Sequence[Optional[torch.Tensor]], depth: int) -> _lib.MlpPtrs:
    """flat order: per block (lin_w, lin_b, ln_g, ln_b), then out_w, out_b (the actor's layout)."""
    s = _lib.MlpPtrs()
    for l in range(depth):
        s.lin_w[l] = _lib.ptr(params[4 * l + 0])
        s.lin_b[l] = _lib.ptr(params[4 * l + 1])
        s.ln_g[l] = _lib.ptr(params[4 * l + 2])
        s.ln_b[l] = _lib.ptr(params[4 * l + 3])
    s.out_w = _lib.ptr(params[4 * depth])
    s.out_b = _lib.ptr(params[4 * depth + 1])
    return s


class _PanelCache:
    """The bf16 MMA panel of the most recent feature tensor: reward model, done model, critic and target critic all
    read the same ``features`` memory (world_model.py:178-180, value_trainer.py:121-123), which is converted once.
    The cache keeps a (detached) view of the source alive, so its address cannot be handed to another tensor while
    the entry exists; (address, shape, version counter) then identifies the contents - ``x.detach()`` and views of
    ``x`` hit, an in-place write misses."""

    def __init__(self):
        self.key = None
        self.src: Optional[torch.Tensor] = None
        self.panel: Optional[torch.Tensor] = None

    def clear(self) -> None:
        self.key = self.src = self.panel = None

    def get(self, x2: torch.Tensor, d: _lib.MlpDims) -> torch.Tensor:
        key = (x2.data_ptr(), tuple(x2.shape), x2._version, x2.device)
        if self.key == key and self.panel is not None:
            return self.panel
        lib = _lib.load()
        dev = x2.device
        panel = _bytes(lib.wmrl_mlp_panel_bytes(C.byref(d)), dev)
        _native(dev, lib.wmrl_mlp_to_panel, C.byref(d), x2.data_ptr(), panel.data_ptr(), _stream_ptr(dev))
        self.key, self.src, self.panel = key, x2, panel
        return panel


_PANELS = _PanelCache()


def clear_panel_cache() -> None:
    """Drop the cached feature panel (and the reference that keeps its source tensor alive)."""
    _PANELS.clear()

_MLP_PACKED: "dict" = {}


def _mlp_packed(seq, params, d0: _lib.MlpDims) -> torch.Tensor:
    """packed bf16 image of one head's weights, refreshed when a parameter's version changes (an optimiser step)."""
    import weakref
    lib = _lib.load()
    key = tuple((p.data_ptr(), p._version) for p in params)
    ent = _MLP_PACKED.get(id(seq))
    if ent is not None and ent[0]() is seq and ent[1] == key:
        return ent[2]
    dev = params[0].device
    buf = _bytes(lib.wmrl_mlp_packed_bytes(C.byref(d0)), dev)
    p32 = [_f32c(p) for p in params]
    ptrs = _mlp_ptrs(p32, d0.depth)
    _native(dev, lib.wmrl_mlp_pack, C.byref(d0), C.byref(ptrs), buf.data_ptr(), _stream_ptr(dev))
    if len(_MLP_PACKED) > 64:
        for k in [k for k, v in _MLP_PACKED.items() if v[0]() is None]:
            del _MLP_PACKED[k]
    _MLP_PACKED[id(seq)] = (weakref.ref(seq), key, buf)
    return buf


class MlpBF16(torch.autograd.Function):
    """[Linear, LayerNorm, act] x depth -> Linear on the layer-wise tcgen05 chain (wmrl_mlp_fwd_bf16 /
    wmrl_mlp_bwd_bf16): bf16 operands, fp32 accumulation, fp32 in / out.  Gradients: d x, and the parameters'
    when any of them requires grad."""

    @staticmethod
    def forward(ctx, spec, packed, panel, x, *params):
        lib = _lib.load()
        in_dim, hidden, depth, out_dim, act = spec
        dev = x.device
        rows = x.numel() // in_dim
        need = x.requires_grad or any(p.requires_grad for p in params)
        d = _mlp_dims(rows, in_dim, hidden, depth, out_dim, act, _lib.MLP_SAVE if need else 0)
        out = torch.empty(rows, out_dim, device=dev)
        saved = _bytes(lib.wmrl_mlp_saved_bytes(C.byref(d)), dev) if need else None
        ws_n = lib.wmrl_mlp_workspace_bytes(C.byref(d), 0)
        ws = _bytes(ws_n, dev)
        if rows > 0:
            _native(dev, lib.wmrl_mlp_fwd_bf16, C.byref(d), packed.data_ptr(), panel.data_ptr(), out.data_ptr(),
                    _lib.ptr(saved), 0 if saved is None else saved.numel(), ws.data_ptr(), ws.numel(), _stream_ptr(dev))
        ctx.spec, ctx.rows, ctx.x_shape = spec, rows, x.shape
        ctx.packed, ctx.panel, ctx.saved_buf = packed, panel, saved
        ctx.param_shapes = [p.shape for p in params]
        return out.reshape(*x.shape[:-1], out_dim)

    @staticmethod
    def backward(ctx, g_out):
        lib = _lib.load()
        in_dim, hidden, depth, out_dim, act = ctx.spec
        rows = ctx.rows
        dev = g_out.device
        need_x = ctx.needs_input_grad[3]
        need_w = any(ctx.needs_input_grad[4:])
        g = _f32c(g_out.reshape(rows, out_dim))
        g_x = torch.empty(rows, in_dim, device=dev) if need_x else None
        grads = [torch.zeros(s, device=dev) for s in ctx.param_shapes] if need_w else None
        if rows > 0:
            d = _mlp_dims(rows, in_dim, hidden, depth, out_dim, act, _lib.MLP_WGRAD if need_w else 0)
            ws = _bytes(lib.wmrl_mlp_workspace_bytes(C.byref(d), 1), dev)
            gp = _mlp_ptrs(grads, depth) if need_w else _lib.MlpPtrs()
            _native(dev, lib.wmrl_mlp_bwd_bf16, C.byref(d), ctx.packed.data_ptr(), ctx.panel.data_ptr(),
                    ctx.saved_buf.data_ptr(), g.data_ptr(), C.byref(gp), _lib.ptr(g_x), ws.data_ptr(), ws.numel(),
                    _stream_ptr(dev))
        elif need_x:
            g_x.zero_()
        out = [None, None, None, g_x.reshape(ctx.x_shape) if need_x else None]
        for i in range(len(ctx.param_shapes)):
            out.append(grads[i] if need_w and ctx.needs_input_grad[4 + i] else None)
        return tuple(out)


def mlp_bf16_supported(seq: "torch.nn.Sequential") -> bool:
    pat = _mlp_pattern(seq)
    if pat is None:
        return False
    blocks, act, out = pat
    d = _mlp_dims(0, blocks[0][0].in_features, blocks[0][0].out_features, len(blocks), out.out_features, act)
    return _lib.load().wmrl_mlp_packed_bytes(C.byref(d)) > 0


def mlp_forward_bf16(seq: "torch.nn.Sequential", x: torch.Tensor) -> torch.Tensor:
    """The head ``seq`` applied to ``x[..., in_dim]`` on the bf16 tcgen05 chain.  Raises for module patterns or
    widths the chain does not cover (``mlp_bf16_supported``); the fp32 route is ``mlp_forward``."""
    pat = _mlp_pattern(seq)
    if pat is None or not mlp_bf16_supported(seq):
        raise RuntimeError("wmrl: this module is not a [Linear, LayerNorm, ELU|SiLU|ReLU] x depth -> Linear head with "
                           "hidden % 32 == 0, hidden <= 512, out_dim <= 16; use precision='fp32'")
    blocks, act, out = pat
    params = []
    for lin, ln in blocks:
        params += [lin.weight, lin.bias, ln.weight, ln.bias]
    params += [out.weight, out.bias]
    dev = _require_cuda(x, *params)
    in_dim, hidden = blocks[0][0].in_features, blocks[0][0].out_features
    if x.shape[-1] != in_dim:
        raise ValueError(f"mlp: input has {x.shape[-1]} features, the head expects {in_dim}")
    spec = (in_dim, hidden, len(blocks), out.out_features, act)
    d0 = _mlp_dims(x.numel() // in_dim, *spec)
    packed = _mlp_packed(seq, params, d0)
    x2 = _f32c(x.reshape(-1, in_dim))
    panel = _PANELS.get(x2, d0)
    return MlpBF16.apply(spec, packed, panel, x, *params)


def mlp_forward(seq: "torch.nn.Sequential", x: torch.Tensor) -> torch.Tensor:
    """Run an ``nn.Sequential`` of the reference's head pattern - [Linear, LayerNorm, act] x depth, Linear -
    on libwmrl's kernels (3xTF32 tensor-core GEMM + fused LayerNorm/activation) when ``x`` is a CUDA fp32
    tensor; any other layer, dtype or device runs the torch module itself.  Parameters stay in the module,
    so checkpoints and optimisers are untouched."""
    nn = torch.nn
    acts = {nn.ELU: ACT_ELU, nn.SiLU: ACT_SILU, nn.ReLU: ACT_RELU}
    mods = list(seq)
    native = x.is_cuda and x.dtype == torch.float32
    i = 0
    while i < len(mods):
        m = mods[i]
        if native and isinstance(m, nn.Linear):
            x = linear(x, m.weight, m.bias, 0)
            nxt, nxt2 = (mods[i + 1] if i + 1 < len(mods) else None), (mods[i + 2] if i + 2 < len(mods) else None)
            if (isinstance(nxt, nn.LayerNorm) and nxt.elementwise_affine and abs(nxt.eps - 1e-5) < 1e-12
                    and len(nxt.normalized_shape) == 1 and type(nxt2) in acts
                    and not (isinstance(nxt2, nn.ELU) and nxt2.alpha != 1.0)):
                x = layer_norm_act(x, nxt.weight, nxt.bias, acts[type(nxt2)])
                i += 3
                continue
            i += 1
            continue
        x = m(x)
        i += 1
    return x
