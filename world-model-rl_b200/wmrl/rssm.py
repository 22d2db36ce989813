"""Drop-in RSSM modules: same constructor arguments, attributes, state_dict keys
and method names as ``reflect.components.rssm_world_model.rssm`` (rssm.py:13-227),
with ``imagine_rollout`` / ``observe_rollout`` executed by libwmrl kernels.

* ``imagine_rollout(initial_states, actor, n_steps)`` -> one ``wmrl_imagine_fwd``
  call (all H steps; the actor MLP, GRU, prior head and sampling fused in the
  native path) registered as an autograd Function whose backward is the
  hand-written BPTT ``wmrl_imagine_bwd``.
* ``observe_rollout(obs_embeds, action_embs)`` -> ``wmrl_observe_fwd`` / ``_bwd``.
* ``prior`` / ``posterior`` / ``observe_step`` stay eager torch: they are the
  B=1 online-acting API (memory_actor.py:34-57), not the hot path.

Extras over the reference signature are keyword-only: ``noise=`` (pre-drawn
tensors, see ``draw_imagine_noise``), ``generator=``, ``return_actions=``.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import nn

from . import _lib
from .functional import (ImagineRollout, ObserveRollout, ObserveRolloutBF16, PackedObserveWeights, PackedWeights,
                         PathConfig)
from .state import (InternalStateContinuous, InternalStateContinuousSequence, InternalStateDiscrete,
                    InternalStateDiscreteSequence)

_PRECISIONS = {"fp32": _lib.FP32, "bf16": _lib.BF16}


def _sequence_types(like=None):
    """(continuous Sequence class, discrete Sequence class) to RETURN.  The drop-in contract is the
    reference's own dataclasses: when the caller works with the reference package - the state handed in is an
    instance of ``reflect.components.rssm_world_model.state.*``, or that module is loaded in this process - the
    sequences come back as the reference's ``InternalState{Continuous,Discrete}Sequence`` (``isinstance`` checks and
    its methods keep working); otherwise as the mirrors in ``wmrl.state`` (same fields and methods)."""
    import sys
    ref = sys.modules.get("reflect.components.rssm_world_model.state")
    use_ref = ref is not None and (like is None or type(like).__module__.startswith("reflect."))
    if like is not None and type(like).__module__.startswith("reflect.") and ref is None:   # pragma: no cover
        import importlib
        ref = importlib.import_module("reflect.components.rssm_world_model.state")
        use_ref = True
    if use_ref:
        return ref.InternalStateContinuousSequence, ref.InternalStateDiscreteSequence
    return InternalStateContinuousSequence, InternalStateDiscreteSequence


def _actor_tensors(actor: nn.Module):
    """(La, Ha, flat tensor list in functional._actor_layout order) from an
    Actor-shaped module (reference Actor or wmrl.actor.Actor)."""
    try:
        La, layers, mu = int(actor.num_layers), actor.layers, actor.mu
    except AttributeError as e:  # pragma: no cover - defensive
        raise TypeError("imagine_rollout needs an Actor-shaped module (layers/mu/num_layers/bound)") from e
    # The native path hard-codes the block structure of actor.py:22-37 - [Linear, LayerNorm(eps 1e-5, affine),
    # ELU(alpha 1)] x num_layers and a Linear mu head.  Anything else would give silently different rollouts
    # where the reference simply calls actor(...), so it is rejected here.
    if len(layers) != 3 * La:
        raise TypeError(f"actor.layers must hold {3 * La} modules ([Linear, LayerNorm, ELU] x {La}), has {len(layers)}")
    Ha = int(actor.hidden_dim)
    flat = []
    for l in range(La):
        lin, ln, act = layers[3 * l], layers[3 * l + 1], layers[3 * l + 2]
        if not isinstance(lin, nn.Linear) or lin.bias is None or lin.out_features != Ha:
            raise TypeError(f"actor.layers[{3 * l}] must be nn.Linear(.., {Ha}) with a bias")
        if (not isinstance(ln, nn.LayerNorm) or not ln.elementwise_affine or ln.bias is None
                or tuple(ln.normalized_shape) != (Ha,) or abs(ln.eps - 1e-5) > 1e-12):
            raise TypeError(f"actor.layers[{3 * l + 1}] must be nn.LayerNorm({Ha}, eps=1e-5, elementwise_affine=True)")
        if not isinstance(act, nn.ELU) or act.alpha != 1.0:
            raise TypeError(f"actor.layers[{3 * l + 2}] must be nn.ELU(alpha=1.0)")
        flat += [lin.weight, lin.bias, ln.weight, ln.bias]
    if not isinstance(mu, nn.Linear) or mu.bias is None or mu.in_features != Ha:
        raise TypeError(f"actor.mu must be nn.Linear({Ha}, A) with a bias")
    flat += [mu.weight, mu.bias]
    return La, Ha, flat


def _actor_bound(actor: nn.Module, A: int, device) -> torch.Tensor:
    """``actor.bound`` (actor.py:20: a plain fp32 tensor attribute, scalar or one entry per action) as a
    contiguous [A] device tensor.  The expansion is cached on the actor keyed on the source tensor's identity
    and version, so repeated calls hand the same storage to the packed-weight cache (no repack, no per-call
    allocation) while an in-place edit of ``actor.bound`` is still picked up."""
    src = actor.bound if torch.is_tensor(actor.bound) else torch.as_tensor(actor.bound, dtype=torch.float32)
    key = (id(actor.bound), getattr(src, "_version", 0), src.data_ptr() if torch.is_tensor(actor.bound) else None,
           str(device), A)
    cached = actor.__dict__.get("_wmrl_bound_cache")
    if cached is not None and cached[0] == key:
        return cached[1]
    b = src.detach().to(device=device, dtype=torch.float32).reshape(-1)
    if b.numel() == 1:
        b = b.expand(A)
    if b.numel() != A:
        raise ValueError("actor.bound must be a scalar or have one entry per action dim")
    b = b.contiguous().clone()
    actor.__dict__["_wmrl_bound_cache"] = (key, b)
    return b


import operator as _operator
import weakref

_ACT_PLANS: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()   # RSSM module -> (actor ref, ActStepPlan, bound)

_RSSM_GETTER = _operator.attrgetter(*[_lib.RSSM_KEYS[f] for f in _lib.RSSM_FIELDS])


class RSSMBase(nn.Module):
    """Owns the weights of rssm.py:14-42 under identical names."""

    variant = _lib.CONTINUOUS

    def __init__(self, hidden_size: int, deter_size: int, stoch_size: int, obs_embed_size: int,
                 parameterized_stoch_size: int, state_action_embed_input_size: int,
                 precision: str = "fp32"):
        super().__init__()
        self.hidden_size = hidden_size
        self.deter_size = deter_size
        self.stoch_size = stoch_size
        self.obs_embed_size = obs_embed_size
        self.act = nn.ELU()
        self.rnn = nn.GRUCell(deter_size, deter_size)
        self.fc_state_action_embed = nn.Linear(state_action_embed_input_size, deter_size)
        self.state_prior = nn.Sequential(
            nn.Linear(deter_size, hidden_size), nn.ELU(), nn.Linear(hidden_size, parameterized_stoch_size))
        self.state_posterior = nn.Sequential(
            nn.Linear(deter_size + obs_embed_size, hidden_size), nn.ELU(),
            nn.Linear(hidden_size, parameterized_stoch_size))
        self.set_precision(precision)
        self._packed = PackedWeights()
        self._packed_obs = PackedObserveWeights()

    # ------------------------------------------------------------------ config
    def set_precision(self, precision: str) -> "RSSMBase":
        """'fp32' = SIMT validation mode, 'bf16' = tcgen05 tensor-core path."""
        if precision not in _PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}")
        self.precision = precision
        return self

    @property
    def num_groups(self) -> int:
        return getattr(self, "num_categories", 1)

    @property
    def flat_stoch_size(self) -> int:
        return self.stoch_size * (self.num_groups if self.variant == _lib.DISCRETE else 1)

    def _rssm_tensors(self):
        return list(_RSSM_GETTER(self))

    def _config(self, action_size: int, Ha: int = 0, La: int = 0, action_scale: float = 5.0) -> PathConfig:
        return PathConfig(variant=self.variant, D=self.deter_size, S=self.stoch_size, C=self.num_groups,
                          A=action_size, Hd=self.hidden_size, E=self.obs_embed_size, Ha=Ha, La=La,
                          action_scale=float(action_scale), precision=_PRECISIONS[self.precision],
                          record=torch.is_grad_enabled())

    # --------------------------------------------------- eager single-step API
    def initial_state_sequence(self, batch_size):
        raise NotImplementedError

    def initial_state(self, batch_size):
        raise NotImplementedError

    def generate_stoch_state(self, deter_state, dist: torch.Tensor):
        raise NotImplementedError

    def prior(self, action_emb: torch.Tensor, state):
        """h_t = f(s_{t-1}, a_{t-1}, h_{t-1}); p(s_t | h_t)   (rssm.py:53-68)"""
        x = self.act(self.fc_state_action_embed(torch.cat((state.stoch_state, action_emb), dim=-1)))
        deter = self.rnn(x, state.deter_state)
        return self.generate_stoch_state(deter, self.state_prior(deter))

    def posterior(self, obs_embed: torch.Tensor, state):
        """p(s_t | h_t, o_t)   (rssm.py:70-82)"""
        out = self.state_posterior(torch.cat((state.deter_state, obs_embed), dim=-1))
        return self.generate_stoch_state(state.deter_state, out)

    def observe_step(self, obs_embed, action_emb, state):
        prior_state = self.prior(action_emb, state)
        return prior_state, self.posterior(obs_embed, prior_state)

    def act_step(self, obs_embed: torch.Tensor, state, actor: nn.Module, *, noise_post: Optional[torch.Tensor] = None,
                 noise_prior: Optional[torch.Tensor] = None, generator: Optional[torch.Generator] = None,
                 deterministic: bool = True, noise_action: Optional[torch.Tensor] = None):
        """The online acting step of WorldModelActor.__call__ (memory_actor.py:51-56) as ONE native kernel:
        ``post = posterior(obs_embed, state)``, ``action = actor(post.get_features(), deterministic=True)``,
        ``next_state = prior(action, post)``  ->  (action, next_state, post).  fp32, no gradients; the two latent draws
        are made here (N(0,1) / Exp(1), as generate_stoch_state does) unless passed in.  ``deterministic=False`` samples
        the action in the kernel from the actor's Normal head (actor.py:54-59: clamp(mean + std * eps, +-bound)) and
        returns (action, next_state, post, (act_mean, act_std, entropy))."""
        from .functional import ActStepPlan
        ent = _ACT_PLANS.get(self)
        bver = getattr(actor.bound, "_version", 0)
        plan = ent[1] if (ent is not None and ent[0]() is actor and ent[2] is actor.bound and ent[3] == bver) else None
        if plan is None or not plan.valid():
            device = next(actor.parameters()).device
            La, Ha, actor_p = _actor_tensors(actor)
            A = actor.mu.weight.shape[0]
            cfg = self._config(A, Ha, La, getattr(actor, "action_scale", 5.0))
            plan = ActStepPlan(cfg, self._rssm_tensors(), actor_p, _actor_bound(actor, A, device))
            _ACT_PLANS[self] = (weakref.ref(actor), plan, actor.bound, bver)
        device, A = plan.dev, plan.cfg.A
        deter = state.deter_state
        if deter.device != device:
            deter = deter.to(device)
        if obs_embed.device != device:
            obs_embed = obs_embed.to(device)
        B = deter.shape[0]
        if noise_post is None or noise_prior is None:
            both = self.draw_imagine_noise(2, B, device, generator)     # one launch for both draws
            noise_post = both[0] if noise_post is None else noise_post
            noise_prior = both[1] if noise_prior is None else noise_prior
        sampling = None
        if not deterministic:
            if noise_action is None:
                noise_action = torch.randn(B, A, device=device, generator=generator)
            sampling = (actor.stddev.weight, actor.stddev.bias, noise_action)
        with torch.no_grad():
            out = plan.run(obs_embed, deter, noise_post, noise_prior, sampling)
        action, post_s, post_a, post_b, deter_n, pri_s, pri_a, pri_b = out[:8]
        if self.variant == _lib.DISCRETE:
            shp = (B, self.num_groups, self.stoch_size)
            post = InternalStateDiscrete(deter_state=deter, stoch_state=post_s, logits=post_a.reshape(shp))
            nxt = InternalStateDiscrete(deter_state=deter_n, stoch_state=pri_s, logits=pri_a.reshape(shp))
        else:
            post = InternalStateContinuous(deter_state=deter, stoch_state=post_s, mean=post_a, std=post_b)
            nxt = InternalStateContinuous(deter_state=deter_n, stoch_state=pri_s, mean=pri_a, std=pri_b)
        if not deterministic:
            return action, nxt, post, tuple(out[8:])
        return action, nxt, post

    # ------------------------------------------------------------ noise helpers
    def draw_imagine_noise(self, n_steps: int, batch: int, device, generator: Optional[torch.Generator] = None):
        """eps ~ N(0,1) [H,B,S] (continuous) or q ~ Exp(1) [H,B,C,S] (discrete):
        the draws the reference makes inside generate_stoch_state."""
        if self.variant == _lib.DISCRETE:
            q = torch.empty(n_steps, batch, self.num_groups, self.stoch_size, device=device)
            return q.exponential_(1.0, generator=generator)
        return torch.randn(n_steps, batch, self.stoch_size, device=device, generator=generator)

    # ------------------------------------------------------------- native paths
    def imagine_rollout(self, initial_states, actor: nn.Module, n_steps: int, *, noise: Optional[torch.Tensor] = None,
                        generator: Optional[torch.Generator] = None, return_actions: bool = False):
        """rssm.py:123-140.  Returns the reference's Sequence type with fields
        [B, n_steps+1, .] (index 0 = ``initial_states``)."""
        device = next(actor.parameters()).device
        La, Ha, actor_p = _actor_tensors(actor)
        A = actor.mu.weight.shape[0]
        cfg = self._config(A, Ha, La, getattr(actor, "action_scale", 5.0))
        deter0 = initial_states.deter_state.to(device)
        stoch0 = initial_states.stoch_state.to(device)
        B = deter0.shape[0]
        if deter0.shape[1] != self.deter_size or stoch0.shape[1] != self.flat_stoch_size:
            raise ValueError(f"initial state shapes {tuple(deter0.shape)}, {tuple(stoch0.shape)} do not match "
                             f"deter_size={self.deter_size}, flat stoch={self.flat_stoch_size}")
        if actor.layers[0].weight.shape[1] != self.deter_size + self.flat_stoch_size:
            raise ValueError("actor input_dim must equal deter_size + flat stoch size")
        if noise is None:
            noise = self.draw_imagine_noise(n_steps, B, device, generator)
        noise = noise.to(device)
        want = (n_steps, B, self.num_groups, self.stoch_size) if self.variant == _lib.DISCRETE else \
            (n_steps, B, self.stoch_size)
        if tuple(noise.shape) != want:
            raise ValueError(f"noise must have shape {want}, got {tuple(noise.shape)}")
        bound = _actor_bound(actor, A, device)
        rssm_p = self._rssm_tensors()
        packed = None
        if cfg.precision == _lib.BF16:
            packed = self._packed.get(cfg, rssm_p, actor_p, bound)
        if self.variant == _lib.DISCRETE:
            init_da, init_db = initial_states.logits.to(device), None
        else:
            init_da, init_db = initial_states.mean.to(device), initial_states.std.to(device)
        outs = ImagineRollout.apply(cfg, int(n_steps), packed, deter0, stoch0, init_da, init_db, noise, bound,
                                    *actor_p, *rssm_p)
        cont_cls, disc_cls = _sequence_types(initial_states)
        if self.variant == _lib.DISCRETE:
            deter_seq, stoch_seq, logits, actions, _idx = outs
            seq = disc_cls(deter_states=deter_seq, stoch_states=stoch_seq, logits=logits)
        else:
            deter_seq, stoch_seq, means, stds, actions = outs
            seq = cont_cls(deter_states=deter_seq, stoch_states=stoch_seq, means=means, stds=stds)
        return (seq, actions) if return_actions else seq

    def observe_rollout(self, obs_embeds: torch.Tensor, action_embs: torch.Tensor, *,
                        noise: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
                        generator: Optional[torch.Generator] = None):
        """rssm.py:93-121.  Returns (prior_sequence, posterior_sequence), fields
        [B, T, .]; index 0 is the zero initial state (discrete initial logits are
        drawn with torch.randn as in rssm.py:219)."""
        device = obs_embeds.device
        B, T = obs_embeds.shape[0], obs_embeds.shape[1]
        A = action_embs.shape[-1]
        cfg = self._config(A)
        steps = max(T - 1, 0)
        if noise is None:
            noise = (self.draw_imagine_noise(steps, B, device, generator),
                     self.draw_imagine_noise(steps, B, device, generator))
        n_pri, n_post = (n.to(device) for n in noise)
        init_pri = self.initial_state(B)
        init_pri.to(device)
        init_post = self.initial_state(B)
        init_post.to(device)
        if self.variant == _lib.DISCRETE:
            inits = (init_pri.logits, None, init_post.logits, None)
        else:
            inits = (init_pri.mean, init_pri.std, init_post.mean, init_post.std)
        rssm_p = self._rssm_tensors()
        packed = None
        if cfg.precision == _lib.BF16 and T > 1 and obs_embeds.is_cuda:
            # filtering AND its full backward (world-model training) on the bf16 layer-wise tcgen05 chain
            packed = self._packed_obs.get(cfg, rssm_p)
        if packed is not None:
            outs = ObserveRolloutBF16.apply(cfg, packed, obs_embeds, action_embs.to(device), n_pri, n_post, *inits, *rssm_p)
        else:
            outs = ObserveRollout.apply(cfg, obs_embeds, action_embs.to(device), n_pri, n_post, *inits, *rssm_p)
        cont_cls, disc_cls = _sequence_types()
        if self.variant == _lib.DISCRETE:
            pri = disc_cls(deter_states=outs[0], stoch_states=outs[1], logits=outs[2])
            post = disc_cls(deter_states=outs[3], stoch_states=outs[4], logits=outs[5])
        else:
            pri = cont_cls(deter_states=outs[0], stoch_states=outs[1], means=outs[2], stds=outs[3])
            post = cont_cls(deter_states=outs[4], stoch_states=outs[5], means=outs[6], stds=outs[7])
        return pri, post


class ContinuousRSSM(RSSMBase):
    """Gaussian latent (rssm.py:143-187)."""
    variant = _lib.CONTINUOUS

    def __init__(self, hidden_size: int, deter_size: int, stoch_size: int, obs_embed_size: int, action_size: int,
                 precision: str = "fp32"):
        super().__init__(hidden_size=hidden_size, deter_size=deter_size, stoch_size=stoch_size,
                         obs_embed_size=obs_embed_size, parameterized_stoch_size=2 * stoch_size,
                         state_action_embed_input_size=stoch_size + action_size, precision=precision)
        self.action_size = action_size

    def initial_state(self, batch_size):
        z = lambda n: torch.zeros(batch_size, n)  # noqa: E731
        return InternalStateContinuous(deter_state=z(self.deter_size), stoch_state=z(self.stoch_size),
                                       mean=z(self.stoch_size), std=z(self.stoch_size))

    def initial_state_sequence(self, batch_size):
        return InternalStateContinuousSequence.from_init(self.initial_state(batch_size))

    def generate_stoch_state(self, deter_state, dist: torch.Tensor):
        mean, raw = dist.split(self.stoch_size, dim=-1)
        std = nn.functional.softplus(raw) + 0.1
        return InternalStateContinuous(deter_state=deter_state, stoch_state=mean + std * torch.randn_like(mean),
                                       mean=mean, std=std)


class DiscreteRSSM(RSSMBase):
    """Categorical straight-through latent: ``num_categories`` independent groups,
    each a categorical over ``stoch_size`` classes (rssm.py:190-227; the axis
    naming follows the reference, rssm.py:223)."""
    variant = _lib.DISCRETE

    def __init__(self, hidden_size: int, deter_size: int, stoch_size: int, num_categories: int,
                 obs_embed_size: int, action_size: int, precision: str = "fp32"):
        super().__init__(hidden_size=hidden_size, deter_size=deter_size, stoch_size=stoch_size,
                         obs_embed_size=obs_embed_size, parameterized_stoch_size=num_categories * stoch_size,
                         state_action_embed_input_size=num_categories * stoch_size + action_size,
                         precision=precision)
        self.num_categories = num_categories
        self.action_size = action_size

    def initial_state(self, batch_size):
        return InternalStateDiscrete(
            deter_state=torch.zeros(batch_size, self.deter_size),
            stoch_state=torch.zeros(batch_size, self.stoch_size * self.num_categories),
            logits=torch.randn(batch_size, self.num_categories, self.stoch_size))

    def initial_state_sequence(self, batch_size):
        return InternalStateDiscreteSequence.from_init(self.initial_state(batch_size))

    def generate_stoch_state(self, deter_state, dist: torch.Tensor):
        logits = dist.reshape(-1, self.num_categories, self.stoch_size)
        return InternalStateDiscrete.from_logits(deter_state=deter_state, logits=logits)
