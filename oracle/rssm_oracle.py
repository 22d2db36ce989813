"""CPU oracle for the RSSM imagine / observe / lambda-return path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` may be imported by the
product package (``world-model-rl_b200/wmrl``).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs use it, and there only as the checker / the CPU arm.

What it is
----------
A functional restatement (explicit weight dicts, pre-drawn noise, no hidden RNG)
of the reference algorithm in ``/root/reference/src/reflect``; every function
cites the reference ``file:line`` it follows.  The arithmetic itself lives in
PyTorch ATen (third-party, not vendored by the reference and not pinned in its
``pyproject.toml:18-25``); the effective pin is this image's torch 2.11.0+cu128
CPU kernels, so the oracle calls the same ATen ops (``F.linear``,
``F.layer_norm``, ``F.elu``, ``F.softplus``, ``softmax``) in the same order.

Parity pin
----------
The reference's own tests hold no numeric vectors for this path (SURVEY.md §8c:
shapes / truthiness only), so the oracle is pinned against OUTPUTS OF THE
REFERENCE ITSELF: ``oracle/gen_golden.py`` imports the unmodified reference
modules in the build container, injects the noise by patching
``torch.randn_like`` / ``torch.multinomial``, and stores weights, inputs, noise,
outputs and actor gradients in ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py``
checks this file against those vectors (bit-for-bit on CPU where the op
sequence is identical).

Noise contract
------------------------------
continuous: ``eps[t]`` replaces ``torch.randn_like(mean)`` (``rssm.py:181``).
discrete:   ``q[t] ~ Exp(1)`` replaces the noise drawn inside
``torch.multinomial(p, 1, True)`` whose single-draw path is ``argmax(p / q)``.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Weights = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------
# bf16 operand emulation (checker for the tcgen05 path)
# --------------------------------------------------------------------------
def bf16_round(x: torch.Tensor) -> torch.Tensor:
    """round-to-nearest-even to bf16 and back: what a tcgen05 kind::f16 operand holds."""
    return x.to(torch.bfloat16).to(x.dtype)


def _lin(x, w, b, emu: bool):
    """F.linear; with ``emu`` both MMA operands (activation panel and weight) are rounded to bf16 and
    the products accumulate in fp32 - the arithmetic of the bf16 tcgen05 kernels (bias added in fp32
    by the epilogue)."""
    if emu:
        return F.linear(bf16_round(x), bf16_round(w), b)
    return F.linear(x, w, b)


# --------------------------------------------------------------------------
# MLP heads next to the path (rssm_world_model/models.py:3-37, trainers/value/critic.py:4-29)
# --------------------------------------------------------------------------
def mlp_head(sd: Weights, x: torch.Tensor, act: str = "silu", emu: bool = False) -> torch.Tensor:
    """[Linear, LayerNorm(eps 1e-5), act] x depth -> Linear over the last dim, from an ``nn.Sequential`` state dict
    ("0.weight", "1.weight", ... as DenseModel.fc / ValueCritic.layers hold them; models.py:10-22, critic.py:12-27).
    ``emu``: bf16 MMA operands (activations and weights), fp32 accumulation - the tcgen05 chain's arithmetic."""
    fn = {"elu": F.elu, "silu": F.silu, "relu": F.relu}[act]
    l = 0
    while f"{3 * l + 1}.weight" in sd:
        w, b = sd[f"{3 * l}.weight"], sd[f"{3 * l}.bias"]
        x = _lin(x, w, b, emu)
        x = F.layer_norm(x, (w.shape[0],), sd[f"{3 * l + 1}.weight"], sd[f"{3 * l + 1}.bias"], 1e-5)
        x = fn(x)
        l += 1
    return _lin(x, sd[f"{3 * l}.weight"], sd[f"{3 * l}.bias"], emu)


# --------------------------------------------------------------------------
# Actor  (components/models/actor.py)
# --------------------------------------------------------------------------
def actor_num_layers(aw: Weights) -> int:
    """Number of [Linear, LayerNorm, ELU] blocks (actor.py:22-37)."""
    n = 0
    while f"layers.{3 * n}.weight" in aw:
        n += 1
    return n


def actor_hidden(aw: Weights, x: torch.Tensor, emu: bool = False, bias0_bf16: bool = False) -> torch.Tensor:
    """[Linear, LayerNorm(eps=1e-5), ELU] x La  (actor.py:23-37, 48).  ``emu``: bf16 MMA operands;
    ``bias0_bf16``: the first layer's bias rides inside the GEMM as a bf16 weight row (folded-pair kernel)."""
    for l in range(actor_num_layers(aw)):
        w, b = aw[f"layers.{3 * l}.weight"], aw[f"layers.{3 * l}.bias"]
        g, be = aw[f"layers.{3 * l + 1}.weight"], aw[f"layers.{3 * l + 1}.bias"]
        if emu and bias0_bf16 and l == 0:
            b = bf16_round(b)
        x = _lin(x, w, b, emu)
        x = F.layer_norm(x, (w.shape[0],), g, be, 1e-5)
        x = F.elu(x)
    return x


def actor_mu(aw: Weights, x: torch.Tensor, action_scale: float, bound: torch.Tensor, emu: bool = False,
             bias0_bf16: bool = False) -> torch.Tensor:
    """deterministic action: tanh(mu / action_scale) * bound  (actor.py:48-52)."""
    hid = actor_hidden(aw, x, emu, bias0_bf16)
    mu = _lin(hid, aw["mu.weight"], aw["mu.bias"], emu)
    return torch.tanh(mu / action_scale) * bound


def actor_dist(aw: Weights, x: torch.Tensor, action_scale: float, bound: torch.Tensor):
    """(mu, std) of the Normal head: std = 1.0*sigmoid(.) + 0.01  (actor.py:54-59)."""
    hid = actor_hidden(aw, x)
    mu = torch.tanh(F.linear(hid, aw["mu.weight"], aw["mu.bias"]) / action_scale) * bound
    std = 1.0 * torch.sigmoid(F.linear(hid, aw["stddev.weight"], aw["stddev.bias"])) + 0.01
    return mu, std


# --------------------------------------------------------------------------
# Stochastic heads
# --------------------------------------------------------------------------
def gen_stoch_continuous(out: torch.Tensor, S: int, eps: torch.Tensor):
    """mean,raw = split; std = softplus(raw)+0.1; stoch = mean + std*eps (rssm.py:179-181)."""
    mean, raw = torch.split(out, S, dim=-1)
    std = F.softplus(raw) + 0.1
    stoch = mean + std * eps
    return stoch, mean, std


def categorical_probs(logits: torch.Tensor) -> torch.Tensor:
    """torch.distributions.Categorical(logits=l).probs: normalise by logsumexp,
    then softmax (state/discrete.py:27-29 -> OneHotCategoricalStraightThrough)."""
    ln = logits - logits.logsumexp(dim=-1, keepdim=True)
    return F.softmax(ln, dim=-1)


def gen_stoch_discrete(out: torch.Tensor, C: int, S: int, q: torch.Tensor, forced_idx: Optional[torch.Tensor] = None):
    """logits = out.reshape(B, num_categories, stoch_size) (rssm.py:223) - the
    categorical axis is the LAST one (size ``stoch_size``).  sample =
    one_hot(argmax(p/q)) + (p - p.detach())  (state/discrete.py:27-32,
    torch multinomial single-draw path)."""
    logits = out.reshape(-1, C, S)
    p = categorical_probs(logits)
    idx = torch.argmax(p.detach() / q.reshape(-1, C, S), dim=-1)
    if forced_idx is not None:      # replay of a recorded trajectory (test helper): idx stays the oracle's own draw
        onehot = F.one_hot(forced_idx.reshape(-1, C).long(), S).to(p.dtype)
        return (onehot + (p - p.detach())).reshape(-1, C * S), logits, idx
    onehot = F.one_hot(idx, S).to(p.dtype)
    stoch = onehot + (p - p.detach())
    return stoch.reshape(-1, C * S), logits, idx


# --------------------------------------------------------------------------
# RSSM single steps (rssm.py:53-91)
# --------------------------------------------------------------------------
def gru_cell(w: Weights, x: torch.Tensor, h: torch.Tensor, emu: bool = False) -> torch.Tensor:
    """torch.nn.GRUCell, gate order r,z,n; b_hn inside r*(.) (rssm.py:28,66).  ``emu``: the two GEMMs read
    bf16 operands; the convex update keeps the fp32 master copy of h."""
    gi = _lin(x, w["rnn.weight_ih"], w["rnn.bias_ih"], emu)
    gh = _lin(h, w["rnn.weight_hh"], w["rnn.bias_hh"], emu)
    i_r, i_z, i_n = gi.chunk(3, -1)
    h_r, h_z, h_n = gh.chunk(3, -1)
    r = torch.sigmoid(i_r + h_r)
    z = torch.sigmoid(i_z + h_z)
    n = torch.tanh(i_n + r * h_n)
    return (1.0 - z) * n + z * h


def prior_deter(w: Weights, stoch: torch.Tensor, action: torch.Tensor, deter: torch.Tensor, emu: bool = False):
    """cat(stoch, a) -> Linear -> ELU -> GRUCell -> prior head (rssm.py:64-67)."""
    sa = torch.cat([stoch, action], dim=-1)
    x = F.elu(_lin(sa, w["fc_state_action_embed.weight"], w["fc_state_action_embed.bias"], emu))
    h = gru_cell(w, x, deter, emu)
    hid = F.elu(_lin(h, w["state_prior.0.weight"], w["state_prior.0.bias"], emu))
    out = _lin(hid, w["state_prior.2.weight"], w["state_prior.2.bias"], emu)
    return h, out


def posterior_out(w: Weights, deter: torch.Tensor, obs_embed: torch.Tensor, emu: bool = False) -> torch.Tensor:
    """cat(deter, obs_embed) -> Linear -> ELU -> Linear (rssm.py:80-81)."""
    hcat = torch.cat([deter, obs_embed], dim=-1)
    hid = F.elu(_lin(hcat, w["state_posterior.0.weight"], w["state_posterior.0.bias"], emu))
    return _lin(hid, w["state_posterior.2.weight"], w["state_posterior.2.bias"], emu)


def act_step(variant: str, w: Weights, aw: Weights, obs_embed: torch.Tensor, deter: torch.Tensor, noise_post: torch.Tensor,
             noise_prior: torch.Tensor, S: int, C: int = 1, action_scale: float = 5.0, bound=None):
    """One online acting step of WorldModelActor.__call__ (memory_actor.py:51-56) with the two latent draws passed
    in: post = posterior(e, state) (rssm.py:70-82), a = actor(post.features, deterministic) (actor.py:47-52),
    next = prior(a, post) (rssm.py:53-68).  -> dict of action, post_* and next-state tensors."""
    bound = torch.ones(1) if bound is None else bound
    gen = (lambda o, n: gen_stoch_continuous(o, S, n)) if variant == "continuous" \
        else (lambda o, n: gen_stoch_discrete(o, C, S, n))
    out_q = posterior_out(w, deter, obs_embed)
    post_s, post_a, post_b = gen(out_q, noise_post)
    action = actor_mu(aw, torch.cat([deter, post_s], dim=-1), action_scale, bound)
    h, out_p = prior_deter(w, post_s, action, deter)
    pri_s, pri_a, pri_b = gen(out_p, noise_prior)
    return {"action": action, "post_stoch": post_s, "post_a": post_a, "post_b": post_b, "deter_next": h,
            "prior_stoch": pri_s, "prior_a": pri_a, "prior_b": pri_b}


# --------------------------------------------------------------------------
# Rollouts
# --------------------------------------------------------------------------
def imagine_rollout(
    variant: str,
    w: Weights,
    aw: Weights,
    deter0: torch.Tensor,
    stoch0: torch.Tensor,
    noise: torch.Tensor,
    n_steps: int,
    S: int,
    C: int = 1,
    action_scale: float = 5.0,
    bound: Optional[torch.Tensor] = None,
    emulate_bf16: bool = False,
    bias0_bf16: bool = True,
    forced_idx: Optional[torch.Tensor] = None,
):
    """RSSMBase.imagine_rollout (rssm.py:123-140).

    ``forced_idx`` [B,H,C] (discrete): the one-hot sample fed forward is taken from these indices instead
    of the oracle's own argmax (which is still returned in ``idx``) - used to replay a long recorded
    trajectory so that a single near-tie flip does not turn a gradient comparison into chaos.

    ``emulate_bf16``: same op order, but every GEMM reads its activation panel and its weight rounded to
    bf16 and accumulates in fp32 - the rounding points of the bf16 tcgen05 kernels (state, latent, action,
    hidden activations; fp32 master copy of h in the GRU update; fp32 biases except, with ``bias0_bf16``,
    the first actor layer's, which the folded-pair kernel carries inside the GEMM).  It pins FREE-RUNNING
    multi-step parity of that path: what is left is accumulation order and tanh.approx / ex2.approx.

    Per step: features.detach() -> actor(deterministic=True) -> prior.  Returns
    per-step lists stacked on dim 1 WITHOUT the index-0 slot (the caller owns
    the initial state): deter[B,H,D], stoch[B,H,S_in], dist (mean,std) or
    logits[B,H,C,S], actions[B,H,A], idx[B,H,C] (discrete only).
    noise: continuous eps[H,B,S]; discrete q[H,B,C,S].
    """
    if bound is None:
        bound = torch.tensor(1.0, dtype=deter0.dtype)
    deter, stoch = deter0, stoch0
    o_deter, o_stoch, o_a, o_d1, o_d2, o_idx = [], [], [], [], [], []
    for t in range(n_steps):
        feat = torch.cat([deter, stoch], dim=-1).detach()          # rssm.py:135
        action = actor_mu(aw, feat, action_scale, bound, emulate_bf16, bias0_bf16)   # rssm.py:136
        deter, out = prior_deter(w, stoch, action, deter, emulate_bf16)              # rssm.py:137
        if variant == "continuous":
            stoch, mean, std = gen_stoch_continuous(out, S, noise[t])
            o_d1.append(mean)
            o_d2.append(std)
        else:
            stoch, logits, idx = gen_stoch_discrete(out, C, S, noise[t],
                                                    None if forced_idx is None else forced_idx[:, t])
            o_d1.append(logits)
            o_idx.append(idx)
        o_deter.append(deter)
        o_stoch.append(stoch)
        o_a.append(action)
    res = {
        "deter": torch.stack(o_deter, 1),
        "stoch": torch.stack(o_stoch, 1),
        "actions": torch.stack(o_a, 1),
    }
    if variant == "continuous":
        res["mean"] = torch.stack(o_d1, 1)
        res["std"] = torch.stack(o_d2, 1)
    else:
        res["logits"] = torch.stack(o_d1, 1)
        res["idx"] = torch.stack(o_idx, 1)
    return res


def observe_rollout(
    variant: str,
    w: Weights,
    obs_embeds: torch.Tensor,
    actions: torch.Tensor,
    noise_prior: torch.Tensor,
    noise_post: torch.Tensor,
    S: int,
    C: int = 1,
    emulate_bf16: bool = False,
    forced_post_idx: Optional[torch.Tensor] = None,
):
    """RSSMBase.observe_rollout (rssm.py:93-121) from the zero initial state
    (rssm.py:166-172 / 215-220).  For t in 0..T-2: prior_{t+1} = prior(a_t, post_t);
    post_{t+1} = posterior(e_{t+1}, prior_{t+1}).  Two draws per step (prior
    first, then posterior).  Returns dicts for the prior and posterior
    sequences WITHOUT the index-0 slot, fields [B,T-1,...].

    ``emulate_bf16`` / ``forced_post_idx`` [B,T-1,C]: checker modes for the bf16 tcgen05 chain, as in
    ``imagine_rollout`` (GEMM operands rounded to bf16; the posterior sample that is fed forward replays
    recorded indices while ``idx`` still reports the oracle's own draw)."""
    Bn, T = obs_embeds.shape[:2]
    D = w["rnn.weight_hh"].shape[1]
    S_in = S if variant == "continuous" else C * S
    deter = torch.zeros(Bn, D, dtype=obs_embeds.dtype)
    stoch = torch.zeros(Bn, S_in, dtype=obs_embeds.dtype)
    pri: Dict[str, List[torch.Tensor]] = {k: [] for k in ("deter", "stoch", "d1", "d2", "idx")}
    pos: Dict[str, List[torch.Tensor]] = {k: [] for k in ("deter", "stoch", "d1", "d2", "idx")}
    for t in range(T - 1):
        h, out = prior_deter(w, stoch, actions[:, t], deter, emulate_bf16)       # rssm.py:90 via :112
        pout = posterior_out(w, h, obs_embeds[:, t + 1], emulate_bf16)           # rssm.py:91
        if variant == "continuous":
            ps, pm, pstd = gen_stoch_continuous(out, S, noise_prior[t])
            qs, qm, qstd = gen_stoch_continuous(pout, S, noise_post[t])
            pri["d1"].append(pm); pri["d2"].append(pstd)
            pos["d1"].append(qm); pos["d2"].append(qstd)
        else:
            ps, pl, pi = gen_stoch_discrete(out, C, S, noise_prior[t])
            qs, ql, qi = gen_stoch_discrete(pout, C, S, noise_post[t],
                                            None if forced_post_idx is None else forced_post_idx[:, t])
            pri["d1"].append(pl); pri["idx"].append(pi)
            pos["d1"].append(ql); pos["idx"].append(qi)
        pri["deter"].append(h); pri["stoch"].append(ps)
        pos["deter"].append(h); pos["stoch"].append(qs)
        deter, stoch = h, qs                                        # rssm.py:120
    names = ("mean", "std") if variant == "continuous" else ("logits", None)

    def pack(d):
        r = {"deter": torch.stack(d["deter"], 1), "stoch": torch.stack(d["stoch"], 1),
             names[0]: torch.stack(d["d1"], 1)}
        if variant == "continuous":
            r["std"] = torch.stack(d["d2"], 1)
        else:
            r["idx"] = torch.stack(d["idx"], 1)
        return r

    return pack(pri), pack(pos)


# --------------------------------------------------------------------------
# lambda-return (trainers/value/value_trainer.py)
# --------------------------------------------------------------------------
def done_preprocess(dones: torch.Tensor) -> torch.Tensor:
    """shift right by one, threshold at 0.5, running max (value_trainer.py:148-149)."""
    d = torch.cat([torch.zeros_like(dones[:, [0]]), dones[:, :-1]], dim=1)
    d, _ = (d > 0.5).float().cummax(dim=1)
    return d


def _rollout_value(V, r, d, gam, k):
    """compute_rollout_value (value_trainer.py:62-86) on a slice."""
    big_h = V.shape[1]
    h = min(k, big_h - 1)
    R = gam[:h][None, :, None] * r[:, :h] * (1 - d[:, :h])
    final = gam[h] * V[:, [h]] * (1 - d[:, [h]])
    return R.sum(1) + final[:, 0, :]


def lambda_return_loops(V, r, d, gamma: float = 0.99, lam: float = 0.95):
    """value_loss's target loop (value_trainer.py:126-134) over
    compute_value_target (:88-113).  gamma table is built once, as an fp32 tensor
    of length H (:71-74).  V,r,d: [B,H,1] -> targets [B,H,1]."""
    H = V.shape[1]
    gam = torch.tensor([gamma ** i for i in range(H)]).to(V.dtype)
    outs = []
    for i in range(H):
        Vi, ri, di = V[:, i:], r[:, i:], d[:, i:]
        big_h = Vi.shape[1]
        val_sum = 0
        for k in range(1, big_h - 1):
            val_sum = val_sum + lam ** (k - 1) * _rollout_value(Vi, ri, di, gam, k)
        final_val = _rollout_value(Vi, ri, di, gam, big_h)
        outs.append((1 - lam) * val_sum + lam ** (big_h - 1) * final_val)
    return torch.stack(outs, dim=1)


def lambda_return_scan(V, r, d, gamma: float = 0.99, lam: float = 0.95):
    """O(H) restatement of the loops above, exact in real
    arithmetic: r'=r(1-d), V'=V(1-d); Std/M/T at H-1 = V'_{H-1}; for i<H-1:
    Std_i = r'_i + g((1-l)V'_{i+1} + l Std_{i+1}); M_i = r'_i + g M_{i+1};
    T_i = Std_i - (l^{n-2} - l^{n-1}) M_i, n = H-i (the reference's weights do
    not sum to one; n=2 gives l*G_1)."""
    H = V.shape[1]
    rp, Vp = r * (1 - d), V * (1 - d)
    out = [None] * H
    Std = Vp[:, H - 1]
    M = Vp[:, H - 1]
    out[H - 1] = Vp[:, H - 1]
    for i in range(H - 2, -1, -1):
        n = H - i
        Std = rp[:, i] + gamma * ((1 - lam) * Vp[:, i + 1] + lam * Std)
        M = rp[:, i] + gamma * M
        out[i] = Std - (lam ** (n - 2) - lam ** (n - 1)) * M
    return torch.stack(out, dim=1)


def imagine_rollout_reference_like(variant, w, aw, deter0, stoch0, n_steps, S, C=1, action_scale=5.0, bound=None):
    """CPU-baseline twin of ``imagine_rollout``: the same arithmetic, but executed the
    way the reference's CPU path executes it - noise drawn internally
    (``torch.randn_like`` rssm.py:181 / ``OneHotCategoricalStraightThrough.rsample``
    state/discrete.py:27-30) and the ``[B,t,.]`` buffers re-grown with ``torch.cat``
    every step (``Sequence.append_``, state/continuous.py:56-60, state/discrete.py:56-59).
    Used only for timing (bench.py cpu_baseline / --impl reference)."""
    import torch.distributions as Dist
    if bound is None:
        bound = torch.tensor(1.0, dtype=deter0.dtype)
    deter, stoch = deter0, stoch0
    seq = [deter0.unsqueeze(1), stoch0.unsqueeze(1)]
    extra = None
    for _ in range(n_steps):
        feat = torch.cat([deter, stoch], dim=-1).detach()
        action = actor_mu(aw, feat, action_scale, bound)
        deter, out = prior_deter(w, stoch, action, deter)
        if variant == "continuous":
            mean, raw = torch.split(out, S, dim=-1)
            std = F.softplus(raw) + 0.1
            stoch = mean + std * torch.randn_like(mean)
            new = [mean.unsqueeze(1), std.unsqueeze(1)]
        else:
            logits = out.reshape(-1, C, S)
            dist = Dist.Independent(Dist.OneHotCategoricalStraightThrough(logits=logits), 1)
            stoch = dist.rsample().reshape(-1, C * S)
            new = [logits.unsqueeze(1)]
        seq[0] = torch.cat([seq[0], deter.unsqueeze(1)], dim=1)
        seq[1] = torch.cat([seq[1], stoch.unsqueeze(1)], dim=1)
        extra = new if extra is None else [torch.cat([e, n_], dim=1) for e, n_ in zip(extra, new)]
    return seq + (extra or [])
