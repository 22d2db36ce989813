"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules.

TEST INFRASTRUCTURE.  Runs only in the build container, where the reference is
mounted read-only at /root/reference (it does not exist on the GPU box, so the
vectors are committed).  Usage:

    python oracle/gen_golden.py            # rewrites tests/golden/*.npz

Noise injection: while a reference rollout runs, ``torch.randn_like``
is replaced by a reader of a pre-drawn ``eps`` tape (one entry per
``generate_stoch_state`` call, rssm.py:181) and ``torch.multinomial`` by
``argmax(p / q)`` on a pre-drawn Exp(1) tape - which is torch's own single-draw
algorithm; ``check_multinomial_equivalence`` proves that against the real
``torch.multinomial`` under a shared generator before anything is written.
"""
from __future__ import annotations

import contextlib
import os
import sys

import numpy as np
import torch

REF_SRC = "/root/reference/src"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")


def _import_reference():
    if not os.path.isdir(REF_SRC):
        raise SystemExit("reference not mounted at /root/reference - golden vectors are committed; nothing to do")
    sys.path.insert(0, REF_SRC)
    from reflect.components.rssm_world_model.rssm import ContinuousRSSM, DiscreteRSSM
    from reflect.components.models.actor import Actor
    from reflect.components.trainers.value.value_trainer import ValueGradTrainer
    from reflect.components.trainers.value.critic import ValueCritic
    from reflect.components.rssm_world_model.state import (  # noqa: F401
        InternalStateContinuous, InternalStateDiscrete)
    return dict(ContinuousRSSM=ContinuousRSSM, DiscreteRSSM=DiscreteRSSM, Actor=Actor,
                ValueGradTrainer=ValueGradTrainer, ValueCritic=ValueCritic,
                InternalStateContinuous=InternalStateContinuous,
                InternalStateDiscrete=InternalStateDiscrete)


def check_multinomial_equivalence():
    g = torch.Generator().manual_seed(123)
    p = torch.softmax(torch.randn(4096, 32, generator=g), -1)
    g1 = torch.Generator().manual_seed(7)
    ref = torch.multinomial(p, 1, True, generator=g1)
    g2 = torch.Generator().manual_seed(7)
    q = torch.empty_like(p).exponential_(1, generator=g2)
    mine = torch.argmax(p / q, dim=-1, keepdim=True)
    same = (ref == mine).float().mean().item()
    assert same == 1.0, f"multinomial != argmax(p/q): {same}"
    return same


@contextlib.contextmanager
def noise_tape(eps_list=None, q_list=None):
    """Patch torch.randn_like / torch.multinomial with tape readers."""
    eps_it = iter(eps_list or [])
    q_it = iter(q_list or [])
    orig_randn_like, orig_multinomial = torch.randn_like, torch.multinomial

    def randn_like(x, *a, **k):
        e = next(eps_it)
        assert e.shape == x.shape, (e.shape, x.shape)
        return e.to(x.dtype)

    def multinomial(p2d, n, replacement=False, *, generator=None):
        assert n == 1
        q = next(q_it).reshape(p2d.shape).to(p2d.dtype)
        return torch.argmax(p2d / q, dim=-1, keepdim=True)

    torch.randn_like, torch.multinomial = randn_like, multinomial
    try:
        yield
    finally:
        torch.randn_like, torch.multinomial = orig_randn_like, orig_multinomial


def _np(d, prefix=""):
    return {prefix + k: v.detach().cpu().numpy() for k, v in d.items()}


def make_case(ref, variant, name, *, B, H, T, D, S, C, Hd, E, A, Ha, La, seed):
    torch.manual_seed(seed)
    if variant == "continuous":
        rssm = ref["ContinuousRSSM"](Hd, D, S, E, A)
        S_in = S
    else:
        rssm = ref["DiscreteRSSM"](Hd, D, S, C, E, A)
        S_in = C * S
    bound = [0.5 + 0.25 * i for i in range(A)] if A > 1 else 1.0
    actor = ref["Actor"](D + S_in, A, bound, La, Ha)
    # make LayerNorm affine and biases non-trivial so they are actually checked
    with torch.no_grad():
        for n_, p_ in actor.named_parameters():
            if p_.dim() == 1:
                p_.add_(0.1 * torch.randn_like(p_))
    g = torch.Generator().manual_seed(seed + 1)
    deter0 = torch.tanh(torch.randn(B, D, generator=g))
    if variant == "continuous":
        stoch0 = torch.randn(B, S, generator=g)
        init = ref["InternalStateContinuous"](deter_state=deter0, stoch_state=stoch0,
                                              mean=torch.randn(B, S, generator=g), std=torch.ones(B, S))
        eps_im = torch.randn(H, B, S, generator=g)
        eps_ob = torch.randn(2 * (T - 1), B, S, generator=g)
        tape_im = dict(eps_list=list(eps_im))
        tape_ob = dict(eps_list=list(eps_ob))
        noise_im, noise_pri, noise_post = eps_im, eps_ob[0::2], eps_ob[1::2]
    else:
        idx0 = torch.randint(0, S, (B, C), generator=g)
        stoch0 = torch.nn.functional.one_hot(idx0, S).float().reshape(B, C * S)
        init = ref["InternalStateDiscrete"](deter_state=deter0, stoch_state=stoch0,
                                            logits=torch.randn(B, C, S, generator=g))
        q_im = torch.empty(H, B, C, S).exponential_(1, generator=g)
        q_ob = torch.empty(2 * (T - 1), B, C, S).exponential_(1, generator=g)
        tape_im = dict(q_list=list(q_im))
        tape_ob = dict(q_list=list(q_ob))
        noise_im, noise_pri, noise_post = q_im, q_ob[0::2], q_ob[1::2]

    out = {}
    out.update(_np(rssm.state_dict(), "rssm."))
    out.update(_np(actor.state_dict(), "actor."))
    out["bound"] = np.asarray(bound, dtype=np.float32)
    out["deter0"], out["stoch0"] = deter0.numpy(), stoch0.numpy()
    out["noise_im"] = noise_im.numpy()
    out["noise_pri"], out["noise_post"] = noise_pri.numpy(), noise_post.numpy()
    out["dims"] = np.asarray([B, H, T, D, S, C, Hd, E, A, Ha, La], dtype=np.int64)

    # ---- imagine: rssm.py:123-140 with the WM frozen the way
    # WorldModel.imagine_rollout does it (world_model.py:172, utils.py:186-194)
    for p_ in rssm.parameters():
        p_.requires_grad_(False)
    captured = []
    orig_forward = actor.forward

    def spy(x, deterministic=False):
        a = orig_forward(x, deterministic=deterministic)
        captured.append(a)
        return a

    actor.forward = spy
    with noise_tape(**tape_im):
        seq = rssm.imagine_rollout(init, actor, H)
    actor.forward = orig_forward
    out["im_deter"] = seq.deter_states.detach().numpy()          # [B,H+1,D] incl. index 0
    out["im_stoch"] = seq.stoch_states.detach().numpy()
    out["im_actions"] = torch.stack(captured, 1).detach().numpy()
    if variant == "continuous":
        out["im_mean"] = seq.means.detach().numpy()
        out["im_std"] = seq.stds.detach().numpy()
    else:
        out["im_logits"] = seq.logits.detach().numpy()
    # actor-BPTT gradients of a fixed scalar of the features
    feats = seq.get_features()
    gw = torch.randn(feats.shape, generator=g) / feats.shape[0]
    out["im_gfeat"] = gw.numpy()
    loss = (feats * gw).sum()
    if variant == "continuous":
        gm = torch.randn(seq.means[:, 1:].shape, generator=g) / B
        gs = torch.randn(seq.stds[:, 1:].shape, generator=g) / B
        out["im_gmean"], out["im_gstd"] = gm.numpy(), gs.numpy()
        loss = loss + (seq.means[:, 1:] * gm).sum() + (seq.stds[:, 1:] * gs).sum()
    else:
        gl = torch.randn(seq.logits[:, 1:].shape, generator=g) / B
        out["im_glogits"] = gl.numpy()
        loss = loss + (seq.logits[:, 1:] * gl).sum()
    actor.zero_grad()
    loss.backward()
    for n_, p_ in actor.named_parameters():
        if p_.grad is not None:
            out["im_grad.actor." + n_] = p_.grad.numpy().copy()

    # ---- observe: rssm.py:93-121, full (wgrad + dgrad) backward
    for p_ in rssm.parameters():
        p_.requires_grad_(True)
    obs = torch.relu(torch.randn(B, T, E, generator=g)).requires_grad_(True)
    act = (torch.rand(B, T, A, generator=g) * 2 - 1).requires_grad_(True)
    out["ob_obs"], out["ob_act"] = obs.detach().numpy(), act.detach().numpy()
    torch.manual_seed(seed + 2)  # discrete initial logits are torch.randn (rssm.py:219), index 0 only
    with noise_tape(**tape_ob):
        pri, pos = rssm.observe_rollout(obs, act)
    loss = 0.0
    for tag, sq in (("pri", pri), ("pos", pos)):
        fields = dict(deter=sq.deter_states, stoch=sq.stoch_states)
        if variant == "continuous":
            fields.update(mean=sq.means, std=sq.stds)
        else:
            fields.update(logits=sq.logits)
        for k, v in fields.items():
            out[f"ob_{tag}_{k}"] = v.detach().numpy()
            gk = torch.randn(v[:, 1:].shape, generator=g) / B
            out[f"ob_g_{tag}_{k}"] = gk.numpy()
            loss = loss + (v[:, 1:] * gk).sum()
    rssm.zero_grad()
    loss.backward()
    for n_, p_ in rssm.named_parameters():
        out["ob_grad.rssm." + n_] = p_.grad.numpy().copy()
    out["ob_grad.obs"] = obs.grad.numpy().copy()
    out["ob_grad.act"] = act.grad.numpy().copy()

    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, f"{name}.npz")
    np.savez_compressed(path, **out)
    return path


def make_lambda_case(ref, name, *, B, H, seed, gamma=0.99, lam=0.95):
    g = torch.Generator().manual_seed(seed)
    F_ = 12
    feats = torch.randn(B, H, F_, generator=g)
    V = torch.randn(B, H, 1, generator=g).requires_grad_(True)
    r = torch.randn(B, H, 1, generator=g).requires_grad_(True)
    raw_d = torch.rand(B, H, 1, generator=g) * 0.58      # a few cross 0.5
    trainer = ref["ValueGradTrainer"](ref["Actor"](F_, 1, 1.0, 1, 8), ref["ValueCritic"](F_, 1, 8),
                                      gamma=gamma, lam=lam)
    # value_trainer.py:148-149
    d = torch.cat([torch.zeros_like(raw_d[:, [0]]), raw_d[:, :-1]], dim=1)
    d, _ = (d > 0.5).float().cummax(dim=1)
    targets = []
    for i in range(H):                                      # value_trainer.py:126-134
        targets.append(trainer.compute_value_target(
            target_state_values=V[:, i:, :], states=feats[:, i:, :],
            rewards=r[:, i:, :], dones=d[:, i:, :]))
    tv = torch.stack(targets, dim=1)
    gw = torch.randn(tv.shape, generator=g)
    (tv * gw).sum().backward()
    out = dict(V=V.detach().numpy(), r=r.detach().numpy(), raw_d=raw_d.numpy(), d=d.numpy(),
               target=tv.detach().numpy(), g_target=gw.numpy(), g_V=V.grad.numpy(), g_r=r.grad.numpy(),
               gamma=np.float64(gamma), lam=np.float64(lam))
    path = os.path.join(OUT, f"{name}.npz")
    np.savez_compressed(path, **out)
    return path


def make_heads_case(ref, name, *, B, H, F_, seed):
    """The MLP heads the rollout's callers run over the features: reference DenseModel (reward: SiLU, identity out;
    done: SiLU, sigmoid out; models.py:3-37) and ValueCritic (critic.py:4-29): outputs and every gradient."""
    from reflect.components.rssm_world_model.models import DenseModel
    torch.manual_seed(seed)
    heads = {"reward": DenseModel(depth=2, input_dim=F_, hidden_dim=64, output_dim=1),
             "done": DenseModel(depth=3, input_dim=F_, hidden_dim=32, output_dim=1, output_act=torch.nn.Sigmoid),
             "critic": ref["ValueCritic"](F_, 3, 96)}
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for m in heads.values():
            for p_ in m.parameters():
                if p_.dim() == 1:
                    p_.add_(0.2 * torch.randn(p_.shape, generator=g))
    out = {"dims": np.asarray([B, H, F_], dtype=np.int64)}
    x = torch.randn(B, H, F_, generator=g)
    out["x"] = x.numpy()
    for k, m in heads.items():
        out.update(_np(m.state_dict(), k + "."))
        xi = x.clone().requires_grad_(True)
        y = m(xi)
        gy = torch.randn(y.shape, generator=g)
        y.backward(gy)
        out[k + "_y"], out[k + "_gy"], out[k + "_gx"] = y.detach().numpy(), gy.numpy(), xi.grad.numpy()
        for n_, p_ in m.named_parameters():
            out[k + "_grad." + n_] = p_.grad.numpy().copy()
    path = os.path.join(OUT, f"{name}.npz")
    np.savez_compressed(path, **out)
    return path


def make_act_case(ref, variant, name, *, B, steps, D, S, C, Hd, E, A, Ha, La, seed):
    """WorldModelActor.__call__'s RSSM / actor stages (memory_actor.py:51-56) on the unmodified reference, `steps`
    chained environment steps: post = posterior(e, state); a = actor(post.features, deterministic=True);
    state = prior(a, post).  The latent draws come from the tape (posterior first, then prior, per step)."""
    torch.manual_seed(seed)
    if variant == "continuous":
        rssm = ref["ContinuousRSSM"](Hd, D, S, E, A)
    else:
        rssm = ref["DiscreteRSSM"](Hd, D, S, C, E, A)
    S_in = S if variant == "continuous" else C * S
    bound = [0.5 + 0.25 * i for i in range(A)] if A > 1 else 1.0
    actor = ref["Actor"](D + S_in, A, bound, La, Ha)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for p_ in actor.parameters():
            if p_.dim() == 1:
                p_.add_(0.1 * torch.randn(p_.shape, generator=g))
    deter0 = torch.tanh(torch.randn(B, D, generator=g))
    if variant == "continuous":
        state = ref["InternalStateContinuous"](deter_state=deter0, stoch_state=torch.zeros(B, S),
                                               mean=torch.zeros(B, S), std=torch.ones(B, S))
        noise = torch.randn(2 * steps, B, S, generator=g)
        tape = dict(eps_list=list(noise))
    else:
        state = ref["InternalStateDiscrete"](deter_state=deter0, stoch_state=torch.zeros(B, C * S),
                                             logits=torch.zeros(B, C, S))
        noise = torch.empty(2 * steps, B, C, S).exponential_(1, generator=g)
        tape = dict(q_list=list(noise))
    embeds = torch.randn(steps, B, E, generator=g)
    out = {"dims": np.asarray([B, steps, D, S, C, Hd, E, A, Ha, La], dtype=np.int64), "deter0": deter0.numpy(),
           "noise": noise.numpy(), "embeds": embeds.numpy(), "bound": np.asarray(bound, dtype=np.float32)}
    out.update(_np(rssm.state_dict(), "rssm."))
    out.update(_np(actor.state_dict(), "actor."))
    rec = {k: [] for k in ("action", "post_stoch", "post_a", "post_b", "deter", "pri_stoch", "pri_a", "pri_b")}
    with torch.no_grad(), noise_tape(**tape):
        for t in range(steps):
            post = rssm.posterior(embeds[t], state)
            a = actor(post.get_features(), deterministic=True)
            state = rssm.prior(a, post)
            rec["action"].append(a)
            rec["post_stoch"].append(post.stoch_state); rec["pri_stoch"].append(state.stoch_state)
            rec["deter"].append(state.deter_state)
            if variant == "continuous":
                rec["post_a"].append(post.mean); rec["post_b"].append(post.std)
                rec["pri_a"].append(state.mean); rec["pri_b"].append(state.std)
            else:
                rec["post_a"].append(post.logits); rec["pri_a"].append(state.logits)
    for k, v in rec.items():
        if v:
            out["step_" + k] = torch.stack(v).numpy()
    path = os.path.join(OUT, f"{name}.npz")
    np.savez_compressed(path, **out)
    return path


def main():
    ref = _import_reference()
    torch.set_num_threads(1)   # fixed summation order in the CPU GEMMs
    print("multinomial == argmax(p/q):", check_multinomial_equivalence())
    # reference fixture shapes (rssm_world_model/tests/conftest.py:55-84), small E/Ha to keep files tiny
    print(make_case(ref, "continuous", "cont_fixture", B=6, H=5, T=6, D=32, S=32, C=1, Hd=32, E=48, A=1,
                    Ha=64, La=3, seed=10))
    print(make_case(ref, "discrete", "disc_fixture", B=6, H=5, T=6, D=32, S=8, C=4, Hd=32, E=48, A=1,
                    Ha=64, La=3, seed=20))
    # ragged sizes: nothing a multiple of 8/16, Hd != D, A > 1, C != S
    print(make_case(ref, "continuous", "cont_ragged", B=5, H=4, T=5, D=27, S=7, C=1, Hd=19, E=21, A=3,
                    Ha=40, La=2, seed=30))
    print(make_case(ref, "discrete", "disc_ragged", B=5, H=4, T=5, D=27, S=6, C=3, Hd=19, E=21, A=3,
                    Ha=40, La=2, seed=40))
    print(make_lambda_case(ref, "lambda_h15", B=16, H=15, seed=50))
    print(make_lambda_case(ref, "lambda_h3", B=4, H=3, seed=51))
    print(make_lambda_case(ref, "lambda_h1", B=4, H=1, seed=52))
    print(make_lambda_case(ref, "lambda_h64", B=3, H=64, seed=53))
    print(make_heads_case(ref, "heads", B=7, H=5, F_=24, seed=60))
    print(make_act_case(ref, "continuous", "act_cont", B=3, steps=4, D=32, S=32, C=1, Hd=32, E=48, A=2, Ha=64, La=3, seed=70))
    print(make_act_case(ref, "discrete", "act_disc", B=3, steps=4, D=27, S=6, C=3, Hd=19, E=21, A=3, Ha=64, La=2, seed=71))


if __name__ == "__main__":
    main()
