"""Builders shared by the GPU parity tests: wmrl modules loaded with golden /
seeded weights, and the oracle run on the same tensors."""
import torch

import wmrl
from oracle import rssm_oracle as O


def build_modules(variant, D, S, C, Hd, E, A, Ha, La, bound=1.0, seed=0, device="cuda", precision="fp32"):
    torch.manual_seed(seed)
    if variant == "continuous":
        rssm = wmrl.ContinuousRSSM(Hd, D, S, E, A, precision=precision)
        S_in = S
    else:
        rssm = wmrl.DiscreteRSSM(Hd, D, S, C, E, A, precision=precision)
        S_in = S * C
    actor = wmrl.Actor(D + S_in, A, bound, La, Ha)
    with torch.no_grad():   # non-trivial LayerNorm affine / biases
        for p in actor.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    return rssm.to(device), actor.to(device)


def modules_from_golden(g, device="cuda", precision="fp32"):
    bound = g.z["bound"].tolist()
    bound = bound if isinstance(bound, list) and len(bound) > 1 else float(g.z["bound"].reshape(-1)[0])
    rssm, actor = build_modules(g.variant, g.D, g.S, g.C, g.Hd, g.E, g.A, g.Ha, g.La, bound, device="cpu",
                                precision=precision)
    rssm.load_state_dict(g.weights("rssm."))
    actor.load_state_dict(g.weights("actor."))
    return rssm.to(device), actor.to(device)


def cpu_weights(module):
    return {k: v.detach().cpu().clone() for k, v in module.state_dict().items()}


def init_state(rssm, deter0, stoch0, device="cuda", seed=5):
    g = torch.Generator().manual_seed(seed)
    B = deter0.shape[0]
    if rssm.variant == 1:
        return wmrl.InternalStateDiscrete(deter_state=deter0.to(device), stoch_state=stoch0.to(device),
                                          logits=torch.randn(B, rssm.num_categories, rssm.stoch_size,
                                                             generator=g).to(device))
    return wmrl.InternalStateContinuous(deter_state=deter0.to(device), stoch_state=stoch0.to(device),
                                        mean=torch.randn(B, rssm.stoch_size, generator=g).to(device),
                                        std=torch.ones(B, rssm.stoch_size, device=device))


def synth_inputs(variant, B, H, D, S, C, seed=1):
    """SURVEY §8d synthetic start states and noise."""
    g = torch.Generator().manual_seed(seed)
    deter0 = torch.tanh(torch.randn(B, D, generator=g))
    if variant == "continuous":
        stoch0 = torch.randn(B, S, generator=g)
        noise = torch.randn(H, B, S, generator=g)
    else:
        idx = torch.randint(0, S, (B, C), generator=g)
        stoch0 = torch.nn.functional.one_hot(idx, S).float().reshape(B, C * S)
        noise = torch.empty(H, B, C, S).exponential_(1.0, generator=g)
    return deter0, stoch0, noise


def near_tie_mask(logits, q, tol=1e-4):
    """True where the oracle's argmax(p/q) is decided by a relative margin < tol
    (documented near-tie policy: those draws may legitimately flip)."""
    p = O.categorical_probs(logits)
    v = p / q
    top2 = v.topk(2, dim=-1).values
    return (top2[..., 0] - top2[..., 1]) / top2[..., 0] < tol
