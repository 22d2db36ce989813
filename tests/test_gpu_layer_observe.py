"""GPU parity of posterior filtering (RSSMBase.observe_rollout forward, rssm.py:93-121) on the bf16 layer-wise
tcgen05 chain: `precision="bf16"` modules called without a gradient record.  Checker: the bf16-emulating oracle;
the discrete variant replays the kernel's posterior indices (they are what is fed forward), compares deter / both
logits sequences at an absolute 6e-3 and requires both index streams to agree with the oracle's own draws except
at near-ties (rate <= 1e-3)."""
import pytest
import torch

from oracle import rssm_oracle as O

pytestmark = pytest.mark.gpu
TOL = 6e-3


def _setup(variant, D, S, C, Hd, E, A, seed=5):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import helpers_gpu as HG
    from wmrl import _lib
    if not _lib.load().wmrl_device_supports_bf16():
        pytest.skip("not an sm_100 device")
    rssm, _ = HG.build_modules(variant, D, S, C, Hd, E, A, 64, 1, seed=seed, precision="bf16")
    return HG, rssm


def _inputs(variant, B, T, E, A, S, C, seed=2):
    g = torch.Generator().manual_seed(seed)
    obs = torch.relu(torch.randn(B, T, E, generator=g))          # the encoder ends in ReLU
    act = torch.rand(B, T, A, generator=g) * 2 - 1
    if variant == "continuous":
        n1, n2 = torch.randn(T - 1, B, S, generator=g), torch.randn(T - 1, B, S, generator=g)
    else:
        n1 = torch.empty(T - 1, B, C, S).exponential_(1.0, generator=g)
        n2 = torch.empty(T - 1, B, C, S).exponential_(1.0, generator=g)
    return obs, act, n1, n2


def _close(name, got, want, tol=TOL):
    err = float((got.detach().cpu().float() - want).abs().max())
    assert err <= tol, f"{name}: max |err| {err:.3e} > {tol}"
    return err


@pytest.mark.parametrize("shape", [(32, 8, 4, 32, 16, 2, 40, 9), (600, 32, 32, 600, 1024, 6, 50, 12), (200, 16, 8, 96, 64, 3, 150, 5)])
def test_discrete_filtering_replay_vs_bf16_oracle(shape):
    D, S, C, Hd, E, A, B, T = shape
    HG, rssm = _setup("discrete", D, S, C, Hd, E, A)
    obs, act, n1, n2 = _inputs("discrete", B, T, E, A, S, C)
    with torch.no_grad():
        pri, post = rssm.observe_rollout(obs.cuda(), act.cuda(), noise=(n1.cuda(), n2.cuda()))
    assert float(pri.deter_states[:, 0].abs().max()) == 0.0 and float(post.stoch_states[:, 0].abs().max()) == 0.0
    assert torch.equal(pri.deter_states, post.deter_states)
    st = post.stoch_states[:, 1:].reshape(B, T - 1, C, S).cpu()
    assert torch.all((st == 0) | (st == 1)) and torch.all(st.sum(-1) == 1)
    qi = st.argmax(-1)
    pi = pri.stoch_states[:, 1:].reshape(B, T - 1, C, S).cpu().argmax(-1)
    with torch.no_grad():
        rp, rq = O.observe_rollout("discrete", HG.cpu_weights(rssm), obs, act, n1, n2, S, C, emulate_bf16=True,
                                   forced_post_idx=qi)
    e = [_close("deter", post.deter_states[:, 1:], rq["deter"]), _close("prior logits", pri.logits[:, 1:], rp["logits"]),
         _close("posterior logits", post.logits[:, 1:], rq["logits"])]
    for name, mine, ref in (("prior", pi, rp["idx"]), ("posterior", qi, rq["idx"])):
        rate = float((mine != ref).float().mean())
        assert rate <= 1e-3, f"{name} index mismatch rate {rate:.2e}"
    print(shape, e)


@pytest.mark.parametrize("shape", [(200, 30, 200, 1024, 6, 50, 10), (32, 8, 32, 16, 1, 33, 6)])
def test_continuous_filtering_vs_bf16_oracle(shape):
    D, S, Hd, E, A, B, T = shape
    HG, rssm = _setup("continuous", D, S, 1, Hd, E, A)
    obs, act, n1, n2 = _inputs("continuous", B, T, E, A, S, 1)
    with torch.no_grad():
        pri, post = rssm.observe_rollout(obs.cuda(), act.cuda(), noise=(n1.cuda(), n2.cuda()))
        rp, rq = O.observe_rollout("continuous", HG.cpu_weights(rssm), obs, act, n1, n2, S, emulate_bf16=True)
    e = [_close("deter", post.deter_states[:, 1:], rq["deter"]), _close("prior mean", pri.means[:, 1:], rp["mean"]),
         _close("prior std", pri.stds[:, 1:], rp["std"]), _close("posterior mean", post.means[:, 1:], rq["mean"]),
         _close("posterior stoch", post.stoch_states[:, 1:], rq["stoch"]), _close("prior stoch", pri.stoch_states[:, 1:], rp["stoch"])]
    print(shape, e)


def _agree(name, got, ref, cos_min=0.999, rel_max=2e-2):
    got, ref = got.detach().float().cpu().flatten(), ref.detach().float().cpu().flatten()
    if got.numel() < 8:
        rel_max = max(rel_max, 6e-2)
    cos = float(got @ ref / (got.norm() * ref.norm() + 1e-30))
    rel = float((got - ref).norm() / (ref.norm() + 1e-30))
    assert cos >= cos_min and rel <= rel_max, f"{name}: cosine {cos:.6f}, relative error {rel:.3e}"
    return round(cos, 6), round(rel, 5)


@pytest.mark.parametrize("variant,shape", [("continuous", (200, 30, 1, 200, 1024, 6, 50, 10)), ("continuous", (48, 16, 1, 96, 32, 3, 140, 5)),
                                           ("discrete", (600, 32, 32, 600, 1024, 6, 50, 12)), ("discrete", (64, 8, 4, 64, 32, 2, 40, 7))])
def test_filtering_backward_vs_oracle_autograd(variant, shape):
    """world-model training through observe_rollout (world_model.py:120-163) on the bf16 chain: RSSM weight gradients,
    d obs_embeds and d actions against torch autograd on the fp32 oracle (discrete: posterior indices replayed).
    Tolerance for bf16 gradients: cosine >= 0.999, relative Frobenius error <= 2e-2 per tensor."""
    D, S, C, Hd, E, A, B, T = shape
    HG, rssm = _setup(variant, D, S, C, Hd, E, A)
    obs, act, n1, n2 = _inputs(variant, B, T, E, A, S, C)
    g = torch.Generator().manual_seed(11)
    S_in = S * C if variant == "discrete" else S
    gfeat = torch.randn(B, T - 1, D + S_in, generator=g) / B
    gpri = torch.randn(B, T - 1, D + S_in, generator=g) / B
    gd1 = torch.randn(B, T - 1, *( (C, S) if variant == "discrete" else (S,)), generator=g) / B
    gd2 = torch.randn(B, T - 1, *( (C, S) if variant == "discrete" else (S,)), generator=g) / B
    ob_d, ac_d = obs.cuda().requires_grad_(True), act.cuda().requires_grad_(True)
    pri, post = rssm.observe_rollout(ob_d, ac_d, noise=(n1.cuda(), n2.cuda()))
    dist_p = pri.logits if variant == "discrete" else pri.means
    dist_q = post.logits if variant == "discrete" else post.means
    loss = (post.get_features() * gfeat.cuda()).sum() + (pri.get_features() * gpri.cuda()).sum() + \
        (dist_p[:, 1:] * gd1.cuda()).sum() + (dist_q[:, 1:] * gd2.cuda()).sum()
    if variant == "continuous":
        loss = loss + (pri.stds[:, 1:] * gd2.cuda()).sum() + (post.stds[:, 1:] * gd1.cuda()).sum()
    loss.backward()
    w = HG.cpu_weights(rssm)
    for v in w.values():
        v.requires_grad_(True)
    ob_c, ac_c = obs.clone().requires_grad_(True), act.clone().requires_grad_(True)
    kw = {}
    if variant == "discrete":
        kw["forced_post_idx"] = post.stoch_states[:, 1:].detach().reshape(B, T - 1, C, S).argmax(-1).cpu()
    rp, rq = O.observe_rollout(variant, w, ob_c, ac_c, n1, n2, S, C, **kw)
    if variant == "discrete":
        # the kernel's prior sample is one_hot(its own draw); replay it so both sides differentiate the same value
        pi = pri.stoch_states[:, 1:].detach().reshape(B, T - 1, C, S).argmax(-1).cpu()
        p_ = O.categorical_probs(rp["logits"])
        rp_stoch = (torch.nn.functional.one_hot(pi, S).float() + (p_ - p_.detach())).reshape(B, T - 1, C * S)
    else:
        rp_stoch = rp["stoch"]
    k1 = "logits" if variant == "discrete" else "mean"
    rl = (torch.cat([rq["deter"], rq["stoch"]], -1) * gfeat).sum() + (torch.cat([rp["deter"], rp_stoch], -1) * gpri).sum() + \
        (rp[k1] * gd1).sum() + (rq[k1] * gd2).sum()
    if variant == "continuous":
        rl = rl + (rp["std"] * gd2).sum() + (rq["std"] * gd1).sum()
    rl.backward()
    out = {}
    named = dict(rssm.named_parameters())
    for k, v in w.items():
        assert named[k].grad is not None, k
        out[k] = _agree(k, named[k].grad, v.grad)
    out["d obs"] = _agree("d obs_embeds", ob_d.grad, ob_c.grad)
    out["d act"] = _agree("d actions", ac_d.grad, ac_c.grad)
    print(variant, shape, out)
