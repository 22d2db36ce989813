"""GPU parity, fp32 validation mode: the CUDA path (through the drop-in modules
-> ctypes -> C ABI) against (a) the golden vectors produced by the unmodified
reference and (b) the oracle on larger seeded inputs.

Tolerance (BASELINE.json north_star): rtol 1e-5 for fp32 mode - with an atol of
the same size for values near zero; categorical indices bit-exact except at
documented near-ties (relative margin of the top two p/q below 1e-4)."""
import pytest
import torch

from helpers import CASES, LAMBDA_CASES, Golden
from oracle import rssm_oracle as O

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 2e-5
G_RTOL, G_ATOL = 2e-4, 2e-6     # gradients: sums over B*H terms in a different order


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import helpers_gpu  # noqa: F401  (imports wmrl; fails loudly if the .so is missing)
    return "cuda"


def close(a, b, rtol=RTOL, atol=ATOL, msg=""):
    torch.testing.assert_close(a.detach().cpu().float(), b.detach().cpu().float(), rtol=rtol, atol=atol,
                               msg=lambda m: f"{msg}: {m}")


@pytest.mark.parametrize("name", CASES)
def test_imagine_forward_golden(name):
    dev = _dev()
    import helpers_gpu as HG
    g = Golden(name)
    rssm, actor = HG.modules_from_golden(g)
    init = HG.init_state(rssm, g.t("deter0"), g.t("stoch0"))
    with torch.no_grad():
        seq, actions = rssm.imagine_rollout(init, actor, g.H, noise=g.t("noise_im").to(dev), return_actions=True)
    assert seq.deter_states.shape == (g.B, g.H + 1, g.D)
    close(seq.deter_states, g.t("im_deter"), msg="deter")
    close(actions, g.t("im_actions"), msg="actions")
    if g.variant == "continuous":
        close(seq.stoch_states, g.t("im_stoch"), msg="stoch")
        close(seq.means[:, 1:], g.t("im_mean")[:, 1:], msg="mean")
        close(seq.stds[:, 1:], g.t("im_std")[:, 1:], msg="std")
    else:
        close(seq.logits[:, 1:], g.t("im_logits")[:, 1:], msg="logits")
        assert torch.equal(seq.stoch_states.cpu(), g.t("im_stoch")), "categorical samples must match bit-exactly"


@pytest.mark.parametrize("name", CASES)
def test_imagine_actor_bptt_golden(name):
    dev = _dev()
    import helpers_gpu as HG
    g = Golden(name)
    rssm, actor = HG.modules_from_golden(g)
    for p in rssm.parameters():          # WorldModel.imagine_rollout freezes the WM (world_model.py:172)
        p.requires_grad_(False)
    init = HG.init_state(rssm, g.t("deter0"), g.t("stoch0"))
    seq = rssm.imagine_rollout(init, actor, g.H, noise=g.t("noise_im").to(dev))
    loss = (seq.get_features() * g.t("im_gfeat").to(dev)).sum()
    if g.variant == "continuous":
        loss = loss + (seq.means[:, 1:] * g.t("im_gmean").to(dev)).sum() + (seq.stds[:, 1:] * g.t("im_gstd").to(dev)).sum()
    else:
        loss = loss + (seq.logits[:, 1:] * g.t("im_glogits").to(dev)).sum()
    loss.backward()
    ref = g.grads("im_grad.actor.")
    named = dict(actor.named_parameters())
    assert named["stddev.weight"].grad is None      # reference: the sigma head gets no gradient
    for k, v in ref.items():
        close(named[k].grad, v, G_RTOL, G_ATOL, msg=k)


@pytest.mark.parametrize("name", CASES)
def test_observe_golden(name):
    dev = _dev()
    import helpers_gpu as HG
    g = Golden(name)
    rssm, _ = HG.modules_from_golden(g)
    obs = g.t("ob_obs").to(dev).requires_grad_(True)
    act = g.t("ob_act").to(dev).requires_grad_(True)
    pri, post = rssm.observe_rollout(obs, act, noise=(g.t("noise_pri").to(dev), g.t("noise_post").to(dev)))
    fields = {"deter": "deter_states", "stoch": "stoch_states"}
    fields.update({"mean": "means", "std": "stds"} if g.variant == "continuous" else {"logits": "logits"})
    loss = 0.0
    for tag, sq in (("pri", pri), ("pos", post)):
        for k, attr in fields.items():
            v = getattr(sq, attr)
            assert v.shape[:2] == (g.B, g.T)
            close(v[:, 1:], g.t(f"ob_{tag}_{k}")[:, 1:], msg=f"{tag}.{k}")
            loss = loss + (v[:, 1:] * g.t(f"ob_g_{tag}_{k}").to(dev)).sum()
        # index 0: zero initial state (rssm.py:166-172)
        assert float(sq.deter_states[:, 0].abs().max()) == 0.0
        assert float(sq.stoch_states[:, 0].abs().max()) == 0.0
    loss.backward()
    named = dict(rssm.named_parameters())
    for k, v in g.grads("ob_grad.rssm.").items():
        close(named[k].grad, v, G_RTOL, G_ATOL, msg=k)
    close(obs.grad, g.t("ob_grad.obs"), G_RTOL, G_ATOL, msg="d obs")
    close(act.grad, g.t("ob_grad.act"), G_RTOL, G_ATOL, msg="d act")


@pytest.mark.parametrize("name", LAMBDA_CASES)
def test_lambda_return_golden(name):
    dev = _dev()
    from wmrl.functional import done_mask, lambda_return
    g = Golden(name)
    V = g.t("V").to(dev).requires_grad_(True)
    r = g.t("r").to(dev).requires_grad_(True)
    d = done_mask(g.t("raw_d").to(dev))
    assert torch.equal(d.cpu(), g.t("d"))
    tgt = lambda_return(V, r, d, float(g.z["gamma"]), float(g.z["lam"]))
    close(tgt, g.t("target"), 1e-5, 2e-6, msg="targets")
    (tgt * g.t("g_target").to(dev)).sum().backward()
    close(V.grad, g.t("g_V"), 1e-5, 2e-6, msg="dV")
    close(r.grad, g.t("g_r"), 1e-5, 2e-6, msg="dr")


# ---------------------------------------------------------------- vs oracle, larger shapes
SHAPES = {
    # name: variant, B, H, D, S, C, Hd, E, A, Ha, La
    "c1_slice": ("continuous", 300, 15, 200, 30, 1, 200, 64, 6, 512, 3),       # BASELINE configs[0] dims
    "c2_slice": ("discrete", 96, 6, 600, 32, 32, 600, 64, 6, 512, 3),          # configs[1] dims
    "odd": ("discrete", 33, 3, 50, 5, 7, 24, 10, 2, 48, 1),
}


@pytest.mark.parametrize("shape", list(SHAPES))
def test_imagine_vs_oracle(shape):
    dev = _dev()
    import helpers_gpu as HG
    variant, B, H, D, S, C, Hd, E, A, Ha, La = SHAPES[shape]
    rssm, actor = HG.build_modules(variant, D, S, C, Hd, E, A, Ha, La, bound=1.0, seed=3)
    for p in rssm.parameters():
        p.requires_grad_(False)
    deter0, stoch0, noise = HG.synth_inputs(variant, B, H, D, S, C)
    init = HG.init_state(rssm, deter0, stoch0)
    seq, actions = rssm.imagine_rollout(init, actor, H, noise=noise.to(dev), return_actions=True)
    w, aw = HG.cpu_weights(rssm), HG.cpu_weights(actor)
    for v in aw.values():
        v.requires_grad_(True)
    ref = O.imagine_rollout(variant, w, aw, deter0, stoch0, noise, H, S, C, 5.0, torch.tensor(1.0))
    gen = torch.Generator().manual_seed(9)
    gfeat = torch.randn(B, H, D + (S if variant == "continuous" else S * C), generator=gen) / B
    if variant == "discrete":
        got_idx = seq.stoch_states[:, 1:].reshape(B, H, C, S).argmax(-1).cpu()
        same = got_idx == ref["idx"]
        if not bool(same.all()):
            # near-tie policy: a flipped draw is only accepted where the oracle margin is < 1e-4,
            # and the rows it touches are excluded from the value comparison afterwards
            tie = HG.near_tie_mask(ref["logits"].detach(), noise.permute(1, 0, 2, 3))
            assert bool((same | tie).all()), "categorical index mismatch away from a near-tie"
        rows = same.all(dim=2).all(dim=1)
        assert rows.float().mean() > 0.99
    else:
        rows = torch.ones(B, dtype=torch.bool)
    close(seq.deter_states[:, 1:].cpu()[rows], ref["deter"].detach()[rows], msg="deter")
    close(actions.cpu()[rows], ref["actions"].detach()[rows], msg="actions")
    if variant == "continuous":
        close(seq.stoch_states[:, 1:], ref["stoch"], msg="stoch")
        close(seq.means[:, 1:], ref["mean"], msg="mean")
        close(seq.stds[:, 1:], ref["std"], msg="std")
    else:
        close(seq.logits[:, 1:].cpu()[rows], ref["logits"].detach()[rows], 1e-5, 5e-5, msg="logits")
    if bool(rows.all()):
        (seq.get_features() * gfeat.to(dev)).sum().backward()
        (torch.cat([ref["deter"], ref["stoch"]], -1) * gfeat).sum().backward()
        named = dict(actor.named_parameters())
        for k, v in aw.items():
            if v.grad is None:
                assert named[k].grad is None, k
                continue
            scale = float(v.grad.abs().max())
            close(named[k].grad, v.grad, 1e-3, 1e-4 * scale + 1e-7, msg=k)


def test_value_trainer_composition():
    """WorldModel.imagine_rollout -> ValueGradTrainer.update (test_value_rssm.py:7-46 composition),
    lambda targets checked against the reference loops restated in the oracle."""
    dev = _dev()
    import helpers_gpu as HG
    import wmrl
    variant, B, H, D, S, C, Hd, E, A, Ha, La = ("continuous", 64, 10, 32, 8, 1, 32, 16, 2, 64, 2)
    rssm, actor = HG.build_modules(variant, D, S, C, Hd, E, A, Ha, La, seed=4)
    F_ = D + S
    torch.manual_seed(7)
    wm = wmrl.WorldModel(encoder=torch.nn.Linear(4, E), decoder=torch.nn.Linear(F_, 4),
                         done_model=wmrl.DenseModel(2, F_, 32, 1, output_act=torch.nn.Sigmoid),
                         reward_model=wmrl.DenseModel(2, F_, 32, 1), dynamic_model=rssm).to(dev)
    critic = wmrl.ValueCritic(F_, 2, 32).to(dev)
    trainer = wmrl.ValueGradTrainer(actor, critic)
    deter0, stoch0, noise = HG.synth_inputs(variant, B, H, D, S, C)
    init = HG.init_state(rssm, deter0, stoch0)
    roll = wm.imagine_rollout(init, actor, H, noise=noise.to(dev))
    assert roll.rewards.shape == (B, H, 1) and roll.features.shape == (B, H, F_)
    assert all(p.requires_grad for p in wm.parameters())      # FreezeParameters restored
    d = wmrl.functional.done_mask(roll.dones)
    loss, targets = trainer.value_loss(roll.features, roll.rewards, d)
    with torch.no_grad():
        V = trainer.target_critic(roll.features)
    ref = O.lambda_return_loops(V.cpu(), roll.rewards.detach().cpu(), d.cpu(), 0.99, 0.95)
    close(targets, ref, 1e-5, 1e-5, msg="lambda targets vs reference loops")
    before = [p.detach().clone() for p in actor.parameters()]
    losses = trainer.update(roll.features, roll.rewards, roll.dones)
    assert losses.value_loss is not None and losses.actor_loss is not None
    changed = [not torch.equal(a, b) for a, b in zip(before, actor.parameters())]
    assert any(changed), "actor received no gradient through the imagined rollout"


def test_edge_cases():
    dev = _dev()
    import helpers_gpu as HG
    import wmrl
    rssm, actor = HG.build_modules("continuous", 16, 4, 1, 16, 8, 1, 32, 1, seed=1)
    # empty batch
    init = HG.init_state(rssm, torch.zeros(0, 16), torch.zeros(0, 4))
    seq = rssm.imagine_rollout(init, actor, 3)
    assert seq.deter_states.shape == (0, 4, 16)
    # zero steps: only the initial slot (get_features -> (B,0,F), test_continuous_states.py:85)
    init = HG.init_state(rssm, torch.ones(2, 16) * 0.5, torch.ones(2, 4))
    seq = rssm.imagine_rollout(init, actor, 0)
    assert seq.deter_states.shape == (2, 1, 16) and seq.get_features().shape == (2, 0, 20)
    assert torch.equal(seq.deter_states[:, 0].cpu(), torch.ones(2, 16) * 0.5)
    # observe with T=1: no steps
    pri, post = rssm.observe_rollout(torch.zeros(2, 1, 8, device=dev), torch.zeros(2, 1, 1, device=dev))
    assert post.stoch_states.shape == (2, 1, 4)
    # wrong noise shape is rejected
    with pytest.raises(ValueError):
        rssm.imagine_rollout(init, actor, 2, noise=torch.zeros(3, 2, 4, device=dev))
    # reference shape contracts (tests/test_rssm.py:27-28,53): B=32, n=10 -> (32,11,.)
    r2, a2 = HG.build_modules("discrete", 32, 8, 4, 32, 64, 1, 64, 3, seed=2)
    st = r2.initial_state(32)
    seq = r2.imagine_rollout(st, a2, 10)
    assert seq.stoch_states.shape == (32, 11, 32) and seq.logits.shape == (32, 11, 4, 8)
    pri, post = r2.observe_rollout(torch.randn(32, 10, 64, device=dev), torch.randn(32, 10, 1, device=dev))
    assert post.stoch_states.shape == (32, 10, 32)
    flat = post.flatten_batch_time()
    assert flat.deter_state.shape == (288, 32)          # 32 x (10-1) start states (test_rssm_world_model.py:83)
    # one-hot samples
    s = seq.stoch_states[:, 1:].reshape(32, 10, 4, 8)
    assert torch.equal(s.sum(-1), torch.ones(32, 10, 4, device=dev))


@pytest.mark.parametrize("B", [40, 150])
@pytest.mark.parametrize("variant", ["continuous", "discrete"])
def test_observe_vs_oracle_few_rows_long_k(variant, B):
    """posterior filtering with K >= 128 at B = 40 (cluster split-K skinny GEMM for fwd / dgrad next to the
    tensor-core one for wgrad) and B = 150 (3xTF32 tcgen05 GEMM everywhere, ragged row tile); forward fields
    and every RSSM / input gradient against autograd on the oracle."""
    dev = _dev()
    import helpers_gpu as HG
    T, D, S, C, Hd, E, A = 5, 160, 12, (6 if variant == "discrete" else 1), 144, 200, 3
    rssm, _ = HG.build_modules(variant, D, S, C, Hd, E, A, 64, 1, seed=11)
    g = torch.Generator().manual_seed(12)
    obs = torch.relu(torch.randn(B, T, E, generator=g))
    act = torch.rand(B, T, A, generator=g) * 2 - 1
    mk = (lambda: torch.empty(T - 1, B, C, S).exponential_(1.0, generator=g)) if variant == "discrete" else \
        (lambda: torch.randn(T - 1, B, S, generator=g))
    n1, n2 = mk(), mk()
    obs_d, act_d = obs.to(dev).requires_grad_(True), act.to(dev).requires_grad_(True)
    pri, post = rssm.observe_rollout(obs_d, act_d, noise=(n1.to(dev), n2.to(dev)))
    w = HG.cpu_weights(rssm)
    for v in w.values():
        v.requires_grad_(True)
    obs_c, act_c = obs.clone().requires_grad_(True), act.clone().requires_grad_(True)
    rpri, rpost = O.observe_rollout(variant, w, obs_c, act_c, n1, n2, S, C)
    if variant == "discrete":
        assert torch.equal(post.stoch_states[:, 1:].cpu(), rpost["stoch"].detach())
        close(post.logits[:, 1:], rpost["logits"], 1e-5, 5e-5, msg="post logits")
    else:
        close(post.stoch_states[:, 1:], rpost["stoch"], msg="post stoch")
        close(pri.means[:, 1:], rpri["mean"], msg="prior mean")
    close(post.deter_states[:, 1:], rpost["deter"], msg="deter")
    gw = torch.randn(B, T - 1, D + S * C, generator=g) / B
    (post.get_features() * gw.to(dev)).sum().backward()
    (torch.cat([rpost["deter"], rpost["stoch"]], -1) * gw).sum().backward()
    named = dict(rssm.named_parameters())
    for k, v in w.items():
        if v.grad is None:
            continue
        scale = float(v.grad.abs().max())
        close(named[k].grad, v.grad, 1e-3, 1e-4 * scale + 1e-7, msg=k)
    close(obs_d.grad, obs_c.grad, 1e-3, 1e-6, msg="d obs")
    close(act_d.grad, act_c.grad, 1e-3, 1e-6, msg="d act")
