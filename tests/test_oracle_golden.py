"""CPU gate: the oracle restatement (oracle/rssm_oracle.py) against vectors
produced by the unmodified reference (oracle/gen_golden.py)."""
import pytest
import torch

from helpers import CASES, LAMBDA_CASES, Golden
from oracle import rssm_oracle as O

torch.set_num_threads(1)


def _imagine(g, dtype=torch.float32, grad=False):
    w, aw = g.weights("rssm.", dtype), g.weights("actor.", dtype)
    if grad:
        for v in aw.values():
            v.requires_grad_(True)
    res = O.imagine_rollout(g.variant, w, aw, g.t("deter0", dtype), g.t("stoch0", dtype),
                            g.t("noise_im", dtype), g.H, g.S, g.C, 5.0, g.t("bound", dtype))
    return res, w, aw


@pytest.mark.parametrize("name", CASES)
def test_imagine_matches_reference(name):
    g = Golden(name)
    res, _, _ = _imagine(g)
    # same ATen ops in the same order on one thread -> tight
    torch.testing.assert_close(res["deter"], g.t("im_deter")[:, 1:], rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(res["stoch"].detach(), g.t("im_stoch")[:, 1:], rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(res["actions"], g.t("im_actions"), rtol=1e-6, atol=1e-6)
    if g.variant == "continuous":
        torch.testing.assert_close(res["mean"], g.t("im_mean")[:, 1:], rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(res["std"], g.t("im_std")[:, 1:], rtol=1e-6, atol=1e-6)
    else:
        torch.testing.assert_close(res["logits"], g.t("im_logits")[:, 1:], rtol=1e-6, atol=1e-6)
        # forward value of the straight-through sample is exactly one-hot
        ref_idx = g.t("im_stoch")[:, 1:].reshape(g.B, g.H, g.C, g.S).argmax(-1)
        assert torch.equal(res["idx"], ref_idx)


@pytest.mark.parametrize("name", CASES)
def test_imagine_actor_grads_match_reference(name):
    g = Golden(name)
    res, _, aw = _imagine(g, grad=True)
    feats = torch.cat([res["deter"], res["stoch"]], -1)
    loss = (feats * g.t("im_gfeat")).sum()
    if g.variant == "continuous":
        loss = loss + (res["mean"] * g.t("im_gmean")).sum() + (res["std"] * g.t("im_gstd")).sum()
    else:
        loss = loss + (res["logits"] * g.t("im_glogits")).sum()
    loss.backward()
    ref = g.grads("im_grad.actor.")
    assert "mu.weight" in ref and "stddev.weight" not in ref   # stddev head gets no grad
    for k, v in ref.items():
        torch.testing.assert_close(aw[k].grad, v, rtol=1e-4, atol=1e-6, msg=lambda m: f"{k}: {m}")


@pytest.mark.parametrize("name", CASES)
def test_observe_matches_reference(name):
    g = Golden(name)
    w = g.weights("rssm.")
    for v in w.values():
        v.requires_grad_(True)
    obs, act = g.t("ob_obs").requires_grad_(True), g.t("ob_act").requires_grad_(True)
    pri, pos = O.observe_rollout(g.variant, w, obs, act, g.t("noise_pri"), g.t("noise_post"), g.S, g.C)
    fields = ("deter", "stoch", "mean", "std") if g.variant == "continuous" else ("deter", "stoch", "logits")
    loss = 0.0
    for tag, sq in (("pri", pri), ("pos", pos)):
        for k in fields:
            torch.testing.assert_close(sq[k].detach(), g.t(f"ob_{tag}_{k}")[:, 1:], rtol=1e-6, atol=1e-6)
            loss = loss + (sq[k] * g.t(f"ob_g_{tag}_{k}")).sum()
    loss.backward()
    for k, v in g.grads("ob_grad.rssm.").items():
        torch.testing.assert_close(w[k].grad, v, rtol=1e-4, atol=1e-6, msg=lambda m: f"{k}: {m}")
    torch.testing.assert_close(obs.grad, g.t("ob_grad.obs"), rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(act.grad, g.t("ob_grad.act"), rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("name", LAMBDA_CASES)
def test_lambda_return_matches_reference(name):
    g = Golden(name)
    gamma, lam = float(g.z["gamma"]), float(g.z["lam"])
    V, r = g.t("V").requires_grad_(True), g.t("r").requires_grad_(True)
    d = O.done_preprocess(g.t("raw_d"))
    assert torch.equal(d, g.t("d"))
    loops = O.lambda_return_loops(V, r, d, gamma, lam)
    torch.testing.assert_close(loops, g.t("target"), rtol=1e-6, atol=1e-6)
    scan = O.lambda_return_scan(V, r, d, gamma, lam)
    # O(H) restatement; residual is the reference's fp32 gamma table + summation order
    torch.testing.assert_close(scan, g.t("target"), rtol=1e-5, atol=2e-6)
    (scan * g.t("g_target")).sum().backward()
    torch.testing.assert_close(V.grad, g.t("g_V"), rtol=1e-5, atol=2e-6)
    r_grad = r.grad if r.grad is not None else torch.zeros_like(r)   # H=1: rewards unused
    torch.testing.assert_close(r_grad, g.t("g_r"), rtol=1e-5, atol=2e-6)


def test_lambda_weights_quirk():
    """n=2 gives lambda*G_1; weights do not sum to one."""
    V = torch.tensor([[[2.0], [3.0]]]); r = torch.tensor([[[1.0], [5.0]]]); d = torch.zeros(1, 2, 1)
    t = O.lambda_return_scan(V, r, d, 0.9, 0.5)
    assert torch.allclose(t[0, 0], torch.tensor([0.5 * (1.0 + 0.9 * 3.0)]))
    assert torch.allclose(t[0, 1], torch.tensor([3.0]))


# ---- heads and the online acting step: oracle restatements vs the unmodified reference's vectors ------------------
HEAD_ACTS = {"reward": "silu", "done": "silu", "critic": "relu"}
HEAD_SEQ = {"reward": "fc.", "done": "fc.", "critic": "layers."}


@pytest.mark.parametrize("head", ["reward", "done", "critic"])
def test_mlp_head_matches_reference(head):
    g = Golden("heads")
    sd = {k: v.clone().requires_grad_(True) for k, v in g.weights(head + "." + HEAD_SEQ[head]).items()}
    x = g.t("x").requires_grad_(True)
    y = O.mlp_head(sd, x, HEAD_ACTS[head])
    if head == "done":
        y = torch.sigmoid(y)          # DenseModel.output_act (models.py:37)
    torch.testing.assert_close(y, g.t(head + "_y"), rtol=1e-6, atol=1e-6)
    y.backward(g.t(head + "_gy"))
    torch.testing.assert_close(x.grad, g.t(head + "_gx"), rtol=1e-5, atol=1e-6)
    for k, v in sd.items():
        torch.testing.assert_close(v.grad, g.t(f"{head}_grad.{HEAD_SEQ[head]}{k}"), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", ["act_cont", "act_disc"])
def test_act_step_matches_reference(name):
    g = Golden(name)
    B, steps, D, S, C, Hd, E, A, Ha, La = [int(v) for v in g.z["dims"]]
    variant = "continuous" if name == "act_cont" else "discrete"
    w, aw = g.weights("rssm."), g.weights("actor.")
    deter = g.t("deter0")
    noise, embeds = g.t("noise"), g.t("embeds")
    for t in range(steps):
        r = O.act_step(variant, w, aw, embeds[t], deter, noise[2 * t], noise[2 * t + 1], S, C, 5.0, g.t("bound"))
        torch.testing.assert_close(r["action"], g.t("step_action")[t], rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(r["deter_next"], g.t("step_deter")[t], rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(r["post_stoch"].detach(), g.t("step_post_stoch")[t], rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(r["prior_stoch"].detach(), g.t("step_pri_stoch")[t], rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(r["post_a"], g.t("step_post_a")[t], rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(r["prior_a"], g.t("step_pri_a")[t], rtol=1e-6, atol=1e-6)
        if variant == "continuous":
            torch.testing.assert_close(r["post_b"], g.t("step_post_b")[t], rtol=1e-6, atol=1e-6)
            torch.testing.assert_close(r["prior_b"], g.t("step_pri_b")[t], rtol=1e-6, atol=1e-6)
        deter = r["deter_next"]
