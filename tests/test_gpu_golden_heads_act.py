"""The MLP heads (fp32 kernels and the bf16 tcgen05 chain) and the acting-step kernel against vectors produced by the
UNMODIFIED reference modules (oracle/gen_golden.py: DenseModel, ValueCritic; posterior -> actor -> prior of
memory_actor.py:51-56 with the latent draws on a tape)."""
import pytest
import torch

from helpers import Golden

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _head(g, head, precision):
    import wmrl
    B, H, F_ = [int(v) for v in g.z["dims"]]
    if head == "critic":
        m = wmrl.ValueCritic(F_, 3, 96, precision=precision)
    elif head == "done":
        m = wmrl.DenseModel(depth=3, input_dim=F_, hidden_dim=32, output_dim=1, output_act=torch.nn.Sigmoid, precision=precision)
    else:
        m = wmrl.DenseModel(depth=2, input_dim=F_, hidden_dim=64, output_dim=1, precision=precision)
    m.load_state_dict(g.weights(head + "."))       # the reference's own state_dict keys
    return m.cuda()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("head", ["reward", "done", "critic"])
def test_heads_vs_reference_vectors(head, precision):
    g = Golden("heads")
    m = _head(g, head, precision)
    x = g.t("x").cuda().requires_grad_(True)
    y = m(x)
    y.backward(g.t(head + "_gy").cuda())
    y_ref, gx_ref = g.t(head + "_y"), g.t(head + "_gx")
    if precision == "fp32":
        torch.testing.assert_close(y.detach().cpu(), y_ref, rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(x.grad.cpu(), gx_ref, rtol=1e-4, atol=1e-5)
        for n, p in m.named_parameters():
            ref = g.t(f"{head}_grad.{n}")
            torch.testing.assert_close(p.grad.cpu(), ref, rtol=1e-4, atol=1e-4 * max(1.0, float(ref.abs().max())))
    else:
        # north_star: rtol 1e-2 for bf16; gradients by direction and norm (35 rows: the ReLU critic's kink flips - see
        # tests/test_gpu_heads_bf16.py - are a visible fraction of so few elements)
        scale = max(1.0, float(y_ref.abs().max()))
        assert float((y.detach().cpu() - y_ref).abs().max()) <= 1e-2 * scale
        lim = (0.98, 0.2) if head == "critic" else (0.999, 3e-2)

        def close(a, b, what):
            a, b = a.double().flatten(), b.double().flatten()
            cos = float(torch.dot(a, b) / (a.norm() * b.norm()))
            rel = float((a - b).norm() / b.norm())
            assert cos >= lim[0] and rel <= lim[1], (what, cos, rel)

        close(x.grad.cpu(), gx_ref, "d x")
        for n, p in m.named_parameters():
            close(p.grad.cpu(), g.t(f"{head}_grad.{n}"), n)


@pytest.mark.parametrize("name", ["act_cont", "act_disc"])
def test_act_step_vs_reference_vectors(name):
    import wmrl
    g = Golden(name)
    B, steps, D, S, C, Hd, E, A, Ha, La = [int(v) for v in g.z["dims"]]
    cont = name == "act_cont"
    rssm = wmrl.ContinuousRSSM(Hd, D, S, E, A) if cont else wmrl.DiscreteRSSM(Hd, D, S, C, E, A)
    bound = g.z["bound"].tolist()
    actor = wmrl.Actor(D + (S if cont else C * S), A, bound, La, Ha)
    rssm.load_state_dict(g.weights("rssm."))
    actor.load_state_dict(g.weights("actor."))
    rssm, actor = rssm.cuda(), actor.cuda()
    state = rssm.initial_state(B)
    state.to("cuda")
    state.deter_state = g.t("deter0").cuda()
    noise, embeds = g.t("noise").cuda(), g.t("embeds").cuda()
    for t in range(steps):
        action, state, post = rssm.act_step(embeds[t], state, actor, noise_post=noise[2 * t], noise_prior=noise[2 * t + 1])
        torch.testing.assert_close(action.cpu(), g.t("step_action")[t], rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(state.deter_state.cpu(), g.t("step_deter")[t], rtol=1e-4, atol=1e-5)
        if cont:
            torch.testing.assert_close(post.mean.cpu(), g.t("step_post_a")[t], rtol=1e-4, atol=1e-5)
            torch.testing.assert_close(post.std.cpu(), g.t("step_post_b")[t], rtol=1e-4, atol=1e-5)
            torch.testing.assert_close(post.stoch_state.cpu(), g.t("step_post_stoch")[t], rtol=1e-4, atol=2e-5)
            torch.testing.assert_close(state.stoch_state.cpu(), g.t("step_pri_stoch")[t], rtol=1e-4, atol=5e-5)
            torch.testing.assert_close(state.std.cpu(), g.t("step_pri_b")[t], rtol=1e-4, atol=2e-5)
        else:
            torch.testing.assert_close(post.logits.cpu(), g.t("step_post_a")[t], rtol=1e-4, atol=1e-5)
            assert torch.equal(post.stoch_state.cpu(), g.t("step_post_stoch")[t])      # indices bit-exact (fp32 mode)
            assert torch.equal(state.stoch_state.cpu(), g.t("step_pri_stoch")[t])
            torch.testing.assert_close(state.logits.cpu(), g.t("step_pri_a")[t], rtol=1e-4, atol=2e-5)
