"""The fused KL(prior || posterior) + free-nats clamp of WorldModel.update (world_model.py:131-141) against the
torch.distributions expression the reference evaluates, values and gradients (fp32: rtol 1e-5 class)."""
import pytest
import torch
import torch.distributions as D

pytestmark = pytest.mark.gpu


def _ref(variant, pa, pb, qa, qb, rho):
    if variant == "continuous":
        p = D.Independent(D.Normal(pa[:, 1:], pb[:, 1:]), 1)
        q = D.Independent(D.Normal(qa[:, 1:], qb[:, 1:]), 1)
    else:
        p = D.Independent(D.OneHotCategoricalStraightThrough(logits=pa[:, 1:]), 1)
        q = D.Independent(D.OneHotCategoricalStraightThrough(logits=qa[:, 1:]), 1)
    kl = D.kl_divergence(p, q)
    return torch.max(kl, torch.ones_like(kl) * rho).mean(), kl.mean(), kl


@pytest.mark.parametrize("variant,B,T,S,C", [("continuous", 50, 11, 30, 1), ("continuous", 3, 2, 7, 1),
                                             ("discrete", 50, 11, 32, 32), ("discrete", 9, 4, 8, 4), ("discrete", 5, 3, 20, 3)])
def test_kl_loss_matches_torch_distributions(variant, B, T, S, C):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from wmrl import _lib
    from wmrl.functional import PathConfig, kl_loss
    g = torch.Generator().manual_seed(B + T)
    dev = "cuda"
    if variant == "continuous":
        mk = lambda: torch.randn(B, T, S, generator=g).to(dev).requires_grad_(True)   # noqa: E731
        sd = lambda: (torch.rand(B, T, S, generator=g) * 1.5 + 0.1).to(dev).requires_grad_(True)   # noqa: E731
        pa, pb, qa, qb = mk(), sd(), mk(), sd()
        cfg = PathConfig(variant=_lib.CONTINUOUS, D=1, S=S, C=1, A=1, Hd=1, E=1)
    else:
        mk = lambda: (2 * torch.randn(B, T, C, S, generator=g)).to(dev).requires_grad_(True)   # noqa: E731
        pa, pb, qa, qb = mk(), None, mk(), None
        cfg = PathConfig(variant=_lib.DISCRETE, D=1, S=S, C=C, A=1, Hd=1, E=1)
    # a rho in the middle of the KL range so that the clamp is active on part of the rows
    with torch.no_grad():
        rho = float(_ref(variant, pa, pb, qa, qb, 0.0)[2].median()) * 1.000123   # not exactly on a row (torch splits ties)
    kc, km, kl = kl_loss(cfg, rho, pa, pb, qa, qb)
    rc, rm, rkl = _ref(variant, pa, pb, qa, qb, rho)
    torch.testing.assert_close(kl, rkl.detach(), rtol=2e-5, atol=2e-5)
    torch.testing.assert_close(kc, rc.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(km, rm.detach(), rtol=1e-5, atol=1e-6)
    (3.0 * kc).backward()
    mine = [t.grad.clone() for t in (pa, pb, qa, qb) if t is not None]
    for t in (pa, pb, qa, qb):
        if t is not None:
            t.grad = None
    (3.0 * rc).backward()
    for m, t in zip(mine, [t for t in (pa, pb, qa, qb) if t is not None]):
        torch.testing.assert_close(m, t.grad, rtol=2e-4, atol=1e-7)
        assert float(m[:, 0].abs().max()) == 0.0


@pytest.mark.parametrize("B,T", [(50, 50), (3, 2), (1, 7)])
def test_head_losses_vs_torch(B, T):
    """reward NLL + done BCE of WorldModel.update (world_model.py:130, 143-144): the fused pass vs torch.distributions /
    F.binary_cross_entropy_with_logits, values and gradients."""
    import torch.distributions as D
    import torch.nn.functional as F
    import wmrl.functional as Fn
    g = torch.Generator().manual_seed(B * 100 + T)
    pr = (torch.randn(B, T, 1, generator=g) * 2).cuda().requires_grad_(True)
    tr = torch.randn(B, T, 1, generator=g).cuda()
    ld = (torch.randn(B, T, 1, generator=g) * 4).cuda().requires_grad_(True)
    td = (torch.rand(B, T, 1, generator=g) > 0.7).float().cuda()
    rl, dl = Fn.head_losses(pr, tr, ld, td)
    (10.0 * rl + 0.5 * dl).backward()
    pr2, ld2 = pr.detach().clone().requires_grad_(True), ld.detach().clone().requires_grad_(True)
    rl2 = -D.Independent(D.Normal(pr2, 1.0), 1).log_prob(tr).mean()
    dl2 = F.binary_cross_entropy_with_logits(ld2, td)
    (10.0 * rl2 + 0.5 * dl2).backward()
    torch.testing.assert_close(rl, rl2, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(dl, dl2, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(pr.grad, pr2.grad, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(ld.grad, ld2.grad, rtol=1e-5, atol=1e-7)
