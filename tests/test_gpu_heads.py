"""The MLP heads next to the path (DenseModel, ValueCritic) on libwmrl's kernels vs the same modules run by
torch on CPU in fp64: forward values and every gradient.  Also the fused LayerNorm+activation op alone."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


@pytest.mark.parametrize("act", [0, 1, 2])
@pytest.mark.parametrize("M,W", [(300, 512), (1000, 256), (7, 40), (65, 600), (130, 1000)])
def test_layer_norm_act(act, M, W):
    import wmrl.functional as F
    g = torch.Generator().manual_seed(M + W + act)
    x = torch.randn(M, W, generator=g) * 2 + 0.3
    gam, bet = torch.rand(W, generator=g) + 0.5, torch.randn(W, generator=g) * 0.2
    gy = torch.randn(M, W, generator=g)
    fn = [torch.nn.functional.elu, torch.nn.functional.silu, torch.relu][act]
    xd, gd, bd = (t.double().requires_grad_(True) for t in (x, gam, bet))
    ref = fn(torch.nn.functional.layer_norm(xd, (W,), gd, bd, 1e-5))
    ref.backward(gy.double())
    xc, gc, bc = (t.cuda().requires_grad_(True) for t in (x, gam, bet))
    y = F.layer_norm_act(xc, gc, bc, act)
    y.backward(gy.cuda())
    torch.testing.assert_close(y.detach().cpu().double(), ref.detach(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(xc.grad.cpu().double(), xd.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(gc.grad.cpu().double(), gd.grad, rtol=1e-4, atol=1e-4 * float(gd.grad.abs().max()))
    torch.testing.assert_close(bc.grad.cpu().double(), bd.grad, rtol=1e-4, atol=1e-4 * float(bd.grad.abs().max()))


@pytest.mark.parametrize("kind", ["reward", "done", "critic"])
def test_heads_match_torch(kind):
    import wmrl
    torch.manual_seed(3)
    F_, B, H = 230, 70, 5
    if kind == "critic":
        m = wmrl.ValueCritic(F_, num_layers=3, hidden_dim=512)
    elif kind == "done":
        m = wmrl.DenseModel(depth=3, input_dim=F_, hidden_dim=256, output_dim=1, output_act=torch.nn.Sigmoid)
    else:
        m = wmrl.DenseModel(depth=3, input_dim=F_, hidden_dim=256, output_dim=1)
    ref = copy.deepcopy(m).double()            # CPU tensors -> the torch modules themselves run
    x = torch.randn(B, H, F_)
    xr = x.double().requires_grad_(True)
    yr = ref(xr)
    gy = torch.randn_like(yr)
    yr.backward(gy)
    mc = m.cuda()
    xc = x.cuda().requires_grad_(True)
    yc = mc(xc)
    assert yc.shape == yr.shape
    yc.backward(gy.float().cuda())
    torch.testing.assert_close(yc.detach().cpu().double(), yr.detach(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(xc.grad.cpu().double(), xr.grad, rtol=1e-3, atol=1e-6)
    for (n, p), (_, q) in zip(mc.named_parameters(), ref.named_parameters()):
        scale = float(q.grad.abs().max())
        torch.testing.assert_close(p.grad.cpu().double(), q.grad, rtol=1e-3, atol=1e-4 * scale + 1e-9, msg=lambda s, n=n: f"{n}: {s}")
    assert list(mc.state_dict()) == list(ref.state_dict())
