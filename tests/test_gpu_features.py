"""get_features() of a state sequence as one native streaming pass each way (wmrl_features_fwd / _bwd) against
torch.cat((deter[:, 1:], stoch[:, 1:]), -1) (state/continuous.py:73-77): values and gradients, bit for bit."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


@pytest.mark.parametrize("B,T,D,S", [(37, 16, 200, 30), (9, 65, 1024, 1024), (5, 5, 27, 7), (3, 1, 32, 8), (0, 4, 32, 8),
                                     (2500, 16, 600, 1024)])
def test_sequence_features_vs_torch_cat(B, T, D, S):
    import wmrl
    import wmrl.functional as Fn
    g = torch.Generator().manual_seed(B + T + D)
    d = torch.randn(B, T, D, generator=g).cuda().requires_grad_(True)
    s = torch.randn(B, T, S, generator=g).cuda().requires_grad_(True)
    f = Fn.sequence_features(d, s)
    ref = torch.cat((d[:, 1:], s[:, 1:]), -1)
    assert torch.equal(f, ref)
    gy = torch.randn(f.shape, generator=g).cuda()
    f.backward(gy)
    gd, gs = d.grad.clone(), s.grad.clone()
    d.grad = s.grad = None
    ref.backward(gy)
    assert torch.equal(gd, d.grad) and torch.equal(gs, s.grad)
    if B > 0:
        seq = wmrl.InternalStateContinuousSequence(deter_states=d.detach(), stoch_states=s.detach(),
                                                   means=s.detach(), stds=s.detach())
        assert torch.equal(seq.get_features(), ref.detach())
