"""Numerics of the Linear primitive (wmrl_linear_fwd / _dgrad / _wgrad) against plain PyTorch fp32.

The tcgen05 kernel computes 3xTF32 (hi*hi + hi*lo + lo*hi, fp32 accumulate), which should agree with
an fp32 GEMM to fp32 rounding.  Tolerance: |err| <= 2e-6 * sum_k |x_k w_k| (a few ulp of the
magnitude an fp32 dot product of that length accumulates), checked against an fp64 reference, and
the same bound must hold for torch's own fp32 result (sanity of the bound).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _ref64(x, w, b, act):
    y = x.double() @ w.double().t()
    if b is not None:
        y = y + b.double()
    if act:
        y = torch.nn.functional.elu(y)
    return y


SHAPES = [
    (128, 256, 64), (2500, 600, 230), (2500, 200, 36), (300, 1800, 600), (65, 16, 512), (1000, 6, 512),
    (129, 257, 17), (4096, 1024, 1624), (777, 60, 200), (50, 600, 1624), (3, 40, 9),
]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("act", [0, 1])
def test_linear_fwd_matches_fp32(M, N, K, act):
    import wmrl.functional as F
    g = torch.Generator().manual_seed(M * 31 + N * 7 + K)
    x = torch.randn(M, K, generator=g).cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    b = torch.randn(N, generator=g).cuda()
    with torch.no_grad():
        y = F.linear(x, w, b, act)
    ref = _ref64(x, w, b, act)
    mag = x.abs().double() @ w.abs().double().t() + b.abs().double()
    err = (y.double() - ref).abs()
    # the tensor core's fp32 accumulate truncates: one ulp-level loss per chained MMA (3 per 8 k)
    bound = max(2e-6, 4e-9 * K)
    assert float((err / mag).max()) < bound, (float((err / mag).max()), bound)
    torch.backends.cuda.matmul.allow_tf32 = False
    y32 = torch.nn.functional.linear(x, w, b)
    if act:
        y32 = torch.nn.functional.elu(y32)
    assert float(((y32.double() - ref).abs() / mag).max()) < 2e-6


@pytest.mark.parametrize("M,N,K", [(2500, 600, 230), (129, 257, 17), (4096, 512, 512), (70000, 60, 200),
                                   (50, 600, 1624), (1000, 6, 512)])
def test_linear_backward_matches_fp32(M, N, K):
    import wmrl.functional as F
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g).cuda().requires_grad_(True)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).cuda().requires_grad_(True)
    b = torch.randn(N, generator=g).cuda().requires_grad_(True)
    gy = torch.randn(M, N, generator=g).cuda()
    F.linear(x, w, b, 1).backward(gy)
    got = [x.grad.clone(), w.grad.clone(), b.grad.clone()]
    xd, wd, bd = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    torch.nn.functional.elu(xd @ wd.t() + bd).backward(gy.double())
    for a, r, name in zip(got, (xd.grad, wd.grad, bd.grad), ("dx", "dw", "db")):
        scale = float(r.abs().max())
        assert float((a.double() - r).abs().max()) <= 2e-5 * scale, (name, float((a.double() - r).abs().max()), scale)


def test_linear_strided_rows_and_empty():
    import wmrl.functional as F
    x = torch.randn(300, 50, device="cuda")[:, :40]          # non-contiguous -> made contiguous by the wrapper
    w = torch.randn(24, 40, device="cuda")
    torch.testing.assert_close(F.linear(x, w), x @ w.t(), rtol=1e-5, atol=1e-5)
    assert F.linear(torch.empty(0, 40, device="cuda"), w).shape == (0, 24)
