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


@pytest.mark.parametrize("M,N,K", [(129, 257, 17), (300, 40, 230), (64, 100, 300), (65, 16, 36), (2500, 600, 203), (1, 7, 129)])
def test_linear_c_abi_strided_with_guard_bands(M, N, K):
    """Through the C ABI with row strides larger than the logical widths (how the chain reads / writes the [B,T,.]
    tensors in place) and NaN guard bands around the output: nothing outside Y[M,N] may be written, padding
    columns of X and W must not be read into the result."""
    import ctypes as C
    from wmrl import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(M * 3 + N * 5 + K)
    ldx, ldy, pad_rows = K + 5, N + 9, 3
    xbuf = torch.full((M, ldx), float("nan"), device="cuda")
    x = torch.randn(M, K, generator=g).cuda()
    xbuf[:, :K] = x
    w = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    b = torch.randn(N, generator=g).cuda()
    ybuf = torch.full((M + 2 * pad_rows, ldy), float("nan"), device="cuda")
    y_ptr = ybuf.data_ptr() + pad_rows * ldy * 4
    st = torch.cuda.current_stream().cuda_stream
    for act in (0, 1):
        ybuf.fill_(float("nan"))
        rc = lib.wmrl_linear_fwd(M, N, K, xbuf.data_ptr(), ldx, w.data_ptr(), b.data_ptr(), y_ptr, ldy, act, st)
        assert rc == 0, lib.wmrl_last_error()
        torch.cuda.synchronize()
        y = ybuf[pad_rows:pad_rows + M, :N]
        ref = x.double() @ w.double().t() + b.double()
        if act:
            ref = torch.nn.functional.elu(ref)
        torch.testing.assert_close(y.double(), ref, rtol=1e-4, atol=1e-4)
        assert bool(torch.isnan(ybuf[:pad_rows]).all()) and bool(torch.isnan(ybuf[pad_rows + M:]).all()), "rows outside Y written"
        assert bool(torch.isnan(ybuf[pad_rows:pad_rows + M, N:]).all()), "columns outside Y written"
    # dgrad into a strided dX with guard bands, accumulate on top of existing values
    dy = torch.randn(M, N, generator=g).cuda()
    lddx = K + 4
    dxbuf = torch.full((M + 2, lddx), float("nan"), device="cuda")
    dxbuf[1:M + 1, :K] = 1.0
    rc = lib.wmrl_linear_dgrad(M, N, K, dy.data_ptr(), N, w.data_ptr(), dxbuf.data_ptr() + lddx * 4, lddx, 1, st)
    assert rc == 0, lib.wmrl_last_error()
    torch.cuda.synchronize()
    torch.testing.assert_close(dxbuf[1:M + 1, :K].double(), 1.0 + dy.double() @ w.double(), rtol=1e-4, atol=1e-4)
    assert bool(torch.isnan(dxbuf[0]).all()) and bool(torch.isnan(dxbuf[M + 1]).all()) and bool(torch.isnan(dxbuf[1:M + 1, K:]).all())
    # wgrad accumulates into dW exactly [N, K]
    dwbuf = torch.full((N + 2, K), float("nan"), device="cuda")
    dwbuf[1:N + 1] = 0.0
    rc = lib.wmrl_linear_wgrad(M, N, K, dy.data_ptr(), N, xbuf.data_ptr(), ldx, dwbuf.data_ptr() + K * 4, st)
    assert rc == 0, lib.wmrl_last_error()
    torch.cuda.synchronize()
    ref = dy.double().t() @ x.double()
    torch.testing.assert_close(dwbuf[1:N + 1].double(), ref, rtol=1e-4, atol=1e-4 * float(ref.abs().max()))
    assert bool(torch.isnan(dwbuf[0]).all()) and bool(torch.isnan(dwbuf[N + 1]).all())
