"""Time the Linear primitive (libwmrl GEMM) against torch fp32 / tf32 matmul on a few path shapes."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "world-model-rl_b200")]
import torch
import wmrl.functional as F

def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

shapes = [(65536, 512, 230), (65536, 512, 512), (65536, 600, 200), (65536, 200, 36), (65536, 60, 200),
          (37500, 1800, 600), (37500, 600, 1030), (37500, 1024, 600), (2500, 600, 600), (2500, 1800, 600), (2500, 512, 1624)]
for M, N, K in shapes:
    x = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") / K ** 0.5; b = torch.randn(N, device="cuda")
    with torch.no_grad():
        ours = t(lambda: F.linear(x, w, b, 0))
        torch.backends.cuda.matmul.allow_tf32 = False
        t32 = t(lambda: torch.nn.functional.linear(x, w, b))
        torch.backends.cuda.matmul.allow_tf32 = True
        ttf = t(lambda: torch.nn.functional.linear(x, w, b))
        torch.backends.cuda.matmul.allow_tf32 = False
    fl = 2.0 * M * N * K
    print(f"M={M:6d} N={N:5d} K={K:5d}  wmrl {ours*1e3:8.1f} us {fl/ours/1e9:7.1f} TF/s | cublas fp32 {t32*1e3:8.1f} us {fl/t32/1e9:7.1f} | cublas tf32 {ttf*1e3:8.1f} us {fl/ttf/1e9:7.1f}")
