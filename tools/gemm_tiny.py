import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "world-model-rl_b200")]
import torch
import wmrl.functional as F
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for M, N, K in [(2500, 64, 16), (2500, 64, 64), (2500, 200, 16), (2500, 200, 200), (2500, 600, 200), (2500, 200, 600), (2500, 512, 230), (2500, 512, 512), (2500, 60, 200), (2500, 16, 512), (128, 64, 16), (65, 256, 512)]:
    x = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda"); b = torch.randn(N, device="cuda")
    y = torch.empty(M, N, device="cuda")
    with torch.no_grad():
        ours = t(lambda: F.linear(x, w, b, 0))
        torch.backends.cuda.matmul.allow_tf32 = False
        ref = t(lambda: torch.addmm(b, x, w.t(), out=y))
        el = t(lambda: y.add_(1.0))
    print(f"M={M:5d} N={N:4d} K={K:4d}: wmrl.linear {ours:6.1f} us | torch addmm fp32 {ref:6.1f} us | trivial elementwise kernel {el:5.1f} us", flush=True)
