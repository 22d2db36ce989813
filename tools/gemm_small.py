"""Time the 3xTF32 GEMM at the 2 500-row shapes of BASELINE configs[0..2] (includes ~20 us of Python/ctypes per call)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "world-model-rl_b200")]
import torch
import wmrl.functional as F
def t(fn, n=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
out = []
for M, N, K in [(2500, 600, 600), (2500, 1800, 600), (2500, 512, 1624), (2500, 512, 512), (2500, 600, 1030), (2500, 1024, 600)]:
    x = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") / K ** 0.5; b = torch.randn(N, device="cuda")
    with torch.no_grad():
        out.append(f"{t(lambda: F.linear(x, w, b, 0))*1e3:6.1f}")
print("wmrl.functional.linear at 2 500 rows, (N, K) = (600,600) (1800,600) (512,1624) (512,512) (600,1030) (1024,600): " + " ".join(out) + " us", flush=True)
