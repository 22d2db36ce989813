import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "world-model-rl_b200")]
import torch
import wmrl.functional as F
M, N, K = (int(a) for a in sys.argv[1:4])
x = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") / K ** 0.5; b = torch.randn(N, device="cuda")
with torch.no_grad():
    for _ in range(3): F.linear(x, w, b, 0)
torch.cuda.synchronize()
