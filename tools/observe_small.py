import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "world-model-rl_b200"), os.path.join(ROOT, "tests")]
import torch, helpers_gpu as HG
rssm, _ = HG.build_modules("discrete", 600, 32, 32, 600, 1024, 6, 64, 1, seed=0)
g = torch.Generator().manual_seed(0)
B, T = 50, 4
obs = torch.relu(torch.randn(B, T, 1024, generator=g)).cuda(); act = (torch.rand(B, T, 6, generator=g) * 2 - 1).cuda()
n1 = torch.empty(T - 1, B, 32, 32).exponential_(1.0, generator=g).cuda(); n2 = n1.clone()
with torch.no_grad():
    for _ in range(3):
        rssm.observe_rollout(obs, act, noise=(n1, n2))
torch.cuda.synchronize()
