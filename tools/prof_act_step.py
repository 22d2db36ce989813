import os, sys, cProfile, pstats
ROOT = "/root/repo"
for p in (ROOT, os.path.join(ROOT, "world-model-rl_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch, helpers_gpu as HG
rssm, actor = HG.build_modules("continuous", 200, 30, 1, 200, 64, 6, 512, 3, seed=0)
for p in list(rssm.parameters()) + list(actor.parameters()):
    p.requires_grad_(False)
st = rssm.initial_state(1); st.to("cuda")
e = torch.randn(1, 64, device="cuda")
with torch.no_grad():
    for _ in range(50):
        a, st, _ = rssm.act_step(e, st, actor)
    torch.cuda.synchronize()
    pr = cProfile.Profile(); pr.enable()
    for _ in range(2000):
        a, st, _ = rssm.act_step(e, st, actor)
    torch.cuda.synchronize()
    pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
