import os, sys
ROOT = "/root/repo"
for p in (ROOT, os.path.join(ROOT, "world-model-rl_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch, helpers_gpu as HG
D, S, Hd, A, Ha, La = 200, 30, 200, 6, 512, 3
rssm, actor = HG.build_modules("continuous", D, S, 1, Hd, 64, A, Ha, La, seed=0, precision="bf16")
B, H = 65536, 15
d0, s0, nz = HG.synth_inputs("continuous", B, H, D, S, 1)
init = HG.init_state(rssm, d0, s0); nz = nz.cuda()
with torch.no_grad():
    for _ in range(2):
        rssm.imagine_rollout(init, actor, H, noise=nz)
torch.cuda.synchronize()
