"""One C2 actor-training style step (imagine fwd+bwd, fp32) for per-kernel profiling."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "world-model-rl_b200"), os.path.join(ROOT, "tests")]
import torch, helpers_gpu as HG
variant = sys.argv[1] if len(sys.argv) > 1 else "discrete"
if variant == "discrete":
    D, S, C, Hd, E, A = 600, 32, 32, 600, 1024, 6
else:
    D, S, C, Hd, E, A = 200, 30, 1, 200, 1024, 6
B, H = 2500, 15
rssm, actor = HG.build_modules(variant, D, S, C, Hd, E, A, 512, 3, seed=0)
deter0, stoch0, noise = HG.synth_inputs(variant, B, H, D, S, C, seed=1)
init = HG.init_state(rssm, deter0, stoch0)
noise = noise.cuda()
for p in rssm.parameters(): p.requires_grad_(False)
for it in range(3):
    seq = rssm.imagine_rollout(init, actor, n_steps=H, noise=noise)
    loss = (seq.deter_states ** 2).mean() + (seq.stoch_states ** 2).mean()
    loss.backward()
torch.cuda.synchronize()
