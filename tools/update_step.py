"""One actor-critic update at BASELINE configs[3]'s top point through the public API (WorldModel.imagine_rollout ->
ValueGradTrainer.update), heads on the bf16 tcgen05 chain; for launch lists / timing.
    python tools/update_step.py [--B 65536] [--H 15] [--iters 3] [--heads bf16|fp32]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "world-model-rl_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch
import wmrl

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=65536)
ap.add_argument("--H", type=int, default=15)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--heads", default="bf16")
a = ap.parse_args()
import helpers_gpu as HG
D, S, Hd, A, Ha, La = 200, 30, 200, 6, 512, 3
rssm, actor = HG.build_modules("continuous", D, S, 1, Hd, 64, A, Ha, La, seed=0, precision="bf16")
dev = torch.device("cuda")
wm = wmrl.WorldModel(torch.nn.Identity(), torch.nn.Identity(),
                     wmrl.DenseModel(input_dim=D + S, output_act=torch.nn.Sigmoid, precision=a.heads),
                     wmrl.DenseModel(input_dim=D + S, precision=a.heads), rssm).to(dev)
trainer = wmrl.ValueGradTrainer(actor, wmrl.ValueCritic(D + S, precision=a.heads).to(dev))
trainer.target_critic.to(dev)
d0, s0, nz = HG.synth_inputs("continuous", a.B, a.H, D, S, 1)
init = HG.init_state(rssm, d0, s0)


def update():
    roll = wm.imagine_rollout(init, actor, a.H)
    return trainer.update(roll.features, roll.rewards, roll.dones)


for _ in range(2):
    update()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    update()
e1.record()
torch.cuda.synchronize()
print(f"update: {e0.elapsed_time(e1) / a.iters:.2f} ms  peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GB")
