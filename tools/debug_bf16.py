"""Debug helper (not a test): per-field error of the bf16 tcgen05 path vs the oracle."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "world-model-rl_b200"), os.path.join(ROOT, "tests")]
import torch
import wmrl
import helpers_gpu as HG
from oracle import rssm_oracle as O

def run(B, H, D=200, S=30, Hd=200, A=6, Ha=512, La=3, seed=3):
    dev = "cuda"
    rssm, actor = HG.build_modules("continuous", D, S, 1, Hd, 64, A, Ha, La, bound=1.0, seed=seed, precision="bf16")
    deter0, stoch0, noise = HG.synth_inputs("continuous", B, H, D, S, 1)
    init = HG.init_state(rssm, deter0, stoch0)
    with torch.no_grad():
        seq, actions = rssm.imagine_rollout(init, actor, H, noise=noise.to(dev), return_actions=True)
    torch.cuda.synchronize()
    w, aw = HG.cpu_weights(rssm), HG.cpu_weights(actor)
    with torch.no_grad():
        ref = O.imagine_rollout("continuous", w, aw, deter0, stoch0, noise, H, S, 1, 5.0, torch.tensor(1.0))
    out = {"deter": seq.deter_states[:, 1:], "stoch": seq.stoch_states[:, 1:], "mean": seq.means[:, 1:],
           "std": seq.stds[:, 1:], "actions": actions}
    print(f"B={B} H={H} D={D} S={S} Ha={Ha}")
    for k, v in out.items():
        v = v.cpu()
        err = (v - ref[k]).abs()
        print(f"  {k:8s} max|err| {err.max():.3e} mean|err| {err.mean():.3e} ref|max| {ref[k].abs().max():.3f} "
              f"nan {int(torch.isnan(v).sum())} per-step max {[f'{e:.1e}' for e in err.amax(dim=(0,2)).tolist()]}")
    return out, ref

if __name__ == "__main__":
    run(128, 1)
    run(300, 2)
    run(256, 15)
    run(64, 3, D=32, S=8, Hd=32, A=1, Ha=64, La=2)
