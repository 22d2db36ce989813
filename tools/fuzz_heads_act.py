"""Randomised shape sweep of the round-2 kernels against torch / the oracle (not part of the test suite: minutes of GPU time).
    python tools/fuzz_heads_act.py [seconds]
MLP heads on the bf16 chain (forward vs the bf16-emulating oracle, d x vs autograd through it), the acting-step kernel (vs the
oracle restatement), get_features (vs torch.cat).  Prints the worst deviations and any failing shape."""
import os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "world-model-rl_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import wmrl
import wmrl.functional as Fn
import helpers_gpu as HG
from oracle import rssm_oracle as O

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = random.Random(1234)
t_end = time.time() + budget
worst = {"mlp_fwd": 0.0, "mlp_gx_rel": 0.0, "act": 0.0}
n = {"mlp": 0, "act": 0, "feat": 0}
fails = []
ACT = [(torch.nn.ELU, "elu"), (torch.nn.SiLU, "silu"), (torch.nn.ReLU, "relu")]
while time.time() < t_end:
    kind = rng.choice(["mlp", "mlp", "act", "feat"])
    try:
        if kind == "mlp":
            rows = rng.choice([1, 3, 127, 128, 129, 300, 1000, 4097, 20000])
            F_ = rng.choice([1, 7, 16, 40, 230, 257, 1030])
            Hh = rng.choice([32, 64, 96, 256, 320, 512])
            L = rng.randint(1, 4)
            Oo = rng.choice([1, 2, 5, 16])
            ai = rng.randrange(3)
            layers, w = [], F_
            for _ in range(L):
                layers += [torch.nn.Linear(w, Hh), torch.nn.LayerNorm(Hh), ACT[ai][0]()]
                w = Hh
            seq = torch.nn.Sequential(*layers, torch.nn.Linear(Hh, Oo)).cuda()
            x = torch.randn(rows, F_, device="cuda", requires_grad=True)
            y = Fn.mlp_forward_bf16(seq, x)
            gy = torch.randn_like(y)
            y.backward(gy)
            sd = {k: v.detach().cpu() for k, v in seq.state_dict().items()}
            xe = x.detach().cpu().requires_grad_(True)
            ye = O.mlp_head(sd, xe, ACT[ai][1], emu=True)
            ye.backward(gy.cpu())
            scale = max(1.0, float(ye.abs().max()))
            e1 = float((y.detach().cpu() - ye.detach()).abs().max()) / scale
            gr = xe.grad
            e2 = float((x.grad.cpu() - gr).norm() / (gr.norm() + 1e-30))
            worst["mlp_fwd"] = max(worst["mlp_fwd"], e1); worst["mlp_gx_rel"] = max(worst["mlp_gx_rel"], e2)
            n["mlp"] += 1
            # depth-4 chains reach 4-5e-3 of the output scale vs the emulating oracle (bf16 flips of single activations); a
            # handful of rows has no averaging in the gradient norm
            if e1 > 8e-3 or (rows >= 100 and e2 > 4e-2):
                fails.append(("mlp", rows, F_, Hh, L, Oo, ACT[ai][1], e1, e2))
        elif kind == "act":
            variant = rng.choice(["continuous", "discrete"])
            D = rng.choice([17, 32, 200, 257]); S = rng.choice([4, 8, 16, 30, 32]) if variant == "continuous" else rng.choice([4, 8, 16, 32])
            C = 1 if variant == "continuous" else rng.choice([1, 3, 8, 32])
            Hd = rng.choice([19, 64, 200]); E = rng.choice([5, 64, 130]); A = rng.choice([1, 3, 6]); Ha = rng.choice([32, 96, 512]); La = rng.randint(1, 3)
            B = rng.choice([1, 2, 5])
            rssm, actor = HG.build_modules(variant, D, S, C, Hd, E, A, Ha, La, bound=[0.5 + 0.1 * i for i in range(A)], seed=rng.randrange(1000))
            w, aw = HG.cpu_weights(rssm), HG.cpu_weights(actor)
            g = torch.Generator().manual_seed(rng.randrange(10 ** 6))
            deter = torch.tanh(torch.randn(B, D, generator=g))
            S_in = S * C if variant == "discrete" else S
            state = HG.init_state(rssm, deter, torch.zeros(B, S_in))
            e = torch.randn(B, E, generator=g)
            if variant == "continuous":
                n1, n2 = torch.randn(B, S, generator=g), torch.randn(B, S, generator=g)
            else:
                n1, n2 = (torch.empty(B, C, S).exponential_(1.0, generator=g) for _ in range(2))
            ref = O.act_step(variant, w, aw, e, deter, n1, n2, S, C, actor.action_scale, torch.tensor([0.5 + 0.1 * i for i in range(A)]))
            a, nxt, post = rssm.act_step(e.cuda(), state, actor, noise_post=n1.cuda(), noise_prior=n2.cuda())
            if variant == "discrete" and (HG.near_tie_mask(ref["post_a"], n1).any() or HG.near_tie_mask(ref["prior_a"], n2).any()):
                continue
            err = max(float((a.cpu() - ref["action"]).abs().max()), float((nxt.deter_state.cpu() - ref["deter_next"]).abs().max()),
                      float((nxt.stoch_state.cpu() - ref["prior_stoch"].detach()).abs().max()))
            worst["act"] = max(worst["act"], err)
            n["act"] += 1
            if err > 2e-4:
                fails.append(("act", variant, D, S, C, Hd, E, A, Ha, La, B, err))
        else:
            B, T = rng.choice([1, 7, 300]), rng.choice([1, 2, 16])
            D, S = rng.choice([1, 27, 200, 1024]), rng.choice([1, 7, 30, 1024])
            d, s = torch.randn(B, T, D, device="cuda"), torch.randn(B, T, S, device="cuda")
            n["feat"] += 1
            if not torch.equal(Fn.sequence_features(d, s), torch.cat((d[:, 1:], s[:, 1:]), -1)):
                fails.append(("feat", B, T, D, S))
    except Exception as ex:   # noqa: BLE001
        fails.append((kind, "exception", repr(ex)[:200]))
        if len(fails) > 5:
            break
print("FAILS:" if fails else "no failures", *fails, sep="\n")
print("cases", n, "worst", worst)
