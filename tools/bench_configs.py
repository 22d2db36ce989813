"""Measured table over the BASELINE.json configs: GPU path vs the CPU oracle.
Not the contract bench (that is bench.py, C4); this feeds DESIGN.md / profiles/."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "world-model-rl_b200"), os.path.join(ROOT, "tests")]
import torch
import wmrl
import helpers_gpu as HG
from oracle import rssm_oracle as O

dev = "cuda"

def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

def cpu_time(fn, n=2):
    fn(); best = 1e9
    for _ in range(n):
        t = time.perf_counter(); fn(); best = min(best, time.perf_counter() - t)
    return best * 1e3

def imagine_case(name, variant, B, H, D, S, C, Hd, A=6, Ha=512, La=3, bwd=False, precision="fp32"):
    rssm, actor = HG.build_modules(variant, D, S, C, Hd, 64, A, Ha, La, seed=0, precision=precision)
    for p in rssm.parameters(): p.requires_grad_(False)
    d0, s0, nz = HG.synth_inputs(variant, B, H, D, S, C)
    init = HG.init_state(rssm, d0, s0)
    nzd = nz.to(dev)
    S_in = S * (C if variant == "discrete" else 1)
    gfeat = torch.randn(B, H, D + S_in, device=dev) / B
    def gpu():
        if bwd:
            actor.zero_grad(set_to_none=True)
            seq = rssm.imagine_rollout(init, actor, H, noise=nzd)
            (seq.get_features() * gfeat).sum().backward()
        else:
            with torch.no_grad(): rssm.imagine_rollout(init, actor, H, noise=nzd)
    ms = timeit(gpu)
    w, aw = HG.cpu_weights(rssm), HG.cpu_weights(actor)
    Bc = min(B, 2500)
    gf = gfeat[:Bc].cpu()
    def cpu():
        if bwd:
            for v in aw.values(): v.requires_grad_(True); v.grad = None
            r = O.imagine_rollout(variant, w, aw, d0[:Bc], s0[:Bc], nz[:, :Bc], H, S, C)
            (torch.cat([r["deter"], r["stoch"]], -1) * gf).sum().backward()
        else:
            with torch.no_grad(): O.imagine_rollout(variant, w, aw, d0[:Bc], s0[:Bc], nz[:, :Bc], H, S, C)
    cms = cpu_time(cpu)
    return dict(case=name, gpu_ms=ms, gpu_latent_steps_per_s=B * H / ms * 1e3, cpu_ms=cms, cpu_rows=Bc,
                cpu_latent_steps_per_s=Bc * H / cms * 1e3, precision=precision, bwd=bwd)

def observe_case(name, variant, B, T, D, S, C, Hd, E=1024, A=6, bwd=False):
    rssm, _ = HG.build_modules(variant, D, S, C, Hd, E, A, 64, 1, seed=0)
    g = torch.Generator().manual_seed(0)
    obs = torch.relu(torch.randn(B, T, E, generator=g)); act = torch.rand(B, T, A, generator=g) * 2 - 1
    S_in = S * (C if variant == "discrete" else 1)
    mk = (lambda: torch.empty(T - 1, B, C, S).exponential_(1.0, generator=g)) if variant == "discrete" else \
        (lambda: torch.randn(T - 1, B, S, generator=g))
    n1, n2 = mk(), mk()
    obs_d, act_d, n1d, n2d = obs.to(dev), act.to(dev), n1.to(dev), n2.to(dev)
    def gpu():
        if bwd:
            rssm.zero_grad(set_to_none=True)
            pri, post = rssm.observe_rollout(obs_d.requires_grad_(True), act_d, noise=(n1d, n2d))
            (post.get_features().square().mean() + pri.get_features().mean()).backward()
        else:
            with torch.no_grad(): rssm.observe_rollout(obs_d, act_d, noise=(n1d, n2d))
    ms = timeit(gpu)
    w = HG.cpu_weights(rssm)
    def cpu():
        if bwd:
            for v in w.values(): v.requires_grad_(True); v.grad = None
            pri, post = O.observe_rollout(variant, w, obs, act, n1, n2, S, C)
            (torch.cat([post["deter"], post["stoch"]], -1).square().mean() + pri["deter"].mean()).backward()
        else:
            with torch.no_grad(): O.observe_rollout(variant, w, obs, act, n1, n2, S, C)
    cms = cpu_time(cpu)
    steps = B * (T - 1)
    return dict(case=name, gpu_ms=ms, gpu_latent_steps_per_s=steps / ms * 1e3, cpu_ms=cms, cpu_rows=B,
                cpu_latent_steps_per_s=steps / cms * 1e3, precision="fp32", bwd=bwd)

def c3_pipeline_case(name, B=50, T=51, H=15, D=600, S=32, C=32, Hd=600, E=1024, A=6):
    """BASELINE configs[2] end to end (encoder outside the path): posterior filtering over B sequences of length T,
    the B*(T-1) posteriors become the start states (state/discrete.py:81-89) of a horizon-H imagined rollout."""
    variant = "discrete"
    rssm, actor = HG.build_modules(variant, D, S, C, Hd, E, A, 512, 3, seed=0)
    g = torch.Generator().manual_seed(0)
    obs = torch.relu(torch.randn(B, T, E, generator=g)); act = torch.rand(B, T, A, generator=g) * 2 - 1
    mk = lambda n, b: torch.empty(n, b, C, S).exponential_(1.0, generator=g)
    n1, n2, nz = mk(T - 1, B), mk(T - 1, B), mk(H, B * (T - 1))
    obs_d, act_d, n1d, n2d, nzd = (t.to(dev) for t in (obs, act, n1, n2, nz))
    def gpu():
        with torch.no_grad():
            _, post = rssm.observe_rollout(obs_d, act_d, noise=(n1d, n2d))
            rssm.imagine_rollout(post.flatten_batch_time(), actor, H, noise=nzd)
    ms = timeit(gpu)
    w, aw = HG.cpu_weights(rssm), HG.cpu_weights(actor)
    def cpu():
        with torch.no_grad():
            _, post = O.observe_rollout(variant, w, obs, act, n1, n2, S, C)
            O.imagine_rollout(variant, w, aw, post["deter"].reshape(-1, D), post["stoch"].reshape(-1, S * C), nz, H, S, C)
    cms = cpu_time(cpu, n=1)
    steps = B * (T - 1) * H
    return dict(case=name, gpu_ms=ms, gpu_latent_steps_per_s=steps / ms * 1e3, cpu_ms=cms, cpu_rows=B * (T - 1),
                cpu_latent_steps_per_s=steps / cms * 1e3, precision="fp32", bwd=False)

def lambda_case(B, H):
    from wmrl.functional import lambda_return
    V = torch.randn(B, H, 1, device=dev, requires_grad=True); r = torch.randn(B, H, 1, device=dev, requires_grad=True)
    d = (torch.rand(B, H, 1, device=dev) > 0.97).float().cummax(1)[0]
    def gpu(): lambda_return(V, r, d, 0.99, 0.95).mean().backward()
    ms = timeit(gpu, n=20)
    Vc, rc, dc = V.detach().cpu().requires_grad_(True), r.detach().cpu().requires_grad_(True), d.cpu()
    cms = cpu_time(lambda: O.lambda_return_loops(Vc, rc, dc, 0.99, 0.95).mean().backward(), n=1)
    return dict(case=f"lambda-return fwd+bwd B={B} H={H}", gpu_ms=ms, cpu_ms=cms, cpu_rows=B, precision="fp32", bwd=True,
                gpu_latent_steps_per_s=B * H / ms * 1e3, cpu_latent_steps_per_s=B * H / cms * 1e3)

def actor_update_case(name, variant, B, H, D, S, C, Hd, A=6, Ha=512, La=3):
    """The composition pinned by the reference's test_value_rssm.py: imagine from B start states,
    reward / done / target-critic heads on the features, done mask, lambda-return, actor loss = -mean(target),
    backward into the actor.  GPU: wmrl modules end to end.  CPU: oracle rollout + the same torch heads +
    the reference's nested-loop lambda-return (oracle) under autograd."""
    import copy
    rssm, actor = HG.build_modules(variant, D, S, C, Hd, 64, A, Ha, La, seed=0)
    S_in = S * (C if variant == "discrete" else 1)
    F_ = D + S_in
    torch.manual_seed(5)
    reward = wmrl.DenseModel(3, F_, 256, 1).to(dev)
    done = wmrl.DenseModel(3, F_, 256, 1, output_act=torch.nn.Sigmoid).to(dev)
    critic = wmrl.ValueCritic(F_, 3, 512).to(dev)
    heads = (reward, done, critic)
    for m in (rssm,) + heads:
        for p_ in m.parameters(): p_.requires_grad_(False)
    d0, s0, nz = HG.synth_inputs(variant, B, H, D, S, C)
    init = HG.init_state(rssm, d0, s0)
    nzd = nz.to(dev)
    from wmrl.functional import lambda_return, done_mask
    def gpu():
        actor.zero_grad(set_to_none=True)
        feat = rssm.imagine_rollout(init, actor, H, noise=nzd).get_features()
        r, dn, v = reward(feat), done(feat), critic(feat)
        tgt = lambda_return(v, r, done_mask(dn.detach()), 0.99, 0.95)
        (-tgt.mean()).backward()
    ms = timeit(gpu)
    w, aw = HG.cpu_weights(rssm), HG.cpu_weights(actor)
    cheads = [copy.deepcopy(m).cpu() for m in heads]
    Bc = min(B, 2500)
    def cpu():
        for v_ in aw.values(): v_.requires_grad_(True); v_.grad = None
        ro = O.imagine_rollout(variant, w, aw, d0[:Bc], s0[:Bc], nz[:, :Bc], H, S, C)
        feat = torch.cat([ro["deter"], ro["stoch"]], -1)
        r, dn, v = (m(feat) for m in cheads)
        tgt = O.lambda_return_loops(v, r, O.done_preprocess(dn.detach()), 0.99, 0.95)
        (-tgt.mean()).backward()
    cms = cpu_time(cpu, n=1)
    return dict(case=name, gpu_ms=ms, gpu_latent_steps_per_s=B * H / ms * 1e3, cpu_ms=cms, cpu_rows=Bc,
                cpu_latent_steps_per_s=Bc * H / cms * 1e3, precision="fp32", bwd=True)

torch.set_num_threads(os.cpu_count())
rows = [
    imagine_case("C1 continuous D200/S30 B=2500 H=15 fwd fp32", "continuous", 2500, 15, 200, 30, 1, 200),
    imagine_case("C1 continuous B=2500 H=15 fwd bf16-tcgen05", "continuous", 2500, 15, 200, 30, 1, 200, precision="bf16"),
    imagine_case("C1 continuous B=2500 H=15 fwd+bwd fp32", "continuous", 2500, 15, 200, 30, 1, 200, bwd=True),
    imagine_case("C2 discrete D600/32x32 B=2500 H=15 fwd fp32", "discrete", 2500, 15, 600, 32, 32, 600),
    imagine_case("C2 discrete B=2500 H=15 fwd+bwd (actor BPTT) fp32", "discrete", 2500, 15, 600, 32, 32, 600, bwd=True),
    actor_update_case("C1 actor-update step (rollout+heads+lambda-return+bwd) B=2500", "continuous", 2500, 15, 200, 30, 1, 200),
    actor_update_case("C2 actor-update step (rollout+heads+lambda-return+bwd) B=2500", "discrete", 2500, 15, 600, 32, 32, 600),
    observe_case("C3 discrete observe B=50 T=51 E=1024 fwd fp32", "discrete", 50, 51, 600, 32, 32, 600),
    observe_case("C3 discrete observe B=50 T=51 fwd+bwd fp32", "discrete", 50, 51, 600, 32, 32, 600, bwd=True),
    c3_pipeline_case("C3 pipeline: observe 50x51 -> 2500 posteriors -> imagine H=15 (fwd)"),
    imagine_case("C4 continuous B=65536 H=15 fwd bf16-tcgen05", "continuous", 65536, 15, 200, 30, 1, 200, precision="bf16"),
    imagine_case("C4 continuous B=65536 H=15 fwd fp32", "continuous", 65536, 15, 200, 30, 1, 200),
    lambda_case(2500, 15), lambda_case(2500, 64),
]
print(f"{'case':66s} {'GPU ms':>9s} {'GPU steps/s':>12s} {'CPU ms':>9s} {'CPU steps/s':>12s} (CPU: oracle, {os.cpu_count()} threads)")
for r in rows:
    print(f"{r['case']:66s} {r['gpu_ms']:9.3f} {r['gpu_latent_steps_per_s']:12.3e} {r['cpu_ms']:9.1f} {r['cpu_latent_steps_per_s']:12.3e}")
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "configs_r1.json"), "w"), indent=1)
