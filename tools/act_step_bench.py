"""Latency of one online acting step (memory_actor.py:34-57) at B = 1: the eager single-step API (the ATen launch chain
the reference runs) vs the one-kernel path (RSSM.act_step -> wmrl_act_step).  Wall clock per step over a chained loop
with a host read of the action every step (what an environment loop does), and device time between events."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "world-model-rl_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import helpers_gpu as HG

out = {}
for name, (variant, D, S, C, Hd, E, A, Ha, La) in {
        "configs[0] continuous D200/S30": ("continuous", 200, 30, 1, 200, 64, 6, 512, 3),
        "configs[1] discrete D600 32x32": ("discrete", 600, 32, 32, 600, 1024, 6, 512, 3)}.items():
    rssm, actor = HG.build_modules(variant, D, S, C, Hd, E, A, Ha, La, seed=0)
    for p in list(rssm.parameters()) + list(actor.parameters()):
        p.requires_grad_(False)
    st0 = rssm.initial_state(1)
    st0.to("cuda")
    e = torch.randn(1, E, device="cuda")

    def eager(state):
        post = rssm.posterior(e, state)
        a = actor(post.get_features(), deterministic=True)
        return a, rssm.prior(a, post)

    def fused(state):
        a, nxt, _ = rssm.act_step(e, state, actor)
        return a, nxt

    res = {}
    for label, fn in (("eager", eager), ("kernel", fused)):
        with torch.no_grad():
            state = st0
            for _ in range(20):
                a, state = fn(state)
            torch.cuda.synchronize()
            n = 300
            t0 = time.perf_counter()
            for _ in range(n):
                a, state = fn(state)
                a.cpu()                      # the environment consumes the action on the host
            wall = (time.perf_counter() - t0) / n
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                a, state = fn(state)
            e1.record()
            torch.cuda.synchronize()
            res[label] = {"wall_us_per_step": wall * 1e6, "device_us_per_step_back_to_back": e0.elapsed_time(e1) / n * 1e3}
    # kernel alone, warm (weights L2-resident): one captured step replayed from a CUDA graph
    with torch.no_grad():
        nz = rssm.draw_imagine_noise(2, 1, "cuda")
        sstream = torch.cuda.Stream()
        with torch.cuda.stream(sstream):
            rssm.act_step(e, st0, actor, noise_post=nz[0], noise_prior=nz[1])
            sstream.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=sstream):
                for _ in range(20):
                    rssm.act_step(e, st0, actor, noise_post=nz[0], noise_prior=nz[1])
            graph.replay()
            sstream.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(sstream)
            for _ in range(10):
                graph.replay()
            e1.record(sstream)
            sstream.synchronize()
            res["kernel_alone_warm_us"] = e0.elapsed_time(e1) / 200 * 1e3
    res["speedup_wall"] = res["eager"]["wall_us_per_step"] / res["kernel"]["wall_us_per_step"]
    out[name] = res
print(json.dumps(out, indent=1))
