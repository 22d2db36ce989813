"""Time the bf16 imagine forward of the layer-wise tcgen05 chain (csrc/tc_layer.cu) at the BASELINE discrete
configs and print ms / latent-steps/s / TFLOP/s (algorithmic FLOPs of SURVEY 8d).  `python tools/layer_bench.py`."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "world-model-rl_b200"), os.path.join(ROOT, "tests")]
import helpers_gpu as HG  # noqa: E402



def _feed_gradient(features, g):
    """back-propagate with dL/d features = g (the gradient a consumer - heads, losses - would hand back): identical to
    (features * g).sum().backward() without the harness's own product and reduction over the feature tensor."""
    features.backward(g.expand_as(features))


def flops_fwd(D, S_in, P, Hd, A, Ha, La):
    F = D + S_in
    actor = F * Ha + (La - 1) * Ha * Ha + Ha * A
    dyn = (S_in + A) * D + 6 * D * D + D * Hd + Hd * P
    return 2 * (actor + dyn)


def run(name, variant, D, S, C, Hd, A, Ha, La, B, H, iters=5, precision="bf16"):
    rssm, actor = HG.build_modules(variant, D, S, C, Hd, 64, A, Ha, La, precision=precision)
    d0, s0, nz = HG.synth_inputs(variant, B, H, D, S, C)
    init = HG.init_state(rssm, d0, s0)
    nz = nz.cuda()
    with torch.no_grad():
        for _ in range(2):
            rssm.imagine_rollout(init, actor, H, noise=nz)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            rssm.imagine_rollout(init, actor, H, noise=nz)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    S_in = S * C if variant == "discrete" else S
    P = S_in if variant == "discrete" else 2 * S
    fl = flops_fwd(D, S_in, P, Hd, A, Ha, La)
    out = {"case": name, "precision": precision, "ms": round(ms, 3), "latent_steps_per_s": B * H / ms * 1e3,
           "tflops": fl * B * H / ms / 1e9, "flops_per_latent_step": fl}
    print(json.dumps(out), flush=True)
    return out


def flops_bwd(D, S_in, P, Hd, A, Ha, La):
    F = D + S_in
    actor = F * Ha + (La - 1) * Ha * Ha + Ha * A
    dyn = (S_in + A) * D + 6 * D * D + D * Hd + Hd * P
    return 2 * (dyn + 2 * actor - F * Ha)


def run_train(name, variant, D, S, C, Hd, A, Ha, La, B, H, iters=5, precision="bf16"):
    """forward (recording) + actor BPTT: d/d(actor params) of sum(features * g) with the world model frozen."""
    rssm, actor = HG.build_modules(variant, D, S, C, Hd, 64, A, Ha, La, precision=precision)
    for p in rssm.parameters():
        p.requires_grad_(False)
    d0, s0, nz = HG.synth_inputs(variant, B, H, D, S, C)
    init = HG.init_state(rssm, d0, s0)
    nz = nz.cuda()
    S_in = S * C if variant == "discrete" else S
    gfeat = torch.randn(B, 1, D + S_in, device="cuda") / B

    def step():
        actor.zero_grad(set_to_none=True)
        seq = rssm.imagine_rollout(init, actor, H, noise=nz)
        _feed_gradient(seq.get_features(), gfeat)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    P = S_in if variant == "discrete" else 2 * S
    fl = flops_fwd(D, S_in, P, Hd, A, Ha, La) + flops_bwd(D, S_in, P, Hd, A, Ha, La)
    out = {"case": name + " fwd+bwd", "precision": precision, "ms": round(ms, 3), "latent_steps_per_s": B * H / ms * 1e3,
           "tflops": fl * B * H / ms / 1e9, "flops_per_latent_step_fwd_bwd": fl,
           "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 2)}
    print(json.dumps(out), flush=True)
    return out


def run_observe(name, B, T, precision, iters=10):
    """configs[2]: discrete posterior filtering over B sequences of T frames (conv-encoder output E = 1024), no grad."""
    D, S, C, Hd, E, A = 600, 32, 32, 600, 1024, 6
    rssm, _ = HG.build_modules("discrete", D, S, C, Hd, E, A, 64, 1, precision=precision)
    g = torch.Generator().manual_seed(3)
    obs = torch.relu(torch.randn(B, T, E, generator=g)).cuda()
    act = (torch.rand(B, T, A, generator=g) * 2 - 1).cuda()
    with torch.no_grad():
        for _ in range(2):
            rssm.observe_rollout(obs, act)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            rssm.observe_rollout(obs, act)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 10682400       # observe fwd FLOPs per sample-step at C2/C3 dims
    out = {"case": name, "precision": precision, "ms": round(ms, 3), "sample_steps_per_s": B * (T - 1) / ms * 1e3,
           "tflops": fl * B * (T - 1) / ms / 1e9}
    print(json.dumps(out), flush=True)
    return out


def run_observe_train(name, B, T, precision, iters=5):
    """configs[2] world-model side: observe_rollout forward (recording) + full backward (RSSM weight gradients)."""
    D, S, C, Hd, E, A = 600, 32, 32, 600, 1024, 6
    rssm, _ = HG.build_modules("discrete", D, S, C, Hd, E, A, 64, 1, precision=precision)
    g = torch.Generator().manual_seed(3)
    obs = torch.relu(torch.randn(B, T, E, generator=g)).cuda()
    act = (torch.rand(B, T, A, generator=g) * 2 - 1).cuda()
    gf = torch.randn(B, T - 1, D + S * C, device="cuda") / B
    gl = torch.randn(B, T - 1, C, S, device="cuda") / B

    def step():
        rssm.zero_grad(set_to_none=True)
        pri, post = rssm.observe_rollout(obs, act)
        ((post.get_features() * gf).sum() + (pri.logits[:, 1:] * gl).sum() + (post.logits[:, 1:] * gl).sum()).backward()

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 32047200       # observe fwd+bwd (wgrad + dgrad) FLOPs per sample-step
    out = {"case": name + " fwd+bwd", "precision": precision, "ms": round(ms, 3), "sample_steps_per_s": B * (T - 1) / ms * 1e3,
           "tflops": fl * B * (T - 1) / ms / 1e9}
    print(json.dumps(out), flush=True)
    return out


if __name__ == "__main__":
    which = sys.argv[1:] or ["c2", "c5", "c2f32"]
    if "c2" in which:
        run("configs[1] discrete D600 32x32 2500x15", "discrete", 600, 32, 32, 600, 6, 512, 3, 2500, 15, iters=20)
    if "c2f32" in which:
        run("configs[1] discrete D600 32x32 2500x15", "discrete", 600, 32, 32, 600, 6, 512, 3, 2500, 15, iters=10, precision="fp32")
    if "c5" in which:
        run("configs[4] discrete D1024 32x32 8192x64 (one GPU's shard)", "discrete", 1024, 32, 32, 1024, 6, 512, 3, 8192, 64, iters=3)
    if "c2t" in which:
        run_train("configs[1] discrete D600 32x32 2500x15", "discrete", 600, 32, 32, 600, 6, 512, 3, 2500, 15, iters=10)
    if "c2tf32" in which:
        run_train("configs[1] discrete D600 32x32 2500x15", "discrete", 600, 32, 32, 600, 6, 512, 3, 2500, 15, iters=5, precision="fp32")
    if "c5t" in which:
        run_train("configs[4] discrete D1024 32x32 8192x64 (one GPU's shard)", "discrete", 1024, 32, 32, 1024, 6, 512, 3, 8192, 64, iters=2)
    if "c3" in which:
        run_observe("configs[2] discrete observe 50x51 (E 1024) fwd", 50, 51, "bf16")
        run_observe("configs[2] discrete observe 50x51 (E 1024) fwd", 50, 51, "fp32")
    if "c3t" in which:
        run_observe_train("configs[2] discrete observe 50x51 (E 1024)", 50, 51, "bf16")
        run_observe_train("configs[2] discrete observe 50x51 (E 1024)", 50, 51, "fp32")
        run_observe_train("discrete observe 1024x51 (E 1024)", 1024, 51, "bf16")
        run_observe_train("discrete observe 1024x51 (E 1024)", 1024, 51, "fp32", iters=2)
    if "c4" in which:   # continuous bench shape; WMRL_BF16_PATH=layer forces the chain
        run("configs[3] continuous 65536x15", "continuous", 200, 30, 1, 200, 6, 512, 3, 65536, 15, iters=5)
