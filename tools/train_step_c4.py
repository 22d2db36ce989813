"""Timing of the bf16 training step at BASELINE configs[3]'s top point (continuous D200/S30, Actor 3x512,
65 536 start states x H 15): forward (plain / recording) and forward + BPTT through the public module API.

    python tools/train_step_c4.py [--B 65536] [--H 15] [--iters 5] [--precision bf16|fp32]
"""
from __future__ import annotations

import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "world-model-rl_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402



def _feed_gradient(features, g):
    """back-propagate with dL/d features = g (the gradient a consumer - heads, losses - would hand back): identical to
    (features * g).sum().backward() without the harness's own product and reduction over the feature tensor."""
    features.backward(g.expand_as(features))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=65536)
    ap.add_argument("--H", type=int, default=15)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--precision", default="bf16")
    a = ap.parse_args()
    import helpers_gpu as HG
    D, S, Hd, A, Ha, La = 200, 30, 200, 6, 512, 3
    rssm, actor = HG.build_modules("continuous", D, S, 1, Hd, 64, A, Ha, La, seed=0, precision=a.precision)
    for p in rssm.parameters():
        p.requires_grad_(False)
    d0, s0, nz = HG.synth_inputs("continuous", a.B, a.H, D, S, 1)
    init = HG.init_state(rssm, d0, s0)
    nz = nz.cuda()
    gf = torch.randn(a.B, a.H, D + S, device="cuda") / a.B

    def timed(fn, n):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    def fwd_plain():
        with torch.no_grad():
            rssm.imagine_rollout(init, actor, a.H, noise=nz)

    def fwd_rec():
        rssm.imagine_rollout(init, actor, a.H, noise=nz)

    def fwd_bwd():
        actor.zero_grad(set_to_none=True)
        seq = rssm.imagine_rollout(init, actor, a.H, noise=nz)
        _feed_gradient(seq.get_features(), gf)

    t0, t1, t2 = timed(fwd_plain, a.iters), timed(fwd_rec, a.iters), timed(fwd_bwd, a.iters)
    steps = a.B * a.H
    fl_f, fl_fb = 1888640, 4832000
    print(f"precision {a.precision}  B {a.B} H {a.H}")
    print(f"forward (no grad)        {t0:8.3f} ms  {steps / t0 * 1e3:.3e} latent-steps/s  {fl_f * steps / t0 / 1e9:.1f} TF/s")
    print(f"forward (recording)      {t1:8.3f} ms")
    print(f"forward + BPTT backward  {t2:8.3f} ms  {steps / t2 * 1e3:.3e} latent-steps/s  {fl_fb * steps / t2 / 1e9:.1f} TF/s "
          f"({fl_fb * steps / t2 / 1e9 / 1628.3 * 100:.1f}% of 1628.3 burst)")
    print(f"peak memory {torch.cuda.max_memory_allocated() / 1e9:.1f} GB")


if __name__ == "__main__":
    main()
