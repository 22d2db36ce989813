"""Field-by-field check of the bf16 training path (forward activation stash -> persistent BPTT kernel ->
weight-gradient GEMMs) against an fp32 torch-autograd restatement with every intermediate exposed.

    python tools/bptt_debug.py [--B 300] [--H 5] [--shape c1|small|mid]

Prints, per quantity, max |err| and relative Frobenius error: the decoded stash fields (x-hat, y, gates, x, hid,
feat), the start-state gradients, the decoded gradient-stash fields (g_pre, g_u, g_mu) and the actor parameter
gradients.  It drives the C ABI directly (wmrl_imagine_fwd / wmrl_imagine_bwd) so the raw buffers can be read.
"""
from __future__ import annotations

import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "world-model-rl_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import wmrl  # noqa: E402
from wmrl import _lib  # noqa: E402
from wmrl import functional as WF  # noqa: E402

SHAPES = {"c1": (200, 30, 200, 6, 512, 3), "small": (32, 8, 32, 1, 64, 2), "mid": (120, 24, 120, 4, 256, 3),
          "wide_hd": (48, 16, 96, 3, 128, 1)}


def ceil16(x):
    return (x + 15) // 16 * 16


def stash_layout(Dp, Sp, Ha, Hdp, La):
    """mirror of make_stash_layout (csrc/tc_plan.cuh)"""
    o = 0
    L = {}

    def take(name, cols):
        nonlocal o
        L[name] = (o, cols)
        o += cols // 8 * 1024

    take("feat", Dp + Sp)
    for l in range(La):
        take(f"y{l}", Ha)
    for l in range(La):
        take(f"xh{l}", Ha)
    take("xe", Dp)
    for n in ("gr", "gz", "gn", "ghn"):
        take(n, Dp)
    take("hid", Hdp)
    L["rstd"] = (o, 0)
    o += (La * 64 * 4 + 1023) // 1024 * 1024
    blk_f = o
    o = 0
    G = {}

    def takeg(name, cols):
        nonlocal o
        G[name] = (o, cols)
        o += cols // 8 * 1024

    for l in range(La):
        takeg(f"gp{l}", Ha)
    for l in range(La):
        takeg(f"gu{l}", Ha)
    takeg("gmu", 16)
    return L, blk_f, G, o


def decode(buf_u8, blk_bytes, off, cols, B, H):
    """stash field -> [B, H, cols] fp32"""
    ntiles = (B + 127) // 128
    nblk = ntiles * 2 * H
    raw = buf_u8[: nblk * blk_bytes].view(nblk, blk_bytes)[:, off: off + cols // 8 * 1024].contiguous()
    x = raw.view(torch.bfloat16).view(nblk, cols // 8, 64, 8).permute(0, 2, 1, 3).reshape(ntiles, 2, H, 64, cols)
    x = x.permute(0, 1, 3, 2, 4).reshape(ntiles * 128, H, cols)
    return x[:B].float()


def decode_rstd(buf_u8, blk_bytes, off, La, B, H):
    ntiles = (B + 127) // 128
    nblk = ntiles * 2 * H
    raw = buf_u8[: nblk * blk_bytes].view(nblk, blk_bytes)[:, off: off + La * 64 * 4].contiguous()
    x = raw.view(torch.float32).view(ntiles, 2, H, La, 64).permute(0, 1, 4, 2, 3).reshape(ntiles * 128, H, La)
    return x[:B]


def report(name, got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    err = (got - ref).abs().max().item()
    rel = ((got - ref).norm() / (ref.norm() + 1e-20)).item()
    cos = (got.flatten() @ ref.flatten() / (got.norm() * ref.norm() + 1e-30)).item()
    print(f"  {name:24s} max|err| {err:9.3e}  rel {rel:9.3e}  cos {cos:.6f}  |ref| {ref.abs().max().item():.3e}")
    return rel, cos


def reference(rssm, actor, deter0, stoch0, noise, H, gouts):
    """fp32 torch restatement of rssm.py:123-140 with intermediates retained; returns dict of tensors / grads"""
    D, S = rssm.deter_size, rssm.stoch_size
    La = actor.num_layers
    dev = deter0.device
    deter0 = deter0.clone().requires_grad_(True)
    stoch0 = stoch0.clone().requires_grad_(True)
    keep = {k: [] for k in ("feat", "xe", "gr", "gz", "gn", "ghn", "hid", "mu_pre", "deter", "stoch", "mean", "std", "act")}
    for l in range(La):
        keep[f"pre{l}"], keep[f"xh{l}"], keep[f"u{l}"], keep[f"y{l}"], keep[f"rstd{l}"] = [], [], [], [], []
    h, s = deter0, stoch0
    w = dict(rssm.named_parameters())
    for t in range(H):
        feat = torch.cat([h, s], -1).detach()
        keep["feat"].append(feat)
        x = feat
        for l in range(La):
            lin, ln = actor.layers[3 * l], actor.layers[3 * l + 1]
            pre = F.linear(x, lin.weight, lin.bias)
            pre.retain_grad()
            mean = pre.mean(-1, keepdim=True)
            var = pre.var(-1, unbiased=False, keepdim=True)
            rstd = torch.rsqrt(var + 1e-5)
            xh = (pre - mean) * rstd
            u = xh * ln.weight + ln.bias
            u.retain_grad()
            x = F.elu(u)
            keep[f"pre{l}"].append(pre); keep[f"xh{l}"].append(xh); keep[f"u{l}"].append(u); keep[f"y{l}"].append(x)
            keep[f"rstd{l}"].append(rstd[:, 0])
        mu_pre = F.linear(x, actor.mu.weight, actor.mu.bias)
        mu_pre.retain_grad()
        keep["mu_pre"].append(mu_pre)
        a = torch.tanh(mu_pre / actor.action_scale) * actor.bound.to(dev)
        keep["act"].append(a)
        xe = F.elu(F.linear(torch.cat([s, a], -1), w["fc_state_action_embed.weight"], w["fc_state_action_embed.bias"]))
        gi = F.linear(xe, w["rnn.weight_ih"], w["rnn.bias_ih"])
        gh = F.linear(h, w["rnn.weight_hh"], w["rnn.bias_hh"])
        i_r, i_z, i_n = gi.chunk(3, -1)
        h_r, h_z, h_n = gh.chunk(3, -1)
        r, z = torch.sigmoid(i_r + h_r), torch.sigmoid(i_z + h_z)
        n = torch.tanh(i_n + r * h_n)
        h = (1 - z) * n + z * h
        hid = F.elu(F.linear(h, w["state_prior.0.weight"], w["state_prior.0.bias"]))
        out = F.linear(hid, w["state_prior.2.weight"], w["state_prior.2.bias"])
        mean_, raw = out.split(S, -1)
        std = F.softplus(raw) + 0.1
        s = mean_ + std * noise[t]
        for k, v in (("xe", xe), ("gr", r), ("gz", z), ("gn", n), ("ghn", h_n), ("hid", hid), ("deter", h), ("stoch", s),
                     ("mean", mean_), ("std", std)):
            keep[k].append(v)
    seqs = {k: torch.stack(keep[k], 1) for k in ("deter", "stoch", "mean", "std")}
    loss = sum((seqs[k] * gouts[k]).sum() for k in seqs)
    loss.backward()
    out = {k: torch.stack(v, 1) for k, v in keep.items()}
    for l in range(La):
        out[f"gp{l}"] = torch.stack([p.grad for p in keep[f"pre{l}"]], 1)
        out[f"gu{l}"] = torch.stack([p.grad for p in keep[f"u{l}"]], 1)
    out["gmu"] = torch.stack([p.grad for p in keep["mu_pre"]], 1)
    out["g_deter0"], out["g_stoch0"] = deter0.grad, stoch0.grad
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=300)
    ap.add_argument("--H", type=int, default=5)
    ap.add_argument("--shape", default="c1")
    ap.add_argument("--seed", type=int, default=3)
    args = ap.parse_args()
    import helpers_gpu as HG
    D, S, Hd, A, Ha, La = SHAPES[args.shape]
    B, H = args.B, args.H
    dev = torch.device("cuda")
    rssm, actor = HG.build_modules("continuous", D, S, 1, Hd, 64, A, Ha, La, bound=1.0, seed=args.seed, precision="bf16")
    for p in rssm.parameters():
        p.requires_grad_(False)
    deter0, stoch0, noise = (x.to(dev) for x in HG.synth_inputs("continuous", B, H, D, S, 1))
    g = torch.Generator().manual_seed(9)
    gouts = {"deter": torch.randn(B, H, D, generator=g).to(dev) / B, "stoch": torch.randn(B, H, S, generator=g).to(dev) / B,
             "mean": torch.randn(B, H, S, generator=g).to(dev) / B, "std": torch.randn(B, H, S, generator=g).to(dev) / B}
    ref = reference(rssm, actor, deter0, stoch0, noise, H, gouts)
    ref_grads = {k: p.grad.clone() for k, p in actor.named_parameters() if p.grad is not None}
    actor.zero_grad()

    # ---- raw C-ABI forward with the stash
    lib = _lib.load()
    cfg = rssm._config(A, Ha, La, actor.action_scale)
    from wmrl.rssm import _actor_bound, _actor_tensors
    _, _, actor_p = _actor_tensors(actor)
    rssm_p = rssm._rssm_tensors()
    bound = _actor_bound(actor, A, dev)
    packed = rssm._packed.get(cfg, rssm_p, actor_p, bound)
    d = cfg.dims(B, H, _lib.FLAG_SAVE_FOR_BWD, _lib.BF16)
    assert lib.wmrl_bf16_train_supported(C.byref(d)), "shape has no bf16 training path"
    saved_n = lib.wmrl_saved_bytes(C.byref(d), 0)
    saved = torch.zeros(saved_n, dtype=torch.uint8, device=dev)
    ws_n = lib.wmrl_workspace_bytes(C.byref(d), 0)
    ws = torch.zeros(ws_n, dtype=torch.uint8, device=dev)
    a32 = [p.detach().float().contiguous() for p in actor_p]
    r32 = [p.detach().float().contiguous() for p in rssm_p]
    T1 = H + 1
    deter_seq, stoch_seq = torch.zeros(B, T1, D, device=dev), torch.zeros(B, T1, S, device=dev)
    means, stds, actions = torch.zeros(B, T1, S, device=dev), torch.zeros(B, T1, S, device=dev), torch.zeros(B, H, A, device=dev)
    w, aw = WF._rssm_ptrs(r32), WF._actor_ptrs(a32, La, bound)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.wmrl_imagine_fwd(C.byref(d), C.byref(w), C.byref(aw), packed.data_ptr(), deter0.data_ptr(),
                                    stoch0.data_ptr(), noise.contiguous().data_ptr(), deter_seq.data_ptr(),
                                    stoch_seq.data_ptr(), means.data_ptr(), stds.data_ptr(), actions.data_ptr(), None,
                                    saved.data_ptr(), saved_n, ws.data_ptr(), ws_n, st), "fwd")
    torch.cuda.synchronize()
    print("forward outputs (bf16 kernel vs fp32 restatement):")
    report("deter", deter_seq[:, 1:], ref["deter"]); report("stoch", stoch_seq[:, 1:], ref["stoch"])
    report("mean", means[:, 1:], ref["mean"]); report("std", stds[:, 1:], ref["std"]); report("actions", actions, ref["act"])
    # no-save forward must be bit-identical
    d0 = cfg.dims(B, H, 0, _lib.BF16)
    ds2, ss2 = torch.zeros_like(deter_seq), torch.zeros_like(stoch_seq)
    m2, s2, a2 = torch.zeros_like(means), torch.zeros_like(stds), torch.zeros_like(actions)
    _lib.check(lib.wmrl_imagine_fwd(C.byref(d0), C.byref(w), C.byref(aw), packed.data_ptr(), deter0.data_ptr(),
                                    stoch0.data_ptr(), noise.contiguous().data_ptr(), ds2.data_ptr(), ss2.data_ptr(),
                                    m2.data_ptr(), s2.data_ptr(), a2.data_ptr(), None, None, 0, ws.data_ptr(), ws_n, st), "fwd0")
    torch.cuda.synchronize()
    print("  recording forward == plain forward:", bool(torch.equal(ds2, deter_seq) and torch.equal(ss2, stoch_seq)
                                                        and torch.equal(a2, actions)))
    Dp, Sp, Hdp = ceil16(D), ceil16(S), ceil16(Hd)
    Lf, blk_f, Lg, blk_g = stash_layout(Dp, Sp, Ha, Hdp, La)
    print("forward stash fields:")
    feat = decode(saved, blk_f, Lf["feat"][0], Dp + Sp, B, H)
    report("feat.h", feat[..., :D], ref["feat"][..., :D]); report("feat.s", feat[..., Dp:Dp + S], ref["feat"][..., D:])
    for l in range(La):
        report(f"xh{l}", decode(saved, blk_f, Lf[f"xh{l}"][0], Ha, B, H), ref[f"xh{l}"])
        report(f"y{l}", decode(saved, blk_f, Lf[f"y{l}"][0], Ha, B, H), ref[f"y{l}"])
    rs = decode_rstd(saved, blk_f, Lf["rstd"][0], La, B, H)
    for l in range(La):
        report(f"rstd{l}", rs[..., l], ref[f"rstd{l}"])
    for k, width, valid in (("xe", Dp, D), ("gr", Dp, D), ("gz", Dp, D), ("gn", Dp, D), ("ghn", Dp, D), ("hid", Hdp, Hd)):
        report(k, decode(saved, blk_f, Lf[k][0], width, B, H)[..., :valid], ref[k])

    # ---- raw C-ABI backward
    db = cfg.dims(B, H, _lib.FLAG_SAVE_FOR_BWD | _lib.FLAG_BACKWARD, _lib.BF16)
    wsb_n = lib.wmrl_workspace_bytes(C.byref(db), 0)
    wsb = torch.zeros(wsb_n, dtype=torch.uint8, device=dev)
    ga = [torch.zeros_like(p) for p in a32]
    gaw = WF._actor_ptrs(ga, La, None, grads=True)
    g_d0, g_s0 = torch.zeros(B, D, device=dev), torch.zeros(B, S, device=dev)

    def full(g_):   # [B,H,.] -> [B,H+1,.] with a zero index-0 slot
        return torch.cat([torch.zeros_like(g_[:, :1]), g_], 1).contiguous()

    gd, gs, gm, gsd = full(gouts["deter"]), full(gouts["stoch"]), full(gouts["mean"]), full(gouts["std"])
    _lib.check(lib.wmrl_imagine_bwd(C.byref(db), C.byref(w), C.byref(aw), packed.data_ptr(), noise.contiguous().data_ptr(),
                                    deter_seq.data_ptr(), stoch_seq.data_ptr(), means.data_ptr(), stds.data_ptr(),
                                    actions.data_ptr(), saved.data_ptr(), gd.data_ptr(), gs.data_ptr(), gm.data_ptr(),
                                    gsd.data_ptr(), C.byref(gaw), None, g_d0.data_ptr(), g_s0.data_ptr(), wsb.data_ptr(),
                                    wsb_n, st), "bwd")
    torch.cuda.synchronize()
    print("start-state gradients (dgrad chain through the dynamics):")
    report("g_deter0", g_d0, ref["g_deter0"]); report("g_stoch0", g_s0, ref["g_stoch0"])
    print("gradient stash fields:")
    report("gmu", decode(wsb, blk_g, Lg["gmu"][0], 16, B, H)[..., :A], ref["gmu"])
    for l in reversed(range(La)):
        report(f"gu{l}", decode(wsb, blk_g, Lg[f"gu{l}"][0], Ha, B, H), ref[f"gu{l}"])
        report(f"gp{l}", decode(wsb, blk_g, Lg[f"gp{l}"][0], Ha, B, H), ref[f"gp{l}"])
    print("actor parameter gradients (weight-gradient GEMMs / column sums):")
    names = WF._actor_layout(La)
    key = {}
    for l in range(La):
        key[f"lin_w{l}"], key[f"lin_b{l}"] = f"layers.{3 * l}.weight", f"layers.{3 * l}.bias"
        key[f"ln_g{l}"], key[f"ln_b{l}"] = f"layers.{3 * l + 1}.weight", f"layers.{3 * l + 1}.bias"
    key["mu_w"], key["mu_b"] = "mu.weight", "mu.bias"
    worst = 0.0
    for n_, g_ in zip(names, ga):
        rel, cos = report(n_, g_, ref_grads[key[n_]])
        worst = max(worst, rel)
    print("worst relative parameter-gradient error:", worst)


if __name__ == "__main__":
    main()
