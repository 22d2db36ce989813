"""GPU parity of the bf16 TRAINING path of the layer-wise tcgen05 chain (csrc/tc_layer.cu): forward records the
activation panels, the BPTT walks them backwards (dgrad GEMMs with fused ELU' / GRU-gate / tanh-squash /
LayerNorm+ELU backward epilogues, straight-through categorical backward), actor weight gradients as big-K
tcgen05 GEMMs over the recorded panels (autograd of rssm.py:123-140 with the world model frozen,
world_model.py:172, and the actor input detached, rssm.py:135).

Reference: torch autograd on the fp32 oracle, same weights / inputs / noise.  The discrete oracle REPLAYS the
kernel's sampled indices (``forced_idx``; the straight-through gradient still flows through its own softmax), so
one near-tie flip does not turn the comparison into chaos.  Tolerance (bf16 gradients): per parameter tensor
cosine >= 0.999 and relative Frobenius error <= 2e-2; start-state gradients the same.  The reference's 64-wide
fixture shape gets 3e-2: its gradient panels are rounded to bf16 like everywhere else but every sum averages
over 32-64 terms only (measured 2.2e-2 on one bias vector, cosine 0.9998).
"""
import warnings

import pytest
import torch

from oracle import rssm_oracle as O

pytestmark = pytest.mark.gpu

COS_MIN, REL_MAX = 0.999, 2e-2


def _setup(variant, D, S, C, Hd, A, Ha, La, seed=3):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import helpers_gpu as HG
    from wmrl import _lib
    if not _lib.load().wmrl_device_supports_bf16():
        pytest.skip("not an sm_100 device")
    rssm, actor = HG.build_modules(variant, D, S, C, Hd, 64, A, Ha, La, bound=1.0, seed=seed, precision="bf16")
    for p in rssm.parameters():
        p.requires_grad_(False)
    return HG, rssm, actor


def _agree(name, got, ref, cos_min=COS_MIN, rel_max=REL_MAX):
    got, ref = got.detach().float().cpu().flatten(), ref.detach().float().cpu().flatten()
    if got.numel() < 8:      # a lone scalar (mu.bias at A = 1) is one cancelling sum over B*H rows: no averaging
        rel_max = max(rel_max, 6e-2)
    cos = float(got @ ref / (got.norm() * ref.norm() + 1e-30))
    rel = float((got - ref).norm() / (ref.norm() + 1e-30))
    assert cos >= cos_min and rel <= rel_max, f"{name}: cosine {cos:.6f}, relative error {rel:.3e}"
    return cos, rel


DISC = [
    # D, S, C, Hd, A, Ha, La, B, H
    (32, 8, 4, 32, 1, 64, 2, 200, 7),
    (200, 32, 32, 200, 6, 512, 3, 150, 6),
    (600, 32, 32, 600, 6, 512, 3, 300, 15),     # BASELINE configs[1] dims
    (96, 16, 6, 80, 3, 128, 1, 130, 5),
]


@pytest.mark.parametrize("shape", DISC)
def test_discrete_bptt_vs_oracle_autograd(shape):
    D, S, C, Hd, A, Ha, La, B, H = shape
    HG, rssm, actor = _setup("discrete", D, S, C, Hd, A, Ha, La)
    from wmrl.functional import effective_imagine_precision
    assert effective_imagine_precision(rssm._config(A, Ha, La), B, H, True, False) == "bf16"
    deter0, stoch0, noise = HG.synth_inputs("discrete", B, H, D, S, C)
    g = torch.Generator().manual_seed(9)
    gfeat = torch.randn(B, H, D + S * C, generator=g) / B
    glog = torch.randn(B, H, C, S, generator=g) / B
    init = HG.init_state(rssm, deter0, stoch0)
    init.deter_state.requires_grad_(True)
    init.stoch_state.requires_grad_(True)
    with warnings.catch_warnings():
        warnings.simplefilter("error")     # a silent or loud fp32 downgrade fails the test
        seq = rssm.imagine_rollout(init, actor, H, noise=noise.cuda())
    loss = (seq.get_features() * gfeat.cuda()).sum() + (seq.logits[:, 1:] * glog.cuda()).sum()
    loss.backward()
    idx_k = seq.stoch_states[:, 1:].detach().reshape(B, H, C, S).argmax(-1).cpu()
    w, aw = HG.cpu_weights(rssm), HG.cpu_weights(actor)
    for v in aw.values():
        v.requires_grad_(True)
    d0, s0 = deter0.clone().requires_grad_(True), stoch0.clone().requires_grad_(True)
    ref = O.imagine_rollout("discrete", w, aw, d0, s0, noise, H, S, C, forced_idx=idx_k)
    ((torch.cat([ref["deter"], ref["stoch"]], -1) * gfeat).sum() + (ref["logits"] * glog).sum()).backward()
    named = dict(actor.named_parameters())
    out = {}
    rel_max = 3e-2 if Ha < 128 else REL_MAX
    for k, v in aw.items():
        if v.grad is None:
            assert named[k].grad is None or float(named[k].grad.abs().max()) == 0.0, k
            continue
        assert named[k].grad is not None, k
        out[k] = _agree(k, named[k].grad, v.grad, rel_max=rel_max)
    out["d0"] = _agree("d loss / d deter0", init.deter_state.grad, d0.grad, rel_max=rel_max)
    out["s0"] = _agree("d loss / d stoch0", init.stoch_state.grad, s0.grad, rel_max=rel_max)
    print(shape, {k: (round(c, 6), round(r, 5)) for k, (c, r) in out.items()})


@pytest.mark.parametrize("shape", [(600, 32, 600, 6, 512, 3, 300, 8), (1024, 64, 1024, 4, 256, 2, 140, 5)])
def test_continuous_layer_bptt_vs_oracle_autograd(shape):
    D, S, Hd, A, Ha, La, B, H = shape
    HG, rssm, actor = _setup("continuous", D, S, 1, Hd, A, Ha, La)
    deter0, stoch0, noise = HG.synth_inputs("continuous", B, H, D, S, 1)
    g = torch.Generator().manual_seed(9)
    gfeat = torch.randn(B, H, D + S, generator=g) / B
    gmean, gstd = torch.randn(B, H, S, generator=g) / B, torch.randn(B, H, S, generator=g) / B
    init = HG.init_state(rssm, deter0, stoch0)
    init.deter_state.requires_grad_(True)
    init.stoch_state.requires_grad_(True)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        seq = rssm.imagine_rollout(init, actor, H, noise=noise.cuda())
    ((seq.get_features() * gfeat.cuda()).sum() + (seq.means[:, 1:] * gmean.cuda()).sum()
     + (seq.stds[:, 1:] * gstd.cuda()).sum()).backward()
    w, aw = HG.cpu_weights(rssm), HG.cpu_weights(actor)
    for v in aw.values():
        v.requires_grad_(True)
    d0, s0 = deter0.clone().requires_grad_(True), stoch0.clone().requires_grad_(True)
    ref = O.imagine_rollout("continuous", w, aw, d0, s0, noise, H, S)
    ((torch.cat([ref["deter"], ref["stoch"]], -1) * gfeat).sum() + (ref["mean"] * gmean).sum() + (ref["std"] * gstd).sum()).backward()
    named = dict(actor.named_parameters())
    for k, v in aw.items():
        if v.grad is not None:
            _agree(k, named[k].grad, v.grad)
    _agree("d loss / d deter0", init.deter_state.grad, d0.grad)
    _agree("d loss / d stoch0", init.stoch_state.grad, s0.grad)


def test_discrete_recording_forward_is_bit_identical_and_gradients_add_over_shards():
    D, S, C, Hd, A, Ha, La = 600, 32, 32, 600, 6, 512, 3
    HG, rssm, actor = _setup("discrete", D, S, C, Hd, A, Ha, La)
    B, H = 330, 5
    deter0, stoch0, noise = HG.synth_inputs("discrete", B, H, D, S, C)
    gfeat = (torch.randn(B, H, D + S * C, generator=torch.Generator().manual_seed(2)) / B).cuda()
    init = HG.init_state(rssm, deter0, stoch0)
    seq = rssm.imagine_rollout(init, actor, H, noise=noise.cuda())
    with torch.no_grad():
        seq0 = rssm.imagine_rollout(init, actor, H, noise=noise.cuda())
    assert torch.equal(seq.deter_states.detach(), seq0.deter_states) and torch.equal(seq.stoch_states.detach(), seq0.stoch_states)
    assert torch.equal(seq.logits.detach(), seq0.logits)

    def grads(lo, hi):
        actor.zero_grad(set_to_none=True)
        st = HG.init_state(rssm, deter0[lo:hi], stoch0[lo:hi])
        sq = rssm.imagine_rollout(st, actor, H, noise=noise[:, lo:hi].contiguous().cuda())
        (sq.get_features() * gfeat[lo:hi]).sum().backward()
        return {k: p.grad.clone() for k, p in actor.named_parameters() if p.grad is not None}

    full, a, b = grads(0, B), grads(0, 193), grads(193, B)
    for k in full:
        _agree(k, a[k] + b[k], full[k], cos_min=0.99999, rel_max=1e-3)


def test_layernorm_cluster_forms_agree(monkeypatch):
    """The three tilings of the 512-wide LayerNorm blocks - one CTA per row tile, CTA pairs, 4-CTA clusters (DSMEM
    statistics exchange) - on the same rollout and BPTT: outputs within bf16 rounding of each other (the statistics are
    summed in a different order), gradients cosine >= 0.9999, and each form within the oracle bounds."""
    shape = (600, 32, 600, 6, 512, 3, 300, 6)          # continuous, layer-wise chain (D = 600 is outside the persistent kernel)
    D, S, Hd, A, Ha, La, B, H = shape
    HG, rssm, actor = _setup("continuous", D, S, 1, Hd, A, Ha, La)
    d0, s0, nz = HG.synth_inputs("continuous", B, H, D, S, 1)
    init = HG.init_state(rssm, d0, s0)
    nz = nz.cuda()
    gf = (torch.randn(B, H, D + S, generator=torch.Generator().manual_seed(5)) / B).cuda()
    res = {}
    for form in ("0", "2", "4"):
        monkeypatch.setenv("WMRL_LN_SPLIT", form)
        actor.zero_grad(set_to_none=True)
        seq = rssm.imagine_rollout(init, actor, H, noise=nz)
        feats = seq.get_features()
        feats.backward(gf)
        res[form] = (feats.detach().clone(), {n: p.grad.detach().clone() for n, p in actor.named_parameters() if p.grad is not None})
    monkeypatch.delenv("WMRL_LN_SPLIT")
    scale = float(res["0"][0].abs().max())
    for form in ("2", "4"):
        assert float((res[form][0] - res["0"][0]).abs().max()) <= 4e-3 * max(scale, 1.0), form
        assert not torch.equal(res[form][0], res["0"][0]) or form == "2"      # a different tiling really ran
        assert res[form][1].keys() == res["0"][1].keys() and len(res[form][1]) >= 4 * La + 2
        for n, ga in res[form][1].items():
            _agree(f"{n} (form {form} vs unsplit)", ga, res["0"][1][n], cos_min=0.9999, rel_max=1.5e-2)
    w, aw = HG.cpu_weights(rssm), HG.cpu_weights(actor)
    ref = O.imagine_rollout("continuous", w, aw, d0, s0, nz.cpu(), H, S, 1, 5.0, torch.ones(A), emulate_bf16=True, bias0_bf16=False)
    ref_feats = torch.cat([ref["deter"], ref["stoch"]], -1)
    for form in ("0", "2", "4"):
        assert float((res[form][0].cpu() - ref_feats).abs().max()) <= 6e-3, form
