"""The MLP heads (DenseModel reward / done, ValueCritic) on the bf16 tcgen05 chain (wmrl_mlp_fwd_bf16 /
wmrl_mlp_bwd_bf16) against the oracle: forward against the bf16-emulating restatement (tight) and the fp32 modules
(north_star tolerance rtol 1e-2 of the output scale); every gradient - d features, all Linear / LayerNorm parameters
- against torch autograd through the bf16-emulating oracle (cosine >= 0.9995, relative Frobenius error <= 2e-2) and
against fp32 autograd of the module itself.  SiLU / ELU heads meet the same bound in fp32; the ReLU critic only a
looser one (cosine >= 0.99, relative error <= 0.12): a pre-activation within bf16 rounding of zero falls on the
other side of the kink, each flip changes a full-sized gradient element, and ~0.1 % flips per block are ~3 % of
gradient norm - a property of bf16 operands, reproduced exactly by the emulating oracle."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _make(kind, F_):
    import wmrl
    if kind == "critic":
        return wmrl.ValueCritic(F_, num_layers=3, hidden_dim=512, precision="bf16"), "relu", "layers"
    if kind == "done":
        return wmrl.DenseModel(depth=3, input_dim=F_, hidden_dim=256, output_dim=1, output_act=torch.nn.Sigmoid,
                               precision="bf16"), "silu", "fc"
    if kind == "elu2":
        return wmrl.DenseModel(depth=2, input_dim=F_, hidden_dim=96, output_dim=5, hidden_act=torch.nn.ELU,
                               precision="bf16"), "elu", "fc"
    return wmrl.DenseModel(depth=3, input_dim=F_, hidden_dim=256, output_dim=1, precision="bf16"), "silu", "fc"


def _perturb(m, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if p.dim() == 1:   # biases / LayerNorm affine away from their trivial init
                p.add_(torch.randn(p.shape, generator=g) * 0.2)


@pytest.mark.parametrize("kind,F_,B,H", [("reward", 230, 300, 5), ("done", 230, 129, 3), ("critic", 230, 300, 5),
                                         ("critic", 1624, 64, 4), ("elu2", 40, 7, 1)])
def test_forward_and_gradients(kind, F_, B, H):
    from oracle import rssm_oracle as O
    torch.manual_seed(11)
    m, act, attr = _make(kind, F_)
    _perturb(m, 5)
    ref = copy.deepcopy(m)
    ref.precision = "fp32"
    x = torch.randn(B, H, F_)
    # oracle, fp32 on the CPU (the torch module itself is the reference's implementation)
    xr = x.clone().requires_grad_(True)
    seq_r = getattr(ref, attr)
    yr_pre = seq_r(xr)
    yr = ref.output_act(yr_pre) if hasattr(ref, "output_act") else yr_pre
    gy = torch.randn_like(yr)
    yr.backward(gy)
    sd = {k: v.detach() for k, v in seq_r.state_dict().items()}
    y_emu = O.mlp_head(sd, x, act, emu=True)
    torch.testing.assert_close(O.mlp_head(sd, x, act, emu=False), yr_pre.detach(), rtol=1e-5, atol=1e-5)
    # chain
    mc = m.cuda()
    xc = x.cuda().requires_grad_(True)
    yc = mc(xc)
    assert yc.shape == yr.shape
    yc.backward(gy.cuda())
    pre_c = getattr(mc, attr)
    with torch.no_grad():
        import wmrl.functional as Fn
        yc_pre = Fn.mlp_forward_bf16(pre_c, xc.detach())
    scale = float(yr_pre.detach().abs().max())
    assert float((yc_pre.cpu() - y_emu).abs().max()) <= 2e-3 * max(scale, 1.0)
    assert float((yc_pre.cpu() - yr_pre.detach()).abs().max()) <= 1e-2 * max(scale, 1.0)

    def close(a, b, what, cos_min=0.9995, rel_max=2e-2):
        a, b = a.detach().cpu().double().flatten(), b.detach().double().flatten()
        nb = float(b.norm())
        if nb == 0:
            assert float(a.norm()) == 0, what
            return
        cos = float(torch.dot(a, b) / (a.norm() * b.norm()))
        rel = float((a - b).norm() / nb)
        assert cos >= cos_min and rel <= rel_max, (what, cos, rel)

    # autograd through the bf16-emulating oracle (same kink decisions as the chain)
    xe = x.clone().requires_grad_(True)
    sde = {k: v.detach().clone().requires_grad_(True) for k, v in seq_r.state_dict().items()}
    ye_pre = O.mlp_head(sde, xe, act, emu=True)
    ye = ref.output_act(ye_pre) if hasattr(ref, "output_act") else ye_pre
    ye.backward(gy)
    close(xc.grad, xe.grad, "d x (emu)")
    for n, pc in pre_c.named_parameters():
        close(pc.grad, sde[n].grad, n + " (emu)")
    loose = dict(cos_min=0.99, rel_max=0.12) if act == "relu" else {}
    close(xc.grad, xr.grad, "d x", **loose)
    for (n, pc), (_, pr) in zip(mc.named_parameters(), ref.named_parameters()):
        close(pc.grad, pr.grad, n, **loose)


def test_feature_gradient_only_and_shared_panel():
    """frozen head (target critic, reward model inside FreezeParameters): only d features; two heads over the SAME
    feature tensor share one converted panel and their input gradients add up."""
    import wmrl
    import wmrl.functional as Fn
    torch.manual_seed(2)
    F_ = 230
    a = wmrl.ValueCritic(F_, precision="bf16").cuda()
    b = wmrl.DenseModel(input_dim=F_, precision="bf16").cuda()
    for p in list(a.parameters()) + list(b.parameters()):
        p.requires_grad_(False)
    x = torch.randn(260, 4, F_, device="cuda", requires_grad=True)
    ya = a(x)
    panel = Fn._PANELS.panel
    yb = b(x)
    assert Fn._PANELS.panel is panel          # converted once
    (ya.sum() + 2.0 * yb.sum()).backward()
    ar, br = copy.deepcopy(a).cpu(), copy.deepcopy(b).cpu()
    ar.precision = br.precision = "fp32"
    xr = x.detach().cpu().requires_grad_(True)
    (ar.layers(xr).sum() + 2.0 * br.fc(xr).sum()).backward()
    g, gr = x.grad.cpu().flatten().double(), xr.grad.flatten().double()
    assert float(torch.dot(g, gr) / (g.norm() * gr.norm())) >= 0.99     # ReLU critic vs fp32: the kink bound (see header)
    assert float((g - gr).norm() / gr.norm()) <= 0.12
    assert all(p.grad is None for p in a.parameters())
    # a new tensor object is converted afresh even if the allocator hands back the same address
    x2 = torch.randn(260, 4, F_, device="cuda")
    y2 = a(x2)
    torch.testing.assert_close(y2, a(x2.clone()), rtol=0, atol=0)


def test_large_rows_property():
    """983 040 feature rows (configs[3]: 65 536 x 15): the chain is row-wise - a slice of the big batch equals the same
    rows run alone.  Bit for bit when both calls use the same tiling of the 512-wide blocks (a second large call); to
    bf16 rounding (the LayerNorm statistics are summed in a different order) when the small call runs them as CTA pairs."""
    import wmrl
    torch.manual_seed(4)
    m = wmrl.ValueCritic(230, precision="bf16").cuda()
    x = torch.randn(65536, 15, 230, device="cuda")
    with torch.no_grad():
        y = m(x)
        assert y.shape == (65536, 15, 1) and bool(torch.isfinite(y).all())
        sub = x[40000:40003].clone()
        scale = float(y.abs().max())
        assert float((m(sub) - y[40000:40003]).abs().max()) <= 2e-3 * max(scale, 1.0)
        big2 = x[20000:60000].clone()                 # 4 688 row tiles: same (unsplit) tiling as the full call
        torch.testing.assert_close(m(big2)[20000:20003], y[40000:40003], rtol=0, atol=0)


def _trainer_problem(prec, B=192, H=6):
    import wmrl
    torch.manual_seed(21)
    D, S, A = 200, 30, 6
    rssm = wmrl.ContinuousRSSM(D, D, S, 64, A, precision="bf16" if prec == "bf16" else "fp32").cuda()
    actor = wmrl.Actor(D + S, A, 1.0, num_layers=3, hidden_dim=512).cuda()
    wm = wmrl.WorldModel(torch.nn.Identity(), torch.nn.Identity(),
                         wmrl.DenseModel(input_dim=D + S, output_act=torch.nn.Sigmoid, precision=prec),
                         wmrl.DenseModel(input_dim=D + S, precision=prec), rssm).cuda()
    trainer = wmrl.ValueGradTrainer(actor, wmrl.ValueCritic(D + S, precision=prec).cuda())
    g = torch.Generator().manual_seed(3)
    d0 = torch.tanh(torch.randn(B, D, generator=g)).cuda()
    s0 = torch.randn(B, S, generator=g).cuda()
    init = wmrl.InternalStateContinuous(d0, s0, s0.clone(), torch.ones_like(s0))
    noise = torch.randn(H, B, S, generator=g).cuda()
    return wm, trainer, actor, init, noise, H


def test_update_with_detached_critic_input_matches_reference_flow():
    """wmrl.ValueGradTrainer feeds the critic detached states (skips a BPTT whose actor gradients the reference zeroes
    unused, value_trainer.py:153-156 + utils.py:19-21): after one update every actor and critic parameter equals the
    reference flow's (critic on the attached states), fp32 kernels."""
    import wmrl
    results = []
    for attached in (False, True):
        wm, trainer, actor, init, noise, H = _trainer_problem("fp32")
        if attached:
            def value_loss(state_samples, reward_samples, done_samples, tr=trainer):
                values = tr.critic(state_samples)
                with wmrl.FreezeParameters([tr.target_critic]):
                    tv = tr.value_targets(tr.target_critic(state_samples), reward_samples, done_samples)
                return (0.5 * (tv.detach() - values) ** 2).mean(), tv
            trainer.value_loss = value_loss
        roll = wm.imagine_rollout(init, actor, H, noise=noise)
        losses = trainer.update(roll.features, roll.rewards, roll.dones)
        results.append((losses, [p.detach().clone() for p in list(actor.parameters()) + list(trainer.critic.parameters())]))
    (la, pa), (lb, pb) = results
    assert abs(la.value_loss - lb.value_loss) <= 1e-6 * max(1.0, abs(lb.value_loss))
    assert abs(la.actor_loss - lb.actor_loss) <= 1e-6 * max(1.0, abs(lb.actor_loss))
    # Adam's first step is lr * g / (|g| + 1e-8): a gradient element that is a cancelling sum of ~1e-8 magnitude turns the
    # fp32-atomics summation order of the weight-gradient kernels into a visible difference of the step.  Such elements
    # are rare; everything else must agree tightly, and nothing may differ by more than one full step.
    lr = max(trainer.actor_lr, trainer.critic_lr)
    for x, y in zip(pa, pb):
        bad = (x - y).abs() > 1e-6 + 1e-4 * y.abs()
        assert float(bad.float().mean()) <= 2e-3, float(bad.float().mean())
        assert float((x - y).abs().max()) <= 2.5 * lr


def test_value_targets_bf16_heads_vs_fp32():
    """lambda-return targets through the bf16 heads vs the fp32 heads on the same rollout (north_star: rtol 1e-2)."""
    import wmrl
    import wmrl.functional as Fn
    wm, trainer, actor, init, noise, H = _trainer_problem("bf16")
    with torch.no_grad():
        roll = wm.imagine_rollout(init, actor, H, noise=noise)
        v16 = trainer.target_critic(roll.features)
        wmrl.set_precision(wm.reward_model, "fp32"); wmrl.set_precision(wm.done_model, "fp32")
        wmrl.set_precision(trainer.target_critic, "fp32")
        r32, dn32 = wm.reward_model(roll.features), wm.done_model(roll.features)
        # the done mask is a 0.5 threshold (value_trainer.py:148-149): both target sets use the fp32 heads' mask, the
        # done probabilities themselves are compared directly
        d = Fn.done_mask(dn32)
        t16 = trainer.value_targets(v16, roll.rewards, d)
        t32 = trainer.value_targets(trainer.target_critic(roll.features), r32, d)
    assert float((roll.dones - dn32).abs().max()) <= 1e-2
    scale = float(t32.abs().max())
    assert float((t16 - t32).abs().max()) <= 1e-2 * max(scale, 1.0)
    wmrl.set_precision(wm, "bf16")
    assert wm.reward_model.precision == "bf16" and wm.dynamic_model.precision == "bf16"
    losses = trainer.update(roll.features, roll.rewards, roll.dones, critic_only=True)
    assert losses.value_loss == losses.value_loss


@pytest.mark.parametrize("rows,F_,Hh,L,O,act", [(300, 230, 512, 3, 1, 2), (129, 40, 96, 2, 5, 0), (5000, 230, 256, 3, 1, 1),
                                                   (1, 1, 32, 1, 16, 0), (257, 17, 64, 8, 1, 1), (40, 2048, 448, 2, 3, 2)])
def test_mlp_c_abi_guard_bands(rows, F_, Hh, L, O, act):
    """Straight through the C ABI with every caller-owned buffer sized EXACTLY as the size queries say and followed by a
    canary region: packed image, feature panel, saved activations, both workspaces, out, g_x - nothing may be written
    past what was asked for; out / g_x agree with the Python-level path."""
    import ctypes as C
    import wmrl.functional as Fn
    from wmrl import _lib
    lib = _lib.load()
    torch.manual_seed(rows + Hh)
    dev = torch.device("cuda")
    acts = [torch.nn.ELU, torch.nn.SiLU, torch.nn.ReLU]
    layers, w = [], F_
    for _ in range(L):
        layers += [torch.nn.Linear(w, Hh), torch.nn.LayerNorm(Hh), acts[act]()]
        w = Hh
    seq = torch.nn.Sequential(*layers, torch.nn.Linear(Hh, O)).to(dev)
    x = torch.randn(rows, F_, device=dev)
    g_out = torch.randn(rows, O, device=dev)
    params = []
    for i in range(L):
        params += [seq[3 * i].weight, seq[3 * i].bias, seq[3 * i + 1].weight, seq[3 * i + 1].bias]
    params += [seq[-1].weight, seq[-1].bias]
    p32 = [p.detach().contiguous() for p in params]
    CAN = 4096

    def guarded(nbytes):
        buf = torch.full((int(nbytes) + CAN,), 0x5A, dtype=torch.uint8, device=dev)
        return buf

    def intact(buf, nbytes):
        return bool((buf[int(nbytes):] == 0x5A).all())

    st = torch.cuda.current_stream().cuda_stream
    d = _lib.MlpDims(rows=rows, in_dim=F_, hidden=Hh, depth=L, out_dim=O, act=act, flags=_lib.MLP_SAVE)
    n_pack, n_panel = lib.wmrl_mlp_packed_bytes(C.byref(d)), lib.wmrl_mlp_panel_bytes(C.byref(d))
    n_saved, n_ws = lib.wmrl_mlp_saved_bytes(C.byref(d)), lib.wmrl_mlp_workspace_bytes(C.byref(d), 0)
    packed, panel, saved, ws = guarded(n_pack), guarded(n_panel), guarded(n_saved), guarded(n_ws)
    out = guarded(rows * O * 4)
    _lib.check(lib.wmrl_mlp_pack(C.byref(d), C.byref(Fn._mlp_ptrs(p32, L)), packed.data_ptr(), st), "pack")
    _lib.check(lib.wmrl_mlp_to_panel(C.byref(d), x.data_ptr(), panel.data_ptr(), st), "to_panel")
    _lib.check(lib.wmrl_mlp_fwd_bf16(C.byref(d), packed.data_ptr(), panel.data_ptr(), out.data_ptr(), saved.data_ptr(), n_saved,
                                     ws.data_ptr(), n_ws, st), "fwd")
    db = _lib.MlpDims(rows=rows, in_dim=F_, hidden=Hh, depth=L, out_dim=O, act=act, flags=_lib.MLP_WGRAD)
    n_wsb = lib.wmrl_mlp_workspace_bytes(C.byref(db), 1)
    wsb, g_x = guarded(n_wsb), guarded(rows * F_ * 4)
    grads = [torch.zeros_like(p) for p in p32]
    _lib.check(lib.wmrl_mlp_bwd_bf16(C.byref(db), packed.data_ptr(), panel.data_ptr(), saved.data_ptr(), g_out.data_ptr(),
                                     C.byref(Fn._mlp_ptrs(grads, L)), g_x.data_ptr(), wsb.data_ptr(), n_wsb, st), "bwd")
    torch.cuda.synchronize()
    for name, buf, n in (("packed", packed, n_pack), ("panel", panel, n_panel), ("saved", saved, n_saved), ("ws", ws, n_ws),
                         ("out", out, rows * O * 4), ("ws_bwd", wsb, n_wsb), ("g_x", g_x, rows * F_ * 4)):
        assert intact(buf, n), f"{name}: written past its {n} bytes"
    y_abi = out[:rows * O * 4].view(torch.float32).reshape(rows, O)
    gx_abi = g_x[:rows * F_ * 4].view(torch.float32).reshape(rows, F_)
    xr = x.clone().requires_grad_(True)
    y_py = Fn.mlp_forward_bf16(seq, xr)
    y_py.backward(g_out)
    torch.testing.assert_close(y_abi, y_py.detach(), rtol=0, atol=0)
    torch.testing.assert_close(gx_abi, xr.grad, rtol=0, atol=0)
    for gnat, p in zip(grads, params):
        torch.testing.assert_close(gnat, p.grad, rtol=2e-3, atol=2e-3 * float(p.grad.abs().max()))   # fp32 atomics: order
