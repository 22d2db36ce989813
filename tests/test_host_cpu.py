"""CPU-side checks: library loads and exports every symbol include/wmrl.h
declares, host-side module API mirrors the reference, error paths are loud."""
import os
import re

import pytest
import torch

import wmrl
from wmrl import _lib
from wmrl.parallel import shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "wmrl.h")).read()
    declared = set(re.findall(r"WMRL_API\s+[\w\s\*]+?\b(wmrl_\w+)\s*\(", header))
    assert declared, "no prototypes found in include/wmrl.h"
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.wmrl_abi_version() == _lib.ABI_VERSION


def test_dims_struct_matches_header_order():
    header = open(os.path.join(ROOT, "include", "wmrl.h")).read()
    body = header.split("typedef struct wmrl_dims {")[1].split("} wmrl_dims;")[0]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names += [n.strip() for n in decl.split(None, 1)[1].split(",")]
    assert names == [f[0] for f in _lib.Dims._fields_]


def test_size_queries_and_argument_errors_without_gpu():
    import ctypes as C
    lib = _lib.load()
    d = _lib.Dims(variant=0, precision=0, B=8, T=5, D=32, S=8, C=1, A=2, Hd=32, E=16, Ha=64, La=2,
                  action_scale=5.0, flags=0)
    assert lib.wmrl_saved_bytes(C.byref(d), 0) == 4 * 5 * 8 * (2 * (2 * 64 + 1) + 2 + 5 * 32 + 32)
    assert lib.wmrl_workspace_bytes(C.byref(d), 0) > 0
    bad = _lib.Dims(variant=7, precision=0, B=8, T=5, D=32, S=8, C=1, A=2, Hd=32, E=16, Ha=64, La=2,
                    action_scale=5.0, flags=0)
    rc = lib.wmrl_imagine_fwd(C.byref(bad), None, None, *([None] * 11), 0, None, 0, None)
    assert rc != 0 and b"variant" in lib.wmrl_last_error()
    rc = lib.wmrl_lambda_return_fwd(None, None, None, 0.99, 0.95, 4, 0, None, None)
    assert rc != 0


def test_state_dict_keys_match_reference():
    r = wmrl.DiscreteRSSM(32, 32, 8, 4, 64, 1)
    assert list(r.state_dict()) == [
        "rnn.weight_ih", "rnn.weight_hh", "rnn.bias_ih", "rnn.bias_hh", "fc_state_action_embed.weight",
        "fc_state_action_embed.bias", "state_prior.0.weight", "state_prior.0.bias", "state_prior.2.weight",
        "state_prior.2.bias", "state_posterior.0.weight", "state_posterior.0.bias", "state_posterior.2.weight",
        "state_posterior.2.bias"]
    assert r.state_dict()["fc_state_action_embed.weight"].shape == (32, 33)
    a = wmrl.Actor(64, 1, 1, 3, 64)
    assert list(a.state_dict()) == [f"layers.{i}.{k}" for i in (0, 1, 3, 4, 6, 7) for k in ("weight", "bias")] + \
        ["mu.weight", "mu.bias", "stddev.weight", "stddev.bias"]
    c = wmrl.ValueCritic(20, 3, 16)
    assert "layers.9.weight" in c.state_dict()


def test_no_cpu_fallback():
    r, a = wmrl.ContinuousRSSM(16, 16, 4, 8, 1), wmrl.Actor(20, 1, 1.0, 1, 16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        r.imagine_rollout(r.initial_state(2), a, 3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        r.observe_rollout(torch.zeros(2, 3, 8), torch.zeros(2, 3, 1))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        wmrl.functional.lambda_return(torch.zeros(2, 3, 1), torch.zeros(2, 3, 1), torch.zeros(2, 3, 1), 0.9, 0.9)


def test_eager_single_step_api():
    """prior/posterior/observe_step stay callable at B=1 (memory_actor.py:34-57)."""
    torch.manual_seed(0)
    for rssm in (wmrl.ContinuousRSSM(16, 16, 4, 8, 2), wmrl.DiscreteRSSM(16, 16, 4, 3, 8, 2)):
        s = rssm.initial_state(1)
        pri, post = rssm.observe_step(torch.randn(1, 8), torch.randn(1, 2), s)
        assert pri.deter_state.shape == (1, 16) and torch.equal(pri.deter_state, post.deter_state)
        assert post.get_features().shape == (1, 16 + rssm.flat_stoch_size)
    actor = wmrl.Actor(20, 2, [1.0, 2.0], 2, 16)
    wm = wmrl.WorldModel(torch.nn.Linear(5, 8), torch.nn.Linear(20, 5), wmrl.DenseModel(1, 20, 8, 1),
                         wmrl.DenseModel(1, 20, 8, 1), wmrl.ContinuousRSSM(16, 16, 4, 8, 2))
    pol = wmrl.WorldModelActor(wm, actor)
    pol.reset()
    a = pol(torch.randn(1, 5))
    assert a.shape == (1, 2) and float(a[0, 1].abs()) <= 2.0


def test_state_containers_follow_reference_contracts():
    """shape contracts of rssm_world_model/tests/test_continuous_states.py / test_discrete_states.py"""
    s = wmrl.InternalStateContinuous(torch.zeros(3, 5), torch.zeros(3, 2), torch.zeros(3, 2), torch.ones(3, 2))
    seq = s.to_sequence()
    assert seq.shapes == ((3, 1, 5), (3, 1, 2), (3, 1, 2), (3, 1, 2))
    assert seq.get_features().shape == (3, 0, 7)
    seq.append_(s); seq.append_(s)
    assert seq.get_features().shape == (3, 2, 7) and seq[1].mean.shape == (3, 2)
    assert seq.flatten_batch_time().deter_state.shape == (6, 5)
    assert seq.get_dist().sample().shape == (3, 2, 2)
    assert seq.get_last().get_features().shape == (3, 7)
    d = wmrl.InternalStateDiscrete.from_logits(torch.zeros(3, 5), torch.randn(3, 4, 6))
    assert d.stoch_state.shape == (3, 24) and set(d.stoch_state.unique().tolist()) <= {0.0, 1.0}
    dseq = d.to_sequence(); dseq.append_(d)
    assert dseq.get_dist().sample().shape == (3, 1, 4, 6)
    moved = dseq.detach(); assert moved is not dseq and moved.logits.shape == (3, 2, 4, 6)
    assert wmrl.InternalStateContinuous.from_mean_std(torch.zeros(2, 3), torch.zeros(2, 2), torch.ones(2, 2)).stoch_state.shape == (2, 2)


def test_freeze_parameters_restores_flags():
    m = torch.nn.Linear(2, 2)
    m.bias.requires_grad_(False)
    with wmrl.FreezeParameters([m]):
        assert not m.weight.requires_grad and not m.training
    assert m.weight.requires_grad and not m.bias.requires_grad and m.training


def test_shard_range_partitions():
    for n, w in ((10, 4), (65536, 8), (3, 8), (0, 2)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def test_linear_and_ln_entry_points_reject_bad_arguments_without_gpu():
    lib = _lib.load()
    assert lib.wmrl_linear_fwd(4, 0, 8, None, 8, None, None, None, 4, 0, None) != 0          # N = 0
    assert b"shape" in lib.wmrl_last_error()
    assert lib.wmrl_linear_fwd(4, 4, 8, None, 8, None, None, None, 4, 3, None) != 0          # unknown activation
    assert lib.wmrl_linear_fwd(4, 4, 8, None, 8, None, None, None, 4, 0, None) != 0          # NULL tensors
    assert b"NULL" in lib.wmrl_last_error()
    assert lib.wmrl_linear_fwd(0, 4, 8, None, 8, None, None, None, 4, 0, None) == 0          # empty batch: no-op
    assert lib.wmrl_linear_dgrad(4, 4, 8, None, 4, None, None, 8, 0, None) != 0
    assert lib.wmrl_linear_wgrad(4, 4, 8, None, 4, None, 8, None, None) != 0
    assert lib.wmrl_ln_act_fwd(4, 8, None, None, None, 5, None, None, None, None) != 0       # act out of range
    assert lib.wmrl_ln_act_fwd(0, 8, None, None, None, 1, None, None, None, None) == 0
    assert lib.wmrl_ln_act_bwd(4, 8, None, None, None, None, None, 1, None, None, None, None) != 0


def test_heads_run_their_own_modules_on_cpu_tensors():
    """DenseModel / ValueCritic keep the reference's module tree; on CPU tensors (adjacent code, not the path) the
    torch modules themselves run - same numbers as the plain nn.Sequential."""
    torch.manual_seed(0)
    m = wmrl.DenseModel(depth=2, input_dim=12, hidden_dim=16, output_dim=3, output_act=torch.nn.Sigmoid)
    x = torch.randn(5, 7, 12)
    assert torch.equal(m(x), torch.sigmoid(m.fc(x)))
    c = wmrl.ValueCritic(12, num_layers=2, hidden_dim=16)
    assert torch.equal(c(x), c.layers(x))
    from wmrl.functional import mlp_forward
    seq = torch.nn.Sequential(torch.nn.Linear(12, 8), torch.nn.Tanh(), torch.nn.Linear(8, 2))   # not the head pattern
    assert torch.equal(mlp_forward(seq, x), seq(x))


def test_numa_binding_is_best_effort():
    from wmrl.parallel import bind_to_gpu_numa
    import os
    before = os.sched_getaffinity(0)
    assert bind_to_gpu_numa(0) in (True, False)      # no NVML / no GPU here -> False, never raises
    if not torch.cuda.is_available():
        assert os.sched_getaffinity(0) == before


def test_actor_shape_validation_is_loud():
    """an Actor-shaped module that deviates from [Linear, LayerNorm(eps 1e-5), ELU] x L + Linear mu must be
    rejected, not silently mis-evaluated by the fused path (the reference just calls actor(...))."""
    from wmrl.rssm import _actor_tensors, _actor_bound
    good = wmrl.Actor(40, 3, 1.0, 2, 64)
    La, Ha, flat = _actor_tensors(good)
    assert (La, Ha, len(flat)) == (2, 64, 10)
    bad = wmrl.Actor(40, 3, 1.0, 2, 64)
    bad.layers[1] = torch.nn.LayerNorm(64, eps=1e-3)
    with pytest.raises(TypeError, match="LayerNorm"):
        _actor_tensors(bad)
    bad = wmrl.Actor(40, 3, 1.0, 2, 64)
    bad.layers[2] = torch.nn.ReLU()
    with pytest.raises(TypeError, match="ELU"):
        _actor_tensors(bad)
    bad = wmrl.Actor(40, 3, 1.0, 2, 64)
    bad.layers.append(torch.nn.Identity())
    with pytest.raises(TypeError, match="must hold"):
        _actor_tensors(bad)
    # bound: cached expansion, same storage across calls, refreshed by an in-place edit or a re-assignment
    b1 = _actor_bound(good, 3, "cpu")
    assert b1.shape == (3,) and _actor_bound(good, 3, "cpu") is b1
    good.bound.mul_(2.0)
    b2 = _actor_bound(good, 3, "cpu")
    assert b2 is not b1 and float(b2[0]) == 2.0
    good.bound = torch.tensor([1.0, 2.0, 3.0])
    assert _actor_bound(good, 3, "cpu").tolist() == [1.0, 2.0, 3.0]
    with pytest.raises(ValueError):
        _actor_bound(good, 4, "cpu")


def test_compute_rollout_value_shim_matches_reference_loops():
    from oracle import rssm_oracle as O
    g = torch.Generator().manual_seed(0)
    B, H = 5, 7
    V, r = torch.randn(B, H, 1, generator=g), torch.randn(B, H, 1, generator=g)
    d = O.done_preprocess((torch.rand(B, H, 1, generator=g) > 0.8).float())
    tr = wmrl.ValueGradTrainer(wmrl.Actor(4, 1, 1.0, 1, 8), wmrl.ValueCritic(4, 1, 8))
    gam = torch.tensor([0.99 ** i for i in range(H)])
    for k in (1, 3, H, H + 2):
        got = tr.compute_rollout_value(V, r, torch.zeros(B, H, 4), d, k)
        torch.testing.assert_close(got, O._rollout_value(V, r, d, gam, k))


def test_oracle_bf16_emulation_is_a_small_perturbation():
    """the bf16-emulating oracle (checker of the tcgen05 path) stays within the north-star bf16 band of the
    fp32 restatement on the reference fixture, and is exact when nothing needs rounding."""
    from helpers import Golden
    from oracle import rssm_oracle as O
    g = Golden("cont_fixture")
    w, aw = g.weights("rssm."), g.weights("actor.")
    args = ("continuous", w, aw, g.t("deter0"), g.t("stoch0"), g.t("noise_im"), g.H, g.S)
    with torch.no_grad():
        a = O.imagine_rollout(*args)
        b = O.imagine_rollout(*args, emulate_bf16=True)
    for k in ("deter", "mean", "std", "actions"):
        assert float((a[k] - b[k]).norm() / a[k].norm()) < 1e-2, k
    x = torch.tensor([1.0, 0.5, -2.0, 3.140625])
    assert torch.equal(O.bf16_round(x), x)


def test_mlp_abi_size_queries_and_struct_layout():
    """wmrl_mlp_* size queries need no device; struct mirrors follow the header's field order."""
    import ctypes as C
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "wmrl.h")).read()
    body = re.sub(r"/\*.*?\*/", "", header.split("typedef struct wmrl_mlp_dims {")[1].split("} wmrl_mlp_dims;")[0], flags=re.S)
    names = []
    for decl in body.split(";"):
        if decl.strip():
            names += [n.strip() for n in decl.strip().split(None, 1)[1].split(",")]
    assert names == [f[0] for f in _lib.MlpDims._fields_]
    d = _lib.MlpDims(rows=1000, in_dim=230, hidden=512, depth=3, out_dim=1, act=2, flags=0)
    assert lib.wmrl_mlp_packed_bytes(C.byref(d)) > 2 * 2 * (230 * 512 + 2 * 512 * 512)   # forward + transposed operands
    assert lib.wmrl_mlp_panel_bytes(C.byref(d)) >= 8 * 128 * 240 * 2                        # 8 row tiles, in_dim padded to 240
    assert lib.wmrl_mlp_saved_bytes(C.byref(d)) >= 3 * 2 * 8 * 128 * 512 * 2                # y + x-hat per block
    ws0, ws1 = lib.wmrl_mlp_workspace_bytes(C.byref(d), 0), lib.wmrl_mlp_workspace_bytes(C.byref(d), 1)
    d.flags = _lib.MLP_WGRAD
    assert lib.wmrl_mlp_workspace_bytes(C.byref(d), 1) > ws1 > 0 and ws0 > 0                # g_u panels only with WGRAD
    for bad in (dict(hidden=500), dict(hidden=1024), dict(out_dim=17), dict(depth=0), dict(act=3)):
        kw = dict(rows=10, in_dim=230, hidden=512, depth=3, out_dim=1, act=2, flags=0)
        kw.update(bad)
        assert lib.wmrl_mlp_packed_bytes(C.byref(_lib.MlpDims(**kw))) == 0, bad
    rc = lib.wmrl_mlp_fwd_bf16(C.byref(_lib.MlpDims(rows=10, in_dim=8, hidden=500, depth=1, out_dim=1, act=0, flags=0)),
                               None, None, None, None, 0, None, 0, None)
    assert rc != 0 and b"not supported" in lib.wmrl_last_error()


def test_heads_precision_switch_on_cpu():
    """fp32 heads on CPU tensors run the torch modules (reference behaviour); the bf16 chain refuses CPU tensors;
    set_precision reaches every wmrl module under a container."""
    m = wmrl.DenseModel(input_dim=24, hidden_dim=32, depth=2)
    c = wmrl.ValueCritic(24, num_layers=2, hidden_dim=64)
    x = torch.randn(5, 3, 24)
    torch.testing.assert_close(m(x), m.fc(x))
    torch.testing.assert_close(c(x), c.layers(x))
    box = torch.nn.ModuleDict({"r": m, "c": c, "rssm": wmrl.ContinuousRSSM(16, 16, 4, 8, 2)})
    wmrl.set_precision(box, "bf16")
    assert m.precision == c.precision == box["rssm"].precision == "bf16"
    with pytest.raises(RuntimeError, match="CUDA"):
        m(x)
    wmrl.set_precision(box, "fp32")
    assert m.precision == "fp32" and box["rssm"].precision == "fp32"
    with pytest.raises(ValueError):
        wmrl.set_precision(box, "fp8")
    from wmrl.functional import mlp_bf16_supported
    assert mlp_bf16_supported(m.fc) and mlp_bf16_supported(c.layers)
    assert not mlp_bf16_supported(torch.nn.Sequential(torch.nn.Linear(8, 40), torch.nn.LayerNorm(40), torch.nn.ReLU(),
                                                      torch.nn.Linear(40, 1)))          # hidden % 32 != 0
    assert not mlp_bf16_supported(torch.nn.Sequential(torch.nn.Linear(8, 32), torch.nn.Tanh(), torch.nn.Linear(32, 1)))
