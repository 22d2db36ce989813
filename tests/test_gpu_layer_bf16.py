"""GPU parity of the LAYER-WISE bf16 tcgen05 chain (csrc/tc_layer.cu): the discrete-categorical RSSM
(rssm.py:222-227, state/discrete.py:20-37) and the continuous shapes the persistent kernel does not cover.

Discrete rollouts are chaotic in the sampled index (one flipped draw changes the whole future), so the check is a
REPLAY: the kernel runs H steps free, then ``oracle.imagine_rollout(emulate_bf16=True, forced_idx=kernel idx)``
walks the same trajectory with the reference arithmetic (GEMM operands rounded to bf16 where the kernel rounds
them, fp32 accumulate) and

  * deter / logits / actions agree to ``TOL`` absolute at every step (all fields are O(1));
  * the stochastic state is EXACTLY one_hot(idx) (forward value of the straight-through sample);
  * the oracle's own draw argmax(p/q) from ITS logits equals the kernel's index except at near-ties: the
    mismatch rate is <= 1e-3 and every mismatch has a top-2 margin of p/q below 3 % (bf16 logits noise).
"""
import pytest
import torch

from oracle import rssm_oracle as O

pytestmark = pytest.mark.gpu

TOL, RMS_TOL = 6e-3, 1.5e-3


def _setup(variant, D, S, C, Hd, A, Ha, La, seed=3):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import helpers_gpu as HG
    from wmrl import _lib
    if not _lib.load().wmrl_device_supports_bf16():
        pytest.skip("not an sm_100 device")
    rssm, actor = HG.build_modules(variant, D, S, C, Hd, 64, A, Ha, La, bound=1.0, seed=seed, precision="bf16")
    return HG, rssm, actor


def _check(name, got, want, tol=TOL):
    err = (got.detach().cpu().float() - want).abs()
    mx, rms = float(err.max()), float(err.square().mean().sqrt())
    assert mx <= tol, f"{name}: max |err| {mx:.3e} > {tol} (per step {err.amax(dim=tuple(i for i in range(err.dim()) if i != 1)).tolist()})"
    assert rms <= RMS_TOL, f"{name}: rms {rms:.3e}"
    return mx


DISC_SHAPES = [
    # D, S(classes), C(groups), Hd, A, Ha, La, B, H
    (32, 8, 4, 32, 1, 64, 2, 200, 10),        # the reference fixture dims (tests/conftest.py:65-74)
    (200, 32, 32, 200, 6, 512, 3, 300, 15),
    (600, 32, 32, 600, 6, 512, 3, 300, 15),   # BASELINE configs[1] dims
    (96, 16, 6, 80, 3, 128, 1, 130, 7),       # P = 96: a ragged 32-column block count, S = 16
]


@pytest.mark.parametrize("shape", DISC_SHAPES)
def test_discrete_replay_vs_bf16_oracle(shape):
    D, S, C, Hd, A, Ha, La, B, H = shape
    HG, rssm, actor = _setup("discrete", D, S, C, Hd, A, Ha, La)
    deter0, stoch0, noise = HG.synth_inputs("discrete", B, H, D, S, C)
    init = HG.init_state(rssm, deter0, stoch0)
    from wmrl.functional import ImagineRollout  # noqa: F401
    with torch.no_grad():
        seq, actions = rssm.imagine_rollout(init, actor, H, noise=noise.cuda(), return_actions=True)
    # indices: recover from the one-hot output (and check it IS one-hot)
    st = seq.stoch_states[:, 1:].reshape(B, H, C, S).cpu()
    assert torch.all((st == 0) | (st == 1)) and torch.all(st.sum(-1) == 1), "stoch is not exactly one-hot"
    idx_k = st.argmax(-1)
    with torch.no_grad():
        ref = O.imagine_rollout("discrete", HG.cpu_weights(rssm), HG.cpu_weights(actor), deter0, stoch0, noise, H, S, C,
                                emulate_bf16=True, bias0_bf16=False, forced_idx=idx_k)
    worst = {"deter": _check("deter", seq.deter_states[:, 1:], ref["deter"]),
             "logits": _check("logits", seq.logits[:, 1:], ref["logits"]),
             "actions": _check("actions", actions, ref["actions"])}
    mism = ref["idx"] != idx_k
    rate = float(mism.float().mean())
    assert rate <= 1e-3, f"index mismatch rate {rate:.2e}"
    if mism.any():
        p = O.categorical_probs(ref["logits"])
        v = p / noise.permute(1, 0, 2, 3)
        top2 = v.topk(2, dim=-1).values
        margin = ((top2[..., 0] - top2[..., 1]) / top2[..., 0])[mism]
        assert float(margin.max()) < 3e-2, f"index mismatch away from a tie: margin {float(margin.max()):.3e}"
    print(shape, worst, "idx mismatch rate", rate)


def test_discrete_matches_fp32_path_at_step_one():
    """teacher-forced single step against the fp32 validation path of the same module (north-star band rtol 1e-2)."""
    D, S, C, Hd, A, Ha, La = 600, 32, 32, 600, 6, 512, 3
    HG, rssm, actor = _setup("discrete", D, S, C, Hd, A, Ha, La)
    B = 260
    deter0, stoch0, noise = HG.synth_inputs("discrete", B, 1, D, S, C)
    init = HG.init_state(rssm, deter0, stoch0)
    with torch.no_grad():
        s16, a16 = rssm.imagine_rollout(init, actor, 1, noise=noise.cuda(), return_actions=True)
        rssm.set_precision("fp32")
        s32, a32 = rssm.imagine_rollout(init, actor, 1, noise=noise.cuda(), return_actions=True)
    for name, x, y in [("deter", s16.deter_states, s32.deter_states), ("logits", s16.logits[:, 1:], s32.logits[:, 1:]),
                       ("actions", a16, a32)]:
        rel = float((x - y).norm() / y.norm())
        assert rel <= 1e-2, f"{name}: relative Frobenius error vs the fp32 path {rel:.3e}"
    agree = float((s16.stoch_states == s32.stoch_states).all(-1).float().mean())
    assert agree >= 0.95, f"rows whose 32 draws all agree with the fp32 path: {agree:.3f}"


CONT_SHAPES = [
    (600, 32, 600, 6, 512, 3, 300, 12),    # too wide for one SM's panels -> layer-wise chain
    (1024, 64, 1024, 4, 256, 2, 140, 6),
]


@pytest.mark.parametrize("shape", CONT_SHAPES)
def test_continuous_free_running_vs_bf16_oracle(shape):
    D, S, Hd, A, Ha, La, B, H = shape
    HG, rssm, actor = _setup("continuous", D, S, 1, Hd, A, Ha, La)
    deter0, stoch0, noise = HG.synth_inputs("continuous", B, H, D, S, 1)
    init = HG.init_state(rssm, deter0, stoch0)
    with torch.no_grad():
        seq, actions = rssm.imagine_rollout(init, actor, H, noise=noise.cuda(), return_actions=True)
        ref = O.imagine_rollout("continuous", HG.cpu_weights(rssm), HG.cpu_weights(actor), deter0, stoch0, noise, H, S,
                                emulate_bf16=True, bias0_bf16=False)
    worst = {n: _check(n, g, ref[k]) for n, g, k in [("deter", seq.deter_states[:, 1:], "deter"),
                                                     ("stoch", seq.stoch_states[:, 1:], "stoch"),
                                                     ("mean", seq.means[:, 1:], "mean"), ("std", seq.stds[:, 1:], "std"),
                                                     ("actions", actions, "actions")]}
    print(shape, worst)


def test_discrete_configs4_dims_and_batch_split():
    """BASELINE configs[4] dims (deter 1024, 32x32, hidden 1024): replay parity on 256 rows, and the rollout of a
    sub-batch is bit-identical to the same rows inside a larger batch (rows are independent; tiles differ)."""
    D, S, C, Hd, A, Ha, La = 1024, 32, 32, 1024, 6, 512, 3
    HG, rssm, actor = _setup("discrete", D, S, C, Hd, A, Ha, La, seed=1)
    B, H = 700, 8
    deter0, stoch0, noise = HG.synth_inputs("discrete", B, H, D, S, C)
    init = HG.init_state(rssm, deter0, stoch0)
    with torch.no_grad():
        seq, actions = rssm.imagine_rollout(init, actor, H, noise=noise.cuda(), return_actions=True)
        sub = slice(128, 384)
        init_s = HG.init_state(rssm, deter0[sub], stoch0[sub])
        seq_s, act_s = rssm.imagine_rollout(init_s, actor, H, noise=noise[:, sub].cuda(), return_actions=True)
    assert torch.equal(seq.deter_states[sub], seq_s.deter_states) and torch.equal(seq.stoch_states[sub], seq_s.stoch_states)
    assert torch.equal(actions[sub], act_s)
    rows = slice(0, 256)
    st = seq.stoch_states[rows, 1:].reshape(256, H, C, S).cpu()
    idx_k = st.argmax(-1)
    with torch.no_grad():
        ref = O.imagine_rollout("discrete", HG.cpu_weights(rssm), HG.cpu_weights(actor), deter0[rows], stoch0[rows],
                                noise[:, rows], H, S, C, emulate_bf16=True, bias0_bf16=False, forced_idx=idx_k)
    _check("deter", seq.deter_states[rows, 1:], ref["deter"])
    _check("logits", seq.logits[rows, 1:], ref["logits"])
    assert float((ref["idx"] != idx_k).float().mean()) <= 1e-3
