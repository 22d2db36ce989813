"""GPU parity, bf16 tcgen05 path (continuous RSSM imagine forward).

Tolerance (BASELINE.json north_star): rtol 1e-2 for bf16, checked
TEACHER-FORCED PER STEP: every step starts from the oracle's fp32 state at that
step (the B*H oracle states become B*H independent one-step rollouts), because
a free-running trajectory accumulates bf16 rounding.  Elementwise the bound is
|err| <= 1e-2*|ref| + 5e-3 (the fields are O(1); the floor covers values near
zero) and the aggregate Frobenius relative error must be <= 1e-2.  A
free-running H=15 rollout is additionally bounded more loosely."""
import pytest
import torch

from oracle import rssm_oracle as O

pytestmark = pytest.mark.gpu


def _setup(D, S, Hd, A, Ha, La, seed=3):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import helpers_gpu as HG
    from wmrl import _lib
    if not _lib.load().wmrl_device_supports_bf16():
        pytest.skip("not an sm_100 device")
    rssm, actor = HG.build_modules("continuous", D, S, 1, Hd, 64, A, Ha, La, bound=1.0, seed=seed, precision="bf16")
    return HG, rssm, actor


def _check(name, got, ref, rtol=1e-2, atol=5e-3):
    """elementwise |err| <= rtol*|ref| + atol (all fields are O(1): tanh/sigmoid-bounded state,
    unit-scale latents; atol 5e-3 is 0.5% of that range) AND aggregate relative error
    ||err||_F / ||ref||_F <= rtol."""
    got, ref = got.detach().cpu().float(), ref.detach().float()
    err = (got - ref).abs()
    bad = err > (rtol * ref.abs() + atol)
    assert not bool(bad.any()), f"{name}: {int(bad.sum())} elements outside rtol {rtol} + atol {atol}; " \
                                f"max err {float(err.max()):.3e}"
    rel = float(err.norm() / (ref.norm() + 1e-12))
    assert rel <= rtol, f"{name}: aggregate relative error {rel:.3e} > {rtol}"


SHAPES = {"c1": (200, 30, 200, 6, 512, 3), "small": (32, 8, 32, 1, 64, 2), "wide_hd": (48, 16, 96, 3, 128, 1),
          # A > 8 and Ha % 64 != 0: not covered by the default folded-pair kernel -> automatic fall-back to mode 1
          "fallback": (40, 8, 40, 12, 96, 2)}


@pytest.mark.parametrize("shape", list(SHAPES))
def test_bf16_teacher_forced_per_step(shape):
    D, S, Hd, A, Ha, La = SHAPES[shape]
    HG, rssm, actor = _setup(D, S, Hd, A, Ha, La)
    B, H = 200, 15           # 200 rows: one full 128-row tile + a ragged one per step-batch
    deter0, stoch0, noise = HG.synth_inputs("continuous", B, H, D, S, 1)
    w, aw = HG.cpu_weights(rssm), HG.cpu_weights(actor)
    with torch.no_grad():
        ref = O.imagine_rollout("continuous", w, aw, deter0, stoch0, noise, H, S)
    # oracle states at step t (t=0: the initial state) -> B*H independent start states
    d_in = torch.cat([deter0.unsqueeze(1), ref["deter"][:, :-1]], 1).reshape(B * H, D)
    s_in = torch.cat([stoch0.unsqueeze(1), ref["stoch"][:, :-1]], 1).reshape(B * H, S)
    nz = noise.permute(1, 0, 2).reshape(1, B * H, S)
    init = HG.init_state(rssm, d_in, s_in)
    with torch.no_grad():
        seq, actions = rssm.imagine_rollout(init, actor, 1, noise=nz.cuda(), return_actions=True)
    _check("deter", seq.deter_states[:, 1], ref["deter"].reshape(B * H, D))
    _check("mean", seq.means[:, 1], ref["mean"].reshape(B * H, S))
    _check("std", seq.stds[:, 1], ref["std"].reshape(B * H, S))
    _check("stoch", seq.stoch_states[:, 1], ref["stoch"].reshape(B * H, S))
    _check("actions", actions[:, 0], ref["actions"].reshape(B * H, A))
    # index 0 carries the given initial state untouched
    assert torch.equal(seq.deter_states[:, 0].cpu(), d_in) and torch.equal(seq.stoch_states[:, 0].cpu(), s_in)


def test_bf16_free_running_bounded():
    D, S, Hd, A, Ha, La = SHAPES["c1"]
    HG, rssm, actor = _setup(D, S, Hd, A, Ha, La)
    B, H = 300, 15
    deter0, stoch0, noise = HG.synth_inputs("continuous", B, H, D, S, 1)
    init = HG.init_state(rssm, deter0, stoch0)
    with torch.no_grad():
        seq = rssm.imagine_rollout(init, actor, H, noise=noise.cuda())
        ref = O.imagine_rollout("continuous", HG.cpu_weights(rssm), HG.cpu_weights(actor), deter0, stoch0, noise, H, S)
    assert float((seq.deter_states[:, 1:].cpu() - ref["deter"]).abs().max()) < 3e-2
    assert float((seq.stoch_states[:, 1:].cpu() - ref["stoch"]).abs().max()) < 3e-2
    assert seq.deter_states.shape == (B, H + 1, D) and torch.isfinite(seq.stds).all()


def test_bf16_repack_on_weight_update():
    """the packed bf16 image follows parameter updates (version-counter keyed cache)."""
    D, S, Hd, A, Ha, La = SHAPES["small"]
    HG, rssm, actor = _setup(D, S, Hd, A, Ha, La)
    deter0, stoch0, noise = HG.synth_inputs("continuous", 64, 2, D, S, 1)
    init = HG.init_state(rssm, deter0, stoch0)
    with torch.no_grad():
        a = rssm.imagine_rollout(init, actor, 2, noise=noise.cuda()).deter_states.clone()
        b = rssm.imagine_rollout(init, actor, 2, noise=noise.cuda()).deter_states
        assert torch.equal(a, b), "same weights, same noise -> bitwise identical"
        rssm.rnn.weight_hh.mul_(0.5)
        c = rssm.imagine_rollout(init, actor, 2, noise=noise.cuda()).deter_states
        ref = O.imagine_rollout("continuous", HG.cpu_weights(rssm), HG.cpu_weights(actor), deter0, stoch0, noise, 2, S)
    assert not torch.equal(a, c)
    assert float((c[:, 1:].cpu() - ref["deter"]).abs().max()) < 2e-2


def test_bf16_with_grad_uses_recording_path():
    """when a backward will follow, the recording fp32 chain produces the graph."""
    D, S, Hd, A, Ha, La = SHAPES["small"]
    HG, rssm, actor = _setup(D, S, Hd, A, Ha, La)
    for p in rssm.parameters():
        p.requires_grad_(False)
    deter0, stoch0, noise = HG.synth_inputs("continuous", 32, 3, D, S, 1)
    seq = rssm.imagine_rollout(HG.init_state(rssm, deter0, stoch0), actor, 3, noise=noise.cuda())
    seq.get_features().square().mean().backward()
    assert actor.mu.weight.grad is not None and torch.isfinite(actor.mu.weight.grad).all()


@pytest.mark.parametrize("mode", ["1", "2", "3", "4"])
def test_bf16_kernel_variants(mode):
    """every tcgen05 kernel variant stays parity-green: 1 = single CTA, 2 = CTA pair (cta_group::2, M=256),
    3 = folded CTA pair (cta_group::2, M=128, 64 rows per CTA), 4 = folded pair with two-tile ping-pong.  WMRL_TC_MODE is read once per process, so
    each variant runs in a subprocess."""
    import os, subprocess, sys
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    env = dict(os.environ, WMRL_TC_MODE=mode)
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_gpu_parity_bf16.py"), "-q", "-x",
                        "-k", "teacher_forced or free_running", "-m", "gpu"], env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
