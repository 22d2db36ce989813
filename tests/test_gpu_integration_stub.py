"""The reference-side binding shown in INTEGRATION.md is executed as written (ctypes only, no wmrl Python layer) and must
reproduce the module path - the document stays honest."""
import ctypes as C
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_heads_stub_from_integration_md():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import wmrl
    from wmrl import _lib
    lib = C.CDLL(_lib.LIB_PATH)
    lib.wmrl_last_error.restype = C.c_char_p
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    code = md.split("### The heads and the acting step, same pattern")[1].split("```python")[1].split("```")[0]
    ns = {"C": C, "torch": torch, "lib": lib}
    exec(code, ns)
    m = wmrl.ValueCritic(230, precision="bf16").cuda()
    r = wmrl.DenseModel(input_dim=230, precision="bf16").cuda()
    x = torch.randn(37, 5, 230, device="cuda")
    with torch.no_grad():
        assert torch.equal(ns["mlp_fwd"](m.layers, x, 2), m(x))
        assert torch.equal(ns["mlp_fwd"](r.fc, x, 1), r(x))


def test_imagine_stub_from_integration_md():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import helpers_gpu as HG
    from wmrl import _lib
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    code = md.split("## Option B")[1].split("```python")[1].split("```")[0]
    code = code.replace('C.CDLL("libwmrl.so")', f'C.CDLL("{_lib.LIB_PATH}")')
    ns = {}
    exec(code, ns)
    D, S, A, Ha, La, B, H = 200, 30, 6, 512, 3, 70, 4
    rssm, actor = HG.build_modules("continuous", D, S, 1, 200, 64, A, Ha, La, seed=2)
    d0, s0, nz = HG.synth_inputs("continuous", B, H, D, S, 1)
    with torch.no_grad():
        deter, stoch, mean, std, act = ns["imagine_fwd"](rssm, actor, d0.cuda(), s0.cuda(), nz.cuda(), H)
        seq = rssm.imagine_rollout(HG.init_state(rssm, d0, s0), actor, H, noise=nz.cuda())
    torch.cuda.synchronize()
    assert torch.equal(deter, seq.deter_states) and torch.equal(stoch, seq.stoch_states)
    assert torch.equal(mean[:, 1:], seq.means[:, 1:]) and torch.equal(std[:, 1:], seq.stds[:, 1:])
