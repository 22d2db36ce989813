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
