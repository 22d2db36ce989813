"""pytest plugin (``-p ref_dropin_plugin``) for running the REFERENCE'S OWN tests against the drop-in: stubs the
``gymnasium`` import of the reference's conftest (no environment is created by the tests selected), swaps the
reference's RSSM classes for ``wmrl``'s before the test modules import them, and makes CUDA the default device
so that the tensors and modules the tests build live where the native path runs (it has no CPU fallback)."""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "world-model-rl_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

if "gymnasium" not in sys.modules:
    gym = types.ModuleType("gymnasium")

    class _StubEnv:
        """what ``gym.make`` returns here: the reference evaluates it as a DEFAULT ARGUMENT at class-definition time
        (data/loader.py:74), so creating it must succeed; any use of it fails loudly."""

        def __getattr__(self, name):
            raise RuntimeError("gymnasium is stubbed: the selected reference tests do not step environments")

    def _no_env(*a, **k):
        return _StubEnv()

    gym.make = _no_env
    gym.Env = object
    sys.modules["gymnasium"] = gym

from oracle import make_ref  # noqa: E402

make_ref.import_reference()
import reflect.components.rssm_world_model.rssm as ref_rssm  # noqa: E402

import wmrl  # noqa: E402

ref_rssm.ContinuousRSSM = wmrl.ContinuousRSSM
ref_rssm.DiscreteRSSM = wmrl.DiscreteRSSM
torch.set_default_device("cuda")
