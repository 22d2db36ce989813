"""The reference's own tests of the path (rssm_world_model/tests/test_rssm.py, test_continuous_states.py,
test_discrete_states.py - staged unmodified under oracle/_ref by oracle/make_ref.py) run against the drop-in:
``reflect...rssm.ContinuousRSSM / DiscreteRSSM`` are replaced by ``wmrl``'s classes before the test modules import
them (tests/ref_dropin_plugin.py), everything else - Actor, ConvEncoder / ConvDecoder, the state dataclasses the
tests construct and ``isinstance``-check - is the reference's.  The env-driven tests (EnvDataLoader over a MuJoCo
gymnasium env) are out of reach on the box and are not selected."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_reference_rssm_tests_pass_on_the_drop_in():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle import make_ref
    if not make_ref.available():
        pytest.skip("reference modules are not staged (oracle/make_ref.py runs in the build container)")
    here = os.path.dirname(os.path.abspath(__file__))
    tdir = os.path.join(make_ref.DEST, "reflect", "components", "rssm_world_model", "tests")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([here, make_ref.DEST, os.environ.get("PYTHONPATH", "")]))
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "ref_dropin_plugin", "-p", "no:cacheprovider",
                        os.path.join(tdir, "test_rssm.py"), os.path.join(tdir, "test_continuous_states.py"),
                        os.path.join(tdir, "test_discrete_states.py")], env=env, capture_output=True, text=True,
                       timeout=900, cwd=tdir)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-1000:]


def test_sequences_come_back_as_reference_dataclasses():
    """type in, type out: a reference state in -> the reference's Sequence dataclass out."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle import make_ref
    if not make_ref.available():
        pytest.skip("reference modules are not staged")
    make_ref.import_reference()
    from reflect.components.rssm_world_model import state as RS
    import helpers_gpu as HG
    rssm, actor = HG.build_modules("continuous", 32, 8, 1, 32, 16, 2, 64, 2)
    d0, s0, nz = HG.synth_inputs("continuous", 40, 3, 32, 8, 1)
    init = RS.InternalStateContinuous(deter_state=d0.cuda(), stoch_state=s0.cuda(), mean=s0.cuda(), std=torch.ones_like(s0).cuda())
    with torch.no_grad():
        seq = rssm.imagine_rollout(init, actor, 3, noise=nz.cuda())
    assert isinstance(seq, RS.InternalStateContinuousSequence)
    assert seq.get_features().shape == (40, 3, 40) and seq.flatten_batch_time().deter_state.shape == (120, 32)
    pri, post = rssm.observe_rollout(torch.randn(5, 4, 16).cuda(), torch.randn(5, 4, 2).cuda())
    assert isinstance(post, RS.InternalStateContinuousSequence) and post.get_dist() is not None
