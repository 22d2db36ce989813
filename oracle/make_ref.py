"""Stage the UNMODIFIED reference modules of the hot path under ``oracle/_ref/`` so that they travel to the
GPU box (``/root/reference`` does not exist there).

TEST / BASELINE INFRASTRUCTURE ONLY.  ``oracle/_ref/`` is git-ignored (nothing of the reference enters the
repository history) but not gpurun-ignored.  It is used by ``bench.py --impl reference`` (CPU arm, ``kind:
"reference"``), by ``bench.py``'s same-box eager-GPU baseline, and by ``tests/test_reference_drop_in.py``.
Run in the build container: ``python oracle/make_ref.py`` (``__graft_entry__.build()`` does it when
``/root/reference`` is present).  The files are byte-for-byte copies; only the package ``__init__`` chain
needed to import them is staged, nothing is edited.
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("WMRL_REFERENCE_SRC", "/root/reference/src")

# every module `reflect.components.rssm_world_model.{rssm,world_model,memory_actor}` and
# `reflect.components.trainers.value.*` import, plus the package __init__ files on the way
FILES = [
    "reflect/__init__.py",
    "reflect/utils.py",
    "reflect/components/__init__.py",
    "reflect/components/models/__init__.py",
    "reflect/components/models/actor.py",
    "reflect/components/models/encoder.py",
    "reflect/components/models/decoder.py",
    "reflect/components/rssm_world_model/__init__.py",
    "reflect/components/rssm_world_model/rssm.py",
    "reflect/components/rssm_world_model/models.py",
    "reflect/components/rssm_world_model/world_model.py",
    "reflect/components/rssm_world_model/memory_actor.py",
    "reflect/components/rssm_world_model/state/__init__.py",
    "reflect/components/rssm_world_model/state/base.py",
    "reflect/components/rssm_world_model/state/continuous.py",
    "reflect/components/rssm_world_model/state/discrete.py",
    "reflect/components/trainers/__init__.py",
    "reflect/components/trainers/value/__init__.py",
    "reflect/components/trainers/value/value_trainer.py",
    "reflect/components/trainers/value/critic.py",
    # the reference's own tests of this path that need no gymnasium environment, and what their conftest imports
    # (tests/test_reference_drop_in.py runs them against the drop-in modules on the GPU box)
    "reflect/components/rssm_world_model/tests/conftest.py",
    "reflect/components/rssm_world_model/tests/test_rssm.py",
    "reflect/components/rssm_world_model/tests/test_continuous_states.py",
    "reflect/components/rssm_world_model/tests/test_discrete_states.py",
    "reflect/components/trainers/reward/__init__.py",
    "reflect/components/trainers/reward/reward_trainer.py",
    "reflect/data/__init__.py",
    "reflect/data/loader.py",
    "reflect/data/noise.py",
]


def available() -> bool:
    """True when the staged copy can be imported (on the GPU box: staged by the build container)."""
    return os.path.exists(os.path.join(DEST, "reflect", "components", "rssm_world_model", "rssm.py"))


def stage(verbose: bool = False) -> bool:
    """Copy the files; returns False (and leaves DEST alone) when the reference tree is not present."""
    if not os.path.isdir(os.path.join(SRC, "reflect")):
        return False
    for rel in FILES:
        src = os.path.join(SRC, rel)
        if not os.path.exists(src):
            if rel.endswith("__init__.py"):
                os.makedirs(os.path.dirname(os.path.join(DEST, rel)), exist_ok=True)
                open(os.path.join(DEST, rel), "a").close()
                continue
            raise FileNotFoundError(src)
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    if verbose:
        print(f"staged {len(FILES)} reference files under {DEST}")
    return True


def import_reference():
    """Put the staged copy (or, in the build container, the reference tree itself) on sys.path and return
    the ``reflect`` package; raises ImportError when neither exists."""
    root = DEST if available() else (SRC if os.path.isdir(os.path.join(SRC, "reflect")) else None)
    if root is None:
        raise ImportError("reference modules are not staged: run `python oracle/make_ref.py` in the build container")
    if root not in sys.path:
        sys.path.insert(0, root)
    import reflect  # noqa: F401
    return reflect


if __name__ == "__main__":
    ok = stage(verbose=True)
    sys.exit(0 if ok else 1)
