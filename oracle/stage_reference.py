"""Stage the UNMODIFIED reference into the git-ignored ``baseline/_ref/`` so that it travels to the GPU box.

Test / measurement infrastructure only (like everything under ``oracle/``): ``bench.py --impl reference`` times the
reference's own modules from there, and the integration tests run its ``training_main`` / ``inference_main`` from there.
``/root/reference`` exists only in the build container; ``__graft_entry__.build()`` calls this when it is present.  Only
the Python sources and JSON configs the path needs are copied (no model files, no media); nothing staged is ever
committed (``baseline/_ref/`` is in .gitignore) and nothing under ``objectpermanence_b200/`` imports it.

    python oracle/stage_reference.py [/root/reference]
"""
from __future__ import annotations

import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(REPO, "baseline", "_ref")
KEEP_EXT = (".py", ".json")
TOP_LEVEL = ("baselines", "configs", "object_detection", "object_indices.py", "main.py")


def staged_root() -> str | None:
    """Path of the staged reference, or None when it has not been staged."""
    return STAGED if os.path.exists(os.path.join(STAGED, "baselines", "learned_models.py")) else None


def stage(reference: str = "/root/reference") -> str | None:
    if not os.path.isdir(os.path.join(reference, "baselines")):
        return staged_root()
    if os.path.isdir(STAGED):
        shutil.rmtree(STAGED)
    for top in TOP_LEVEL:
        src = os.path.join(reference, top)
        if os.path.isfile(src):
            os.makedirs(STAGED, exist_ok=True)
            shutil.copyfile(src, os.path.join(STAGED, top))
            continue
        for root, dirs, files in os.walk(src):
            dirs[:] = [d for d in dirs if d != "data"]   # DaSiamRPN/code/data: benchmark listings, not needed
            rel = os.path.relpath(root, reference)
            for name in files:
                if name.endswith(KEEP_EXT):
                    os.makedirs(os.path.join(STAGED, rel), exist_ok=True)
                    shutil.copyfile(os.path.join(root, name), os.path.join(STAGED, rel, name))
    return STAGED


def import_reference():
    """Put the staged reference on sys.path with the two version shims the 2020 code needs on today's numpy / torch
    (monkey-patches, the files stay untouched): ``np.int`` (baselines/datasets.py:475, tracking_utils.py:270) and the
    ``verbose`` keyword of ``ReduceLROnPlateau`` (baselines/training_main.py:151).  Returns the staged root."""
    root = staged_root()
    if root is None:
        raise RuntimeError("the reference has not been staged (baseline/_ref is empty): run oracle/stage_reference.py in "
                           "the build container, where /root/reference is mounted")
    import numpy as np
    import torch
    if not hasattr(np, "int"):
        np.int = int
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau
    if not getattr(sched, "_opn_verbose_shim", False):
        import inspect
        if "verbose" not in inspect.signature(sched.__init__).parameters:
            original = sched.__init__

            def init(self, *args, verbose=None, **kwargs):
                original(self, *args, **kwargs)

            sched.__init__ = init
        sched._opn_verbose_shim = True
    if root not in sys.path:
        sys.path.insert(0, root)
    return root


if __name__ == "__main__":
    print(stage(sys.argv[1] if len(sys.argv) > 1 else "/root/reference"))
