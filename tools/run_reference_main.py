#!/usr/bin/env python
"""Run the UNMODIFIED reference CLI (main.py training | inference | ...) with the learned models replaced by the B200
implementation -- the launcher of INTEGRATION.md as a shipped file.

    python tools/run_reference_main.py [--reference DIR] [--stock] training --model_type opnet \\
        --model_config configs/opnet_model_config.json --training_config configs/training_config.json

What it does, without touching a reference file:
  * puts the reference checkout on sys.path (default: the staged copy baseline/_ref, else /root/reference);
  * applies the two version shims the 2020 code needs on today's numpy / torch: ``np.int`` and the ``verbose`` keyword of
    ``ReduceLROnPlateau`` (oracle/stage_reference.import_reference holds the same shims for the tests);
  * rebinds ``get_model`` of the reference's ``ModelsFactory`` -- in every module that imported the class
    (baselines/training_main.py:11, inference_main.py:14) -- to objectpermanence_b200.models_factory.ModelsFactory.get_model,
    unless --stock is given (the reference's own modules, e.g. for a CPU run of the same command);
  * executes the reference's main.py under runpy with the remaining arguments.
Reference call stack: main.py:81-139 -> baselines/training_main.py:120-252 / baselines/inference_main.py:162-257.
"""
import inspect
import os
import runpy
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _shims():
    import numpy as np
    import torch
    if not hasattr(np, "int"):
        np.int = int
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau
    if "verbose" not in inspect.signature(sched.__init__).parameters and not getattr(sched, "_opn_verbose_shim", False):
        original = sched.__init__

        def init(self, *args, verbose=None, **kwargs):
            original(self, *args, **kwargs)

        sched.__init__ = init
        sched._opn_verbose_shim = True


def launch(argv, reference=None, stock=False):
    """Run the reference's main.py with `argv` (list of CLI arguments after the script name)."""
    if reference is None:
        staged = os.path.join(REPO, "baseline", "_ref")
        reference = staged if os.path.exists(os.path.join(staged, "main.py")) else "/root/reference"
    if not os.path.exists(os.path.join(reference, "main.py")):
        raise SystemExit(f"run_reference_main: no reference checkout at {reference} (main.py missing)")
    _shims()
    if REPO not in sys.path:
        sys.path.insert(0, REPO)
    if reference not in sys.path:
        sys.path.insert(0, reference)
    import baselines.models_factory as ref_factory
    if not hasattr(ref_factory.ModelsFactory, "_opn_stock_get_model"):
        ref_factory.ModelsFactory._opn_stock_get_model = ref_factory.ModelsFactory.__dict__["get_model"]
    if stock:    # (a previous launch in this process may have swapped it)
        ref_factory.ModelsFactory.get_model = ref_factory.ModelsFactory._opn_stock_get_model
    else:
        from objectpermanence_b200.models_factory import ModelsFactory as B200Factory
        ref_factory.ModelsFactory.get_model = staticmethod(B200Factory.get_model)   # trackers / detector factories untouched
    saved = sys.argv
    sys.argv = [os.path.join(reference, "main.py")] + list(argv)
    try:
        runpy.run_path(os.path.join(reference, "main.py"), run_name="__main__")
    finally:
        sys.argv = saved


def main():
    args = sys.argv[1:]
    reference, stock = None, False
    while args and args[0] in ("--reference", "--stock"):
        if args[0] == "--stock":
            stock, args = True, args[1:]
        else:
            reference, args = args[1], args[2:]
    launch(args, reference, stock)


if __name__ == "__main__":
    main()
