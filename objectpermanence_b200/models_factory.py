"""ModelsFactory.get_model with the reference's signature (baselines/models_factory.py:42-80).

Same names -> same classes, same AttributeError for unknown names.  Two deliberate differences,
both documented in INTEGRATION.md:
  * ``opnet_no_labels`` is accepted (the reference's factory only matches the misspelt
    ``opent_no_labels`` although its CLI offers ``opnet_no_labels``); the misspelling still works.
  * weights are loaded with ``map_location`` = the current CUDA device instead of the hard-coded
    ``"cuda:0"`` (baselines/models_factory.py:77) so one-process-per-GPU data parallelism works.
The tracker / detector factories of the reference are out of scope and not mirrored.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .learned_models import (AbstractCaterModel, BaselineLstm, NonLinearLstm, OPNet, OPNetLstmMlp,
                             TransformerLstm)

_FAMILY_CLASS = {
    "baseline_lstm": BaselineLstm,
    "non_linear_lstm": NonLinearLstm,
    "transformer_lstm": TransformerLstm,
    "opnet": OPNet,
    "opnet_lstm_mlp": OPNetLstmMlp,
}


def _family(model_name: str) -> Optional[str]:
    if model_name == "opent_no_labels":  # the reference's own spelling, models_factory.py:64
        return "opnet"
    base = model_name[:-len("_no_labels")] if model_name.endswith("_no_labels") else model_name
    return base if base in _FAMILY_CLASS else None


class ModelsFactory(object):

    @staticmethod
    def get_model(model_name: str, model_config: Dict[str, int], model_weights_path: str = None) -> AbstractCaterModel:
        family = _family(model_name)
        if family is None:
            raise AttributeError("Model name is incorrect")
        model = _FAMILY_CLASS[family](model_config)
        if model_weights_path is not None:
            location = f"cuda:{torch.cuda.current_device()}" if torch.cuda.is_available() else "cpu"
            model.load_state_dict(torch.load(model_weights_path, map_location=location))
            print(f"Loaded model parameters from {model_weights_path}")
        return model
