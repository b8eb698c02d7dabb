"""Model-name registries with the reference's names (baselines/supported_models.py:2-64).

Built from the model families rather than spelled out; the resulting lists are element-for-
element those of the reference, including the programmed (non-learned) trackers, which this
package does not implement (they stay in the reference, see DESIGN.md "out of scope")."""
from typing import List

_FIVE_TRACK_FAMILIES = ["baseline_lstm", "non_linear_lstm", "transformer_lstm"]
_SIX_TRACK_FAMILIES = ["opnet", "opnet_lstm_mlp"]


def _with_variants(families: List[str]) -> List[str]:
    return [name for fam in families for name in (fam, fam + "_no_labels")]


PROGRAMMED_MODELS: List[str] = ["detector_tracker", "detector_heuristic"]
TRAINING_SUPPORTED_MODELS_5_TRACKS: List[str] = _with_variants(_FIVE_TRACK_FAMILIES)
TRAINING_SUPPORTED_MODELS_6_TRACKS: List[str] = _with_variants(_SIX_TRACK_FAMILIES)
TRAINING_SUPPORTED_MODELS: List[str] = TRAINING_SUPPORTED_MODELS_5_TRACKS + TRAINING_SUPPORTED_MODELS_6_TRACKS
INFERENCE_SUPPORTED_MODELS: List[str] = PROGRAMMED_MODELS + TRAINING_SUPPORTED_MODELS
DOUBLE_OUTPUT_MODELS: List[str] = list(TRAINING_SUPPORTED_MODELS_6_TRACKS)
NO_LABELS_MODELS: List[str] = [n for n in TRAINING_SUPPORTED_MODELS if n.endswith("_no_labels")]
