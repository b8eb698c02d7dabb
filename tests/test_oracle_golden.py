"""CPU: the oracle restatement (oracle/opnet_oracle.py) against the golden vectors that
oracle/make_golden.py recorded from the unmodified reference modules."""
import numpy as np
import pytest
import torch

from golden_utils import GOLDEN, load_case, model_cases
from oracle import opnet_oracle as oracle


@pytest.mark.parametrize("name", model_cases())
@pytest.mark.parametrize("fast", [False, True])
def test_oracle_matches_reference_fixture(name, fast):
    case = load_case(name)
    meta = case["meta"]
    y, logits, loss, grads = oracle.loss_and_grads(meta["model_name"], case["params"], case["boxes"], case["labels"],
                                                   meta["config"], dtype=torch.float32, fast=fast, mask=case["mask"])
    assert (y - case["y"]).abs().max().item() <= 5e-6
    if case["logits"] is not None:
        assert logits.shape == case["logits"].shape
        assert (logits - case["logits"]).abs().max().item() <= 5e-6
    assert abs(loss.item() - case["loss"]) <= 1e-6
    assert set(grads) == set(case["grads"])
    for k in grads:
        assert (grads[k] - case["grads"][k]).abs().max().item() <= 5e-6, k


@pytest.mark.parametrize("name", model_cases())
def test_oracle_fp64_agrees_with_fp32_reference(name):
    case = load_case(name)
    meta = case["meta"]
    y, _, _, grads = oracle.loss_and_grads(meta["model_name"], case["params"], case["boxes"], case["labels"],
                                           meta["config"], dtype=torch.float64, mask=case["mask"])
    assert (y.float() - case["y"]).abs().max().item() <= 5e-6
    for k in grads:
        assert (grads[k].float() - case["grads"][k]).abs().max().item() <= 5e-6, k


def test_param_shapes_match_fixture():
    for name in model_cases():
        case = load_case(name)
        shapes = oracle.param_shapes(case["meta"]["model_name"], case["meta"]["config"])
        assert {k: tuple(v.shape) for k, v in case["params"].items()} == shapes


def test_transformer_slot0_equals_all_slots():
    case = load_case("transformer_lstm_d32_h32")
    p = {k: v.double() for k, v in case["params"].items()}
    a = oracle.transformer_lstm_forward(p, case["boxes"].double(), case["meta"]["config"], all_slots=True)
    b = oracle.transformer_lstm_forward(p, case["boxes"].double(), case["meta"]["config"], all_slots=False)
    assert (a - b).abs().max().item() < 1e-12


def test_iou_metric_matches_reference_analyzer():
    blob = np.load(f"{GOLDEN}/iou_metric.npz")
    assert abs(oracle.mean_iou(blob["pred"], blob["gt"]) - float(blob["mean_iou"])) < 1e-12


def test_unknown_model_name():
    with pytest.raises(AttributeError):
        oracle.family_of("opnet_v2")
