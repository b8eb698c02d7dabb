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


def test_dropout_mask_restatement_matches_philox_known_answers():
    """Random123's published known-answer vectors for Philox4x32-10 pin the generator opn_dropout specifies."""
    from oracle.dropout_mask import keep_mask, philox4x32_10
    kats = [([0, 0, 0, 0], (0, 0), [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
            ([0xffffffff] * 4, (0xffffffff, 0xffffffff), [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
            ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], (0xa4093822, 0x299f31d0),
             [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, want in kats:
        got = philox4x32_10(np.array([ctr], dtype=np.uint32), key)[0]
        assert [int(v) for v in got] == want
    m = keep_mask(1_000_003, 0.1, 1234, 77)
    assert m.shape == (1_000_003,) and abs(m.mean() - 0.9) < 1e-3
    assert keep_mask(1000, 0.0, 1, 0).all()
    assert np.array_equal(keep_mask(64, 0.3, 5, 4), keep_mask(80, 0.3, 5, 0)[16:])   # offset = 4-element blocks


def test_unknown_model_name():
    with pytest.raises(AttributeError):
        oracle.family_of("opnet_v2")


def test_oracle_dropout_sites_match_the_torch_encoder_layer():
    """Train mode of the restated encoder layer: the three nn.Dropout modules of nn.TransformerEncoderLayer are
    replaced by fixed masks in a real torch layer and the same masks are pinned in the oracle (the attention-weight
    dropout lives inside scaled_dot_product_attention and cannot be pinned there: it is switched off on both sides and
    covered by the kernel-level test against plain fp64 math)."""
    torch.manual_seed(3)
    S, N, D, nhead, p_drop = 10, 3, 32, 2, 0.1
    layer = torch.nn.TransformerEncoderLayer(d_model=D, nhead=nhead).double().train()
    layer.self_attn.dropout = 0.0
    masks = {"dropout1": (torch.rand(S, N, D) >= p_drop).double() / (1 - p_drop),
             "dropout": (torch.rand(S, N, 2048) >= p_drop).double() / (1 - p_drop),
             "dropout2": (torch.rand(S, N, D) >= p_drop).double() / (1 - p_drop)}

    class Fixed(torch.nn.Module):
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, t):
            return t * self.m

    layer.dropout1, layer.dropout, layer.dropout2 = Fixed(masks["dropout1"]), Fixed(masks["dropout"]), Fixed(masks["dropout2"])
    x = torch.randn(S, N, D, dtype=torch.float64)
    want = layer(x)
    params = {f"L.{k}": v.detach() for k, v in layer.state_dict().items()}
    seen = []

    def drop(site, t):
        seen.append(site)
        name = site.rsplit(".", 1)[1]
        return t if name == "attn" else t * masks[name]

    got = oracle.encoder_layer(x, params, "L", nhead, drop)
    assert seen == ["L.attn", "L.dropout1", "L.dropout", "L.dropout2"]
    assert (got - want).abs().max().item() < 1e-12


def test_oracle_attention_in_query_blocks_is_the_same_function():
    """q_chunk (used for BASELINE config 3 at S = 9600 on the GPU box) must not change the restatement."""
    cfg = {"boxes_features_dim": 32, "num_attention_heads": 2, "num_attention_layers": 2, "num_lstm_layers": 1, "lstm_hidden_dim": 32}
    from objectpermanence_b200.synthetic import make_batch
    boxes = torch.from_numpy(make_batch(3, 7, 5, seed=3)[0]).double()
    params = {k: v.double() for k, v in oracle.init_params("transformer_lstm", cfg, seed=1).items()}
    full = oracle.transformer_lstm_forward(params, boxes, cfg)
    blocked = oracle.transformer_lstm_forward(params, boxes, cfg, q_chunk=5)
    assert (full - blocked).abs().max().item() < 1e-14
