"""CPU: the host-side mirror of the reference interface and the C-ABI library surface.
No compute calls (there is no GPU here and no CPU path in the product)."""
import ctypes
import json
import os
import re

import pytest
import torch

from objectpermanence_b200 import _lib, supported_models
from objectpermanence_b200.models_factory import ModelsFactory
from oracle import opnet_oracle as oracle

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SHIPPED_CONFIGS = {  # the reference's configs/*.json, restated
    "opnet": {"object_to_track_pred_dim": 15, "object_to_track_hidden_dim": 256, "videos_hidden_dim": 512},
    "opnet_lstm_mlp": {"object_to_track_pred_dim": 15, "object_to_track_hidden_dim": 256, "videos_hidden_dim": 512},
    "baseline_lstm": {"videos_hidden_dim": 512},
    "non_linear_lstm": {"boxes_features_dim": 256, "videos_hidden_dim": 512},
    "transformer_lstm": {"boxes_features_dim": 256, "num_attention_heads": 2, "num_attention_layers": 2,
                         "num_lstm_layers": 2, "lstm_hidden_dim": 512},
}
PARAM_COUNTS = {"opnet": 1421056, "opnet_lstm_mlp": 363264, "baseline_lstm": 1204224, "non_linear_lstm": 11013376,
                "transformer_lstm": 6303488}


@pytest.mark.parametrize("name", sorted(SHIPPED_CONFIGS))
def test_state_dict_contract(name):
    model = ModelsFactory.get_model(name, SHIPPED_CONFIGS[name])
    sd = model.state_dict()
    expected = oracle.param_shapes(name, SHIPPED_CONFIGS[name])
    assert {k: tuple(v.shape) for k, v in sd.items()} == expected
    assert all(v.dtype == torch.float32 for v in sd.values())
    assert sum(v.numel() for v in sd.values()) == PARAM_COUNTS[name]
    assert model.max_objects_in_frame == 15 and model.bb_out_dim == 4


def test_factory_names():
    for name in supported_models.TRAINING_SUPPORTED_MODELS:
        fam = name[:-len("_no_labels")] if name.endswith("_no_labels") else name
        assert type(ModelsFactory.get_model(name, SHIPPED_CONFIGS[fam])).__name__
    assert type(ModelsFactory.get_model("opent_no_labels", SHIPPED_CONFIGS["opnet"])).__name__ == "OPNet"
    with pytest.raises(AttributeError, match="Model name is incorrect"):
        ModelsFactory.get_model("detector_tracker", {})


def test_registries():
    assert len(supported_models.INFERENCE_SUPPORTED_MODELS) == 12
    assert supported_models.INFERENCE_SUPPORTED_MODELS[:2] == supported_models.PROGRAMMED_MODELS
    assert supported_models.DOUBLE_OUTPUT_MODELS == ["opnet", "opnet_no_labels", "opnet_lstm_mlp",
                                                     "opnet_lstm_mlp_no_labels"]
    assert len(supported_models.NO_LABELS_MODELS) == 5
    assert set(supported_models.TRAINING_SUPPORTED_MODELS_5_TRACKS) | set(
        supported_models.TRAINING_SUPPORTED_MODELS_6_TRACKS) == set(supported_models.TRAINING_SUPPORTED_MODELS)


def test_state_dict_round_trip(tmp_path):
    a = ModelsFactory.get_model("baseline_lstm", {"videos_hidden_dim": 32})
    path = tmp_path / "w.pth"
    torch.save(a.state_dict(), path)
    b = ModelsFactory.get_model("baseline_lstm", {"videos_hidden_dim": 32}, str(path))
    for k, v in a.state_dict().items():
        assert torch.equal(v, b.state_dict()[k])


def test_cpu_tensors_are_rejected_loudly():
    model = ModelsFactory.get_model("baseline_lstm", {"videos_hidden_dim": 32})
    with pytest.raises(RuntimeError, match="no CPU path"):
        model(torch.zeros(2, 8, 15, 5))


def test_bad_input_shape():
    model = ModelsFactory.get_model("opnet", SHIPPED_CONFIGS["opnet"])
    with pytest.raises(RuntimeError, match="expects boxes"):
        model(torch.zeros(2, 8, 15, 5))


def _header_symbols():
    text = open(os.path.join(REPO, "include", "opnet_b200.h")).read()
    return re.findall(r"^OPN_API\s+[\w\s\*]+?\b(opn_\w+)\s*\(", text, flags=re.M)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        from objectpermanence_b200.build import build
        build()
    names = _header_symbols()
    assert len(names) >= 18
    assert set(names) == set(_lib.SIGNATURES)  # the ctypes table mirrors the header one to one
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), n
    assert _lib.load().opn_version() >= 100


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.OpnError, match="no fallback"):
        _lib.load()


def test_product_never_imports_oracle():
    pkg = os.path.join(REPO, "objectpermanence_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_fused_adam_and_iou_eval_have_no_cpu_path():
    import torch
    from objectpermanence_b200.evaluation import iou_eval
    from objectpermanence_b200.optim import FusedAdam
    with pytest.raises(RuntimeError, match="CUDA"):
        FusedAdam([torch.nn.Parameter(torch.zeros(4))], lr=1e-3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        iou_eval(torch.zeros(1, 2, 4), torch.zeros(1, 2, 4))


def test_dropout_sites_follow_torch_manual_seed_and_have_no_cpu_path():
    """Host side of the encoder's train-mode dropout: every site draws its Philox key from torch's CPU generator (so a
    run is reproducible under torch.manual_seed and no two sites share a key); eval mode / p = 0 are the identity and
    never touch the library; train mode on a CPU tensor raises (no CPU path)."""
    from objectpermanence_b200 import ops
    torch.manual_seed(123)
    a, b = ops.dropout_stream.take(1000), ops.dropout_stream.take(1000)
    torch.manual_seed(123)
    c = ops.dropout_stream.take(7)
    assert a == c and a != b and a[1] == 0 and 0 <= a[0] < 2 ** 62
    x = torch.ones(4, 8)
    assert ops.dropout(x, 0.1, training=False) is x and ops.dropout(x, 0.0, training=True) is x
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.dropout(x, 0.1, training=True)
    # the transformer variant keeps the reference's default p = 0.1 and switches with train() / eval()
    cfg = {"boxes_features_dim": 32, "num_attention_heads": 2, "num_attention_layers": 2, "num_lstm_layers": 1,
           "lstm_hidden_dim": 32}
    model = ModelsFactory.get_model("transformer_lstm", cfg)
    assert all(layer.dropout_p == 0.1 for layer in model.attention_encoder.layers)
    assert model.training and not model.eval().training


def test_workspace_planners_are_pure_host_code():
    """opn_wgrad_workspace_bytes / opn_attention_workspace_bytes plan on the host (no GPU needed): the ctypes structure of
    a weight-gradient job matches the C struct (a mismatch would scramble the sizes), the plan is deterministic, and bad
    jobs are refused with a message instead of a size."""
    lib = _lib.load()
    rows, T = 9600, 300
    fake = 0x7f0000000000      # never dereferenced by the planner
    spec = [(2048, 512, 1), (2048, 6, 0), (1024, 256, 1), (1024, 90, 0), (256, 15, 0)]
    jobs = (_lib.WgradJob * len(spec))(*[_lib.WgradJob(fake, fake + 256, fake + 512, M, N, N, rows, T, M, N, sh, 0) for M, N, sh in spec])
    n = lib.opn_wgrad_workspace_bytes(len(spec), jobs)
    assert n == lib.opn_wgrad_workspace_bytes(len(spec), jobs)
    # at least the split planes of the small operands (2 planes of bf16, padded to 64 columns) and one partial sum per product
    floor = sum(2 * rows * ((N + 63) // 64 * 64) * 2 + M * ((N + 63) // 64 * 64) * 4 for M, N, _ in spec)
    assert floor <= n <= 8 * floor
    bad = (_lib.WgradJob * 1)(_lib.WgradJob(fake, fake, fake, 100, 64, 64, rows, T, 100, 64, 0, 0))      # M not a multiple of 128
    assert lib.opn_wgrad_workspace_bytes(1, bad) == 0
    assert b"multiple of 128" in lib.opn_last_error()
    # attention: planes of Q', K, V, dO (hi / lo bf16), row statistics, split-row partials, dropout keep bits
    S, D, heads = 9600, 256, 2
    a = lib.opn_attention_workspace_bytes(S, D, heads)
    assert a >= 4 * 2 * S * D * 2 + heads * (S // 64) * S * 8
    assert lib.opn_attention_workspace_bytes(S, 200, heads) == 0      # head dimension other than 128: the fused kernels refuse
