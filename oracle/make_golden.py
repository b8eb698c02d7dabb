"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules.

Runs only in the build container, where the reference is mounted read-only at
/root/reference.  The reference is imported (never copied); what is committed are the
seeded inputs, the reference's own randomly initialised weights, its outputs and its
parameter gradients for the loss of training_main.py:192-210.  The GPU box has no
/root/reference: tests read only the .npz files.

    python oracle/make_golden.py            # rewrites tests/golden/

Also cross-checks ``oracle/opnet_oracle.py`` against the reference while it is at it
and prints the max-abs deviations (they are asserted in tests/test_oracle_golden.py).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("OPN_REFERENCE", "/root/reference")
sys.path.insert(0, REPO)
sys.path.insert(0, REFERENCE)

from objectpermanence_b200.synthetic import make_batch  # noqa: E402
from oracle import opnet_oracle as oracle  # noqa: E402

np.int = int  # numpy>=1.24 shim for baselines/tracking_utils.py:270 (monkey-patch, reference untouched)
import baselines.learned_models as ref_models  # noqa: E402
from baselines.tracking_utils import ResultsAnalyzer  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")

# (fixture name, reference class, model name, config, B, T, weight scale)
CASES = [
    ("baseline_lstm_h32", ref_models.BaselineLstm, "baseline_lstm", {"videos_hidden_dim": 32}, 2, 8, 1.0),
    ("opnet_h32_h64", ref_models.OPNet, "opnet",
     {"object_to_track_pred_dim": 15, "object_to_track_hidden_dim": 32, "videos_hidden_dim": 64}, 3, 12, 1.0),
    ("opnet_h32_h64_x6", ref_models.OPNet, "opnet",
     {"object_to_track_pred_dim": 15, "object_to_track_hidden_dim": 32, "videos_hidden_dim": 64}, 3, 12, 6.0),
    ("opnet_no_labels_h32", ref_models.OPNet, "opent_no_labels",
     {"object_to_track_pred_dim": 15, "object_to_track_hidden_dim": 32, "videos_hidden_dim": 32}, 2, 10, 1.0),
    ("opnet_lstm_mlp_h32", ref_models.OPNetLstmMlp, "opnet_lstm_mlp",
     {"object_to_track_pred_dim": 15, "object_to_track_hidden_dim": 32, "videos_hidden_dim": 64}, 2, 9, 1.0),
    ("non_linear_lstm_d32_h32", ref_models.NonLinearLstm, "non_linear_lstm",
     {"boxes_features_dim": 32, "videos_hidden_dim": 32}, 2, 7, 1.0),
    ("transformer_lstm_d32_h32", ref_models.TransformerLstm, "transformer_lstm",
     {"boxes_features_dim": 32, "num_attention_heads": 2, "num_attention_layers": 2, "num_lstm_layers": 2,
      "lstm_hidden_dim": 32}, 2, 6, 1.0),
]


def run_reference(cls, model_name, config, B, T, scale, seed):
    torch.manual_seed(seed)
    model = cls(dict(config)).eval()  # eval(): dropout off (transformer); no-op for the LSTM models
    if scale != 1.0:
        with torch.no_grad():
            for prm in model.parameters():
                prm.mul_(scale)
    F = oracle.in_features_of(model_name)
    boxes_np, labels_np, mask_np = make_batch(B, T, F, seed=1234 + seed)
    boxes = torch.from_numpy(boxes_np)
    labels = torch.from_numpy(labels_np)
    mask = torch.from_numpy(mask_np)
    out = model(boxes)
    y, logits = out if isinstance(out, tuple) else (out, None)
    # the loss of training_main.py:192-210, evaluated with the reference's own expressions
    pred_loss = torch.nn.L1Loss(reduction="none")(y, labels)
    cons = torch.mean(torch.norm(y[:, 1:, :] - y[:, :-1, :], p=2, dim=-1))
    if model_name.endswith("no_labels"):
        loss = torch.mean(pred_loss * mask) + 0.5 * cons
    else:
        loss = torch.mean(pred_loss)
    loss.backward()
    params = {k: v.detach().clone() for k, v in model.state_dict().items()}
    grads = {k: (v.grad.detach().clone() if v.grad is not None else torch.zeros_like(v))
             for k, v in model.named_parameters()}
    return boxes, labels, mask, y.detach(), None if logits is None else logits.detach(), loss.detach(), params, grads


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)  # deterministic summation order for the fixtures
    manifest = {}
    for idx, (name, cls, model_name, config, B, T, scale) in enumerate(CASES):
        boxes, labels, mask, y, logits, loss, params, grads = run_reference(cls, model_name, config, B, T, scale, seed=idx)
        blob = {"boxes": boxes.numpy(), "labels": labels.numpy(), "mask": mask.numpy(), "y": y.numpy(),
                "loss": loss.numpy()}
        if logits is not None:
            blob["logits"] = logits.numpy()
        for k, v in params.items():
            blob["param:" + k] = v.numpy()
        for k, v in grads.items():
            blob["grad:" + k] = v.numpy()
        np.savez(os.path.join(OUT, name + ".npz"), **blob)
        manifest[name] = {"model_name": model_name, "config": config, "B": B, "T": T, "weight_scale": scale,
                          "reference_class": cls.__name__}

        # cross-check the oracle restatement (fp32 explicit, fp32 fast, fp64 explicit)
        for fast in (False, True):
            oy, ologits, oloss, ograds = oracle.loss_and_grads(model_name, params, boxes, labels, config,
                                                               dtype=torch.float32, fast=fast, mask=mask)
            dy = (oy - y).abs().max().item()
            dg = max((ograds[k] - grads[k]).abs().max().item() for k in grads)
            print(f"{name:28s} fast={fast!s:5s} max|dy|={dy:.2e} max|dgrad|={dg:.2e} loss diff={abs(oloss.item() - loss.item()):.2e}")
        if model_name.startswith("transformer"):
            p64 = {k: v.double() for k, v in params.items()}
            a = oracle.transformer_lstm_forward(p64, boxes.double(), config, all_slots=True)
            b = oracle.transformer_lstm_forward(p64, boxes.double(), config, all_slots=False)
            print(f"{name:28s} slot-0-only vs all-slots max|dy|={(a - b).abs().max().item():.2e}")

    # IoU metric fixture: the reference's own ResultsAnalyzer on integer boxes
    rng = np.random.default_rng(7)
    n, T = 6, 300
    gt = np.zeros((n, T, 4), dtype=np.float32)
    for i in range(n):
        _, lab, _ = make_batch(1, T, 5, seed=500 + i)
        gt[i] = lab[0]
    pred = (gt + rng.normal(0, 0.02, size=gt.shape)).astype(np.float32)
    frame = np.array([320, 240, 320, 240])
    pred_px = (pred.reshape(-1, 4) * frame).reshape(n, T, 4).astype(np.int32)
    gt_px = (gt.reshape(-1, 4) * frame).reshape(n, T, 4).astype(np.int32)
    analyzer = ResultsAnalyzer([str(i) for i in range(n)], pred_px, gt_px)
    analyzer.compute_aggregated_metric("video_mean", np.mean)
    df = analyzer.get_analysis_df()
    miou = float(np.mean(df["video_mean_iou"]))
    np.savez(os.path.join(OUT, "iou_metric.npz"), pred=pred, gt=gt, mean_iou=np.float64(miou),
             per_video=np.asarray(df["video_mean_iou"], dtype=np.float64))
    print("iou fixture mean_iou", miou, "oracle", oracle.mean_iou(pred, gt))
    manifest["iou_metric"] = {"n": n, "T": T}
    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
