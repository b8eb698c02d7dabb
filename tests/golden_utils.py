"""Helpers shared by the CPU and GPU test modules: golden-fixture loading."""
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def manifest():
    with open(os.path.join(GOLDEN, "manifest.json")) as f:
        return json.load(f)


def model_cases():
    return [k for k in sorted(manifest()) if k != "iou_metric"]


def load_case(name):
    meta = manifest()[name]
    blob = np.load(os.path.join(GOLDEN, name + ".npz"))
    params = {k[len("param:"):]: torch.from_numpy(blob[k]) for k in blob.files if k.startswith("param:")}
    grads = {k[len("grad:"):]: torch.from_numpy(blob[k]) for k in blob.files if k.startswith("grad:")}
    out = {
        "meta": meta,
        "boxes": torch.from_numpy(blob["boxes"]),
        "labels": torch.from_numpy(blob["labels"]),
        "mask": torch.from_numpy(blob["mask"]),
        "y": torch.from_numpy(blob["y"]),
        "logits": torch.from_numpy(blob["logits"]) if "logits" in blob.files else None,
        "loss": float(blob["loss"]),
        "params": params,
        "grads": grads,
    }
    return out
