"""The prediction pass and the result writer of the reference's ``reasoning_inference_main``
(baselines/inference_main.py:162-257), without its debug-video rendering.

The reference runs the model, copies every float prediction to the host, scales to pixels with numpy, and only then
-- inside the cv2 video loop -- writes one ``<video>_bb.json`` per video.  Here the forward runs without the backward
stash (the persistent kernels skip gates / cells under ``torch.no_grad()``), the pixel conversion
``(x * [320,240,320,240]).astype(int32)`` happens on the device (``opn_to_pixels``), one copy brings the int32 boxes
back, and the JSON files are written directly (same name, same content as ``DataHelper.write_bb_predictions_to_file``,
baselines/tracking_utils.py:96-103), so inference throughput is not tied to ``VideoHandling``.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Dict, Iterable, Tuple

import numpy as np
import torch

from . import _lib, ops
from .supported_models import DOUBLE_OUTPUT_MODELS


def to_pixels(boxes: torch.Tensor) -> torch.Tensor:
    """float32 [..., 4] normalised xyxy (CUDA) -> int32 [..., 4] pixel boxes with the reference's semantics."""
    ops._require_cuda(boxes)
    boxes = boxes.contiguous()
    if boxes.shape[-1] != 4:
        raise RuntimeError("to_pixels expects [..., 4] boxes")
    out = torch.empty(boxes.shape, dtype=torch.int32, device=boxes.device)
    rc = _lib.load().opn_to_pixels(boxes.numel() // 4, boxes.data_ptr(), out.data_ptr(), ops._stream())
    _lib.check(rc, "opn_to_pixels")
    return out


def predict_pixel_boxes(model_name: str, model: torch.nn.Module, device: torch.device,
                        data_loader: Iterable) -> Tuple[Dict[str, int], np.ndarray, np.ndarray]:
    """inference_main.py:186-215: (video name -> row, predictions int32 [N,T,4], labels int32 [N,T,4])."""
    double_out = model_name in DOUBLE_OUTPUT_MODELS or model_name == "opent_no_labels"
    model.eval()
    model.to(device)
    indices: Dict[str, int] = {}
    preds, labs = [], []
    n = 0
    with torch.no_grad():
        for sample in data_loader:
            x, y, video_names = sample
            boxes, _ = x
            labels, _ = y
            out = model(boxes.to(device, non_blocking=True))
            output = out[0] if double_out else out
            preds.append(to_pixels(output))
            labs.append(to_pixels(labels.to(device, non_blocking=True).float()))
            for i, name in enumerate(video_names):
                indices[name] = n + i
            n += len(video_names)
    if n == 0:
        return indices, np.zeros((0, 0, 4), np.int32), np.zeros((0, 0, 4), np.int32)
    out_p, out_l = torch.cat(preds).cpu().numpy(), torch.cat(labs).cpu().numpy()
    ops.check_status(device, "inference pass")   # the copies above synchronised: a timed-out recurrence raises here
    return indices, out_p, out_l


def write_bb_predictions(video_path: str, predictions_dir: str, boxes) -> Path:
    """``DataHelper.write_bb_predictions_to_file`` (tracking_utils.py:96-103): ``<stem>_bb.json``, a list of
    ``[x1, y1, x2, y2]`` ints per frame, ``indent=2``."""
    path = Path(predictions_dir) / (Path(video_path).stem + "_bb.json")
    rows = [[int(x1), int(y1), int(x2), int(y2)] for [x1, y1, x2, y2] in boxes]
    with open(path, "w") as f:
        json.dump(rows, f, indent=2)
    return path


def run_inference(model_name: str, model: torch.nn.Module, device: torch.device, data_loader: Iterable,
                  results_dir: str) -> Dict[str, Path]:
    """Prediction pass + one ``<video>_bb.json`` per video (what the reference leaves behind for
    ``analyze_iou_offline.py``)."""
    indices, preds, _ = predict_pixel_boxes(model_name, model, device, data_loader)
    Path(results_dir).mkdir(parents=True, exist_ok=True)
    return {name: write_bb_predictions(name, results_dir, preds[row]) for name, row in indices.items()}
