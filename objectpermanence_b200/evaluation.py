"""The per-epoch evaluation pass of the reference (``inference_and_iou_comp``, baselines/training_main.py:32-117)
with the model, the loss and the IoU metric on the device.

The reference copies every prediction to the host (``output.cpu().numpy()``, :85), builds int32 pixel boxes with numpy
and runs ``ResultsAnalyzer`` (pandas) over them.  Here the same integer semantics run in one kernel per batch
(``opn_iou_eval``: double product, truncation to int32, +1-pixel IoU, per-video means) and only 20 bytes per video plus
one loss scalar leave the device, once, at the end of the pass.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

from . import _lib, ops
from .supported_models import DOUBLE_OUTPUT_MODELS, NO_LABELS_MODELS


def iou_eval(y: torch.Tensor, labels: torch.Tensor, mask: Optional[torch.Tensor] = None, per_frame: bool = False):
    """y, labels [N,T,4] normalised xyxy (float32, CUDA); mask [N,T,4] bool/uint8 or None.
    Returns (video_mean_iou [N] f64, containment_mean_iou [N] f64 | None, containment_frames [N] i32 | None,
    frame_iou [N,T] f64 | None) as device tensors; nothing synchronises."""
    ops._require_cuda(y, labels)
    y = y.contiguous()
    labels = labels.contiguous()
    N, T, C = y.shape
    if C != 4 or labels.shape != y.shape:
        raise RuntimeError("iou_eval expects predictions and labels of shape [N,T,4]")
    dev = y.device
    video = torch.empty(N, dtype=torch.float64, device=dev)
    m = masked = frames = fiou = None
    if mask is not None:
        m = mask.to(device=dev, dtype=torch.uint8).contiguous()
        masked = torch.empty(N, dtype=torch.float64, device=dev)
        frames = torch.empty(N, dtype=torch.int32, device=dev)
    if per_frame:
        fiou = torch.empty(N, T, dtype=torch.float64, device=dev)
    rc = _lib.load().opn_iou_eval(N, T, y.data_ptr(), labels.data_ptr(), ops._ptr(m), video.data_ptr(), ops._ptr(masked),
                                  ops._ptr(frames), ops._ptr(fiou), ops._stream())
    _lib.check(rc, "opn_iou_eval")
    return video, masked, frames, fiou


def _nanmean(values) -> float:
    vals = [v for v in values if not math.isnan(v)]   # pandas' Series.mean() skips NaN (training_main.py:109-110)
    return sum(vals) / len(vals) if vals else float("nan")


def inference_and_iou_comp(model_name: str, model: torch.nn.Module, compute_device: torch.device, data_loader,
                           dataset_length: Optional[int] = None, reg_loss_function=None) -> Tuple[float, float, float]:
    """Same signature and return value as the reference function: (average loss, mean IoU, containment mean IoU).
    ``reg_loss_function`` is accepted for signature compatibility; the loss is the reference's L1 form
    (training_main.py:60-79) computed by ``opn_loss_fwd_bwd``."""
    no_labels = model_name in NO_LABELS_MODELS or model_name == "opent_no_labels"
    double_out = model_name in DOUBLE_OUTPUT_MODELS or model_name == "opent_no_labels"
    model.eval()
    model.to(compute_device)
    loss_sum = torch.zeros((), dtype=torch.float64, device=compute_device)
    n_seen = 0
    video_means, masked_means = [], []
    with torch.no_grad():
        for sample in data_loader:
            x, y, _ = sample
            boxes, _ = x
            labels, mask = y
            boxes = boxes.to(compute_device, non_blocking=True)
            labels = labels.to(compute_device, non_blocking=True)
            mask = mask.to(compute_device, non_blocking=True) if mask is not None and mask.numel() > 0 else None
            out = model(boxes)
            output = out[0] if double_out else out
            loss3 = ops.training_loss(output, labels, mask, no_labels)
            bs = labels.shape[0]
            loss_sum += loss3[0].double() * bs
            n_seen += bs
            vm, mm, _, _ = iou_eval(output, labels, mask)
            video_means.append(vm)
            if mm is not None:
                masked_means.append(mm)
    if n_seen == 0:
        return 0.0, 0.0, 0.0
    video = torch.cat(video_means).cpu().tolist()         # the only device->host copies of the pass
    masked = torch.cat(masked_means).cpu().tolist() if masked_means else []
    average_loss = float(loss_sum.item()) / n_seen
    ops.check_status(compute_device, "evaluation pass")   # the copies above synchronised: a timed-out recurrence raises here
    mean_iou = _nanmean(video)
    containment = _nanmean(masked) if masked else float("nan")
    return average_loss, mean_iou, containment
