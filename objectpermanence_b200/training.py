"""The per-mini-batch step of the reference training loop (baselines/training_main.py:175-217)
with the model call, the loss and the backward on sm_100a kernels.

    step = TrainingStep(model, model_name)           # optionally reducer=FlatGradAllReducer(...)
    total, pred, cons = step(boxes, labels, mask)    # device tensors or pinned host tensors

The reference's three per-step `.item()` host syncs (training_main.py:212-214) are replaced by
one 12-byte device->host read of the (total, prediction, consistency) vector.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .data_parallel import FlatGradAllReducer
from .supported_models import DOUBLE_OUTPUT_MODELS, NO_LABELS_MODELS


def _is_no_labels(model_name: str) -> bool:
    return model_name in NO_LABELS_MODELS or model_name == "opent_no_labels"


def _is_double_output(model_name: str) -> bool:
    return model_name in DOUBLE_OUTPUT_MODELS or model_name == "opent_no_labels"


class TrainingStep:
    def __init__(self, model: torch.nn.Module, model_name: str, reducer: Optional[FlatGradAllReducer] = None,
                 optimizer: Optional[torch.optim.Optimizer] = None):
        self.model = model
        self.model_name = model_name
        self.reducer = reducer
        self.optimizer = optimizer
        self.device = next(model.parameters()).device
        self._loss_host = torch.empty(3, dtype=torch.float32).pin_memory() if self.device.type == "cuda" else None

    def forward_backward(self, boxes: torch.Tensor, labels: torch.Tensor, mask: Optional[torch.Tensor] = None):
        """zero_grad + forward + loss + backward (+ gradient all-reduce).  Device tensors in, the
        3-vector (total, prediction, consistency) as a device tensor out; nothing synchronises."""
        for p in self.model.parameters():
            p.grad = None
        out = self.model(boxes)
        y = out[0] if _is_double_output(self.model_name) else out
        loss3 = ops.training_loss(y, labels, mask, _is_no_labels(self.model_name))
        loss3[0].backward()
        if self.reducer is not None:
            self.reducer.reduce()
        if self.optimizer is not None:
            self.optimizer.step()
        return loss3.detach()

    def __call__(self, boxes: torch.Tensor, labels: torch.Tensor, mask: Optional[torch.Tensor] = None):
        """End-to-end step from HOST buffers: async H2D of the inputs, forward_backward, and a
        device->host read of the loss vector (the only synchronisation)."""
        dev = self.device
        boxes_d = boxes.to(dev, non_blocking=True)
        labels_d = labels.to(dev, non_blocking=True)
        mask_d = mask.to(dev, non_blocking=True) if mask is not None else None
        loss3 = self.forward_backward(boxes_d, labels_d, mask_d)
        self._loss_host.copy_(loss3, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return tuple(float(v) for v in self._loss_host)
