"""The per-mini-batch step of the reference training loop (baselines/training_main.py:175-217)
with the model call, the loss and the backward on sm_100a kernels.

    step = TrainingStep(model, model_name)           # optionally reducer=FlatGradAllReducer(...)
    total, pred, cons = step(boxes, labels, mask)    # device tensors or pinned host tensors

The reference's three per-step `.item()` host syncs (training_main.py:212-214) are replaced by
one 12-byte device->host read of the (total, prediction, consistency) vector, plus the 16 status bytes of the
persistent kernels (ops.status_page): a time-out raises OpnError here instead of training on garbage.

`step.pipelined(boxes, labels, mask)` is the asynchronous form ("async logging"): the inputs go to
the device on a copy stream (double buffered, so the copy of step k+1 overlaps the kernels of step k), the loss
vector of the step is copied to a pinned slot without waiting, and the call returns the loss of the PREVIOUS step
(None on the first call); `step.drain()` returns the last one.  The host never waits for the step it has just
enqueued, so the GPU queue stays full.
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
        self._status_host = torch.zeros(4, dtype=torch.int32).pin_memory() if self.device.type == "cuda" else None
        self._status = ops.status_page(self.device)[:4] if self.device.type == "cuda" else None
        # pipelined form: two slots of (device input buffers, pinned loss vector, completion event)
        self._slots = None
        self._copy_stream = None
        self._k = 0
        self._pending = None

    def forward_backward(self, boxes: torch.Tensor, labels: torch.Tensor, mask: Optional[torch.Tensor] = None):
        """zero_grad + forward + loss + backward (+ gradient all-reduce).  Device tensors in, the
        3-vector (total, prediction, consistency) as a device tensor out; nothing synchronises."""
        for p in self.model.parameters():
            p.grad = None
        head = getattr(self.model, getattr(self.model, "head_name", ""), None)
        if head is not None and hasattr(self.model, "trunk"):
            # bbox head + loss + their backward as one pass over the hidden states (ops.head_loss) when the shape allows
            h, _ = self.model.trunk(boxes)
            if ops.head_loss_available(h, head.weight, head.bias) and head.weight.requires_grad and h.requires_grad:
                _, loss3, dh, dw = ops.head_loss(h, head.weight, labels, mask, _is_no_labels(self.model_name))
                head.weight.grad = dw
                h.backward(dh)
                if self.reducer is not None:
                    self.reducer.reduce()
                if self.optimizer is not None:
                    self.optimizer.step()
                return loss3
            y = head(h)
        else:
            out = self.model(boxes)
            y = out[0] if _is_double_output(self.model_name) else out
        # the loss launch also writes d total / dy: seed the backward pass with it directly (loss3[0].backward() would
        # add autograd's own fill / select / multiply kernels in front of it)
        loss3, dy = ops.loss_and_grad(y, labels, mask, _is_no_labels(self.model_name))
        y.backward(dy)
        if self.reducer is not None:
            self.reducer.reduce()
        if self.optimizer is not None:
            self.optimizer.step()
        return loss3

    def __call__(self, boxes: torch.Tensor, labels: torch.Tensor, mask: Optional[torch.Tensor] = None):
        """End-to-end step from HOST buffers: async H2D of the inputs, forward_backward, and a
        device->host read of the loss vector (the only synchronisation)."""
        dev = self.device
        boxes_d = boxes.to(dev, non_blocking=True)
        labels_d = labels.to(dev, non_blocking=True)
        mask_d = mask.to(dev, non_blocking=True) if mask is not None else None
        loss3 = self.forward_backward(boxes_d, labels_d, mask_d)
        self._loss_host.copy_(loss3, non_blocking=True)
        self._status_host.copy_(self._status, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        ops.raise_if_failed(self._status_host, self.device, "training step")
        return tuple(float(v) for v in self._loss_host)

    # ---- pipelined form -----------------------------------------------------------------------------------
    def _slot(self, boxes, labels, mask):
        if self._slots is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._slots = []
            for _ in range(2):
                self._slots.append({
                    "boxes": torch.empty(boxes.shape, dtype=boxes.dtype, device=self.device),
                    "labels": torch.empty(labels.shape, dtype=labels.dtype, device=self.device),
                    "mask": torch.empty(mask.shape, dtype=mask.dtype, device=self.device) if mask is not None else None,
                    "loss": torch.empty(3, dtype=torch.float32).pin_memory(),
                    "status": torch.zeros(4, dtype=torch.int32).pin_memory(),
                    "copied": torch.cuda.Event(), "done": torch.cuda.Event(), "free": torch.cuda.Event(),
                })
        slot = self._slots[self._k & 1]
        if slot["boxes"].shape != boxes.shape or slot["labels"].shape != labels.shape:
            raise RuntimeError("pipelined steps need a fixed batch shape (use drop_last / the synchronous call for the tail)")
        return slot

    def pipelined(self, boxes: torch.Tensor, labels: torch.Tensor, mask: Optional[torch.Tensor] = None):
        """Enqueue one step from pinned HOST tensors and return the loss triple of the previous step (None at first)."""
        slot = self._slot(boxes, labels, mask)
        main = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self._copy_stream):
            if self._k >= 2:
                self._copy_stream.wait_event(slot["free"])      # the step that last read these buffers has finished
            slot["boxes"].copy_(boxes, non_blocking=True)
            slot["labels"].copy_(labels, non_blocking=True)
            if mask is not None:
                slot["mask"].copy_(mask, non_blocking=True)
            slot["copied"].record(self._copy_stream)
        main.wait_event(slot["copied"])
        loss3 = self.forward_backward(slot["boxes"], slot["labels"], slot["mask"])
        slot["free"].record(main)
        slot["loss"].copy_(loss3, non_blocking=True)
        slot["status"].copy_(self._status, non_blocking=True)
        slot["done"].record(main)
        previous = self._pending
        self._pending = slot
        self._k += 1
        if previous is None:
            return None
        previous["done"].synchronize()
        ops.raise_if_failed(previous["status"], self.device, "training step")
        return tuple(float(v) for v in previous["loss"])

    def drain(self):
        """Wait for the last pipelined step and return its loss triple."""
        if self._pending is None:
            return None
        self._pending["done"].synchronize()
        pending, self._pending = self._pending, None
        ops.raise_if_failed(pending["status"], self.device, "training step")
        return tuple(float(v) for v in pending["loss"])
