"""Adam as ONE kernel launch per step over a flat parameter buffer (csrc/opn_train_eval.cu: opn_adam_step).

Drop-in for the reference's ``torch.optim.Adam(model.parameters(), lr=learning_rate)``
(baselines/training_main.py:150) and the ``optimizer.step()`` of :217; ``ReduceLROnPlateau`` (:151,247) works on it
unchanged because it is a ``torch.optim.Optimizer`` with the usual ``param_groups[0]["lr"]``.

    model = ModelsFactory.get_model(name, cfg).to("cuda")      # flatten AFTER the move to the device
    optimizer = FusedAdam(model.parameters(), lr=1e-3)

The constructor re-points every ``p.data`` at a slice of one contiguous fp32 buffer (``state_dict`` /
``load_state_dict`` of the model keep working: the parameters are views).  If the gradients are already views of one
flat buffer in parameter order (``FlatGradAllReducer.reduce()`` leaves them so) the step reads that buffer directly;
otherwise they are concatenated first.
"""
from __future__ import annotations

from typing import Iterable, Optional, Tuple

import torch

from . import _lib


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 0.0):
        params = [p for p in params if p.requires_grad]
        if not params:
            raise ValueError("FusedAdam got no trainable parameters")
        for p in params:
            if not p.is_cuda or p.dtype != torch.float32:
                raise RuntimeError("FusedAdam runs on CUDA float32 parameters only (no CPU path): move the model first")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._params = params
        dev = params[0].device
        # every slice starts on a 16-byte boundary so that the kernel's float4 path applies to any layout
        self._offsets, off = [], 0
        for p in params:
            self._offsets.append(off)
            off += (p.numel() + 3) // 4 * 4
        self.numel = off
        self.flat_params = torch.zeros(off, dtype=torch.float32, device=dev)
        for p, o in zip(params, self._offsets):
            view = self.flat_params[o:o + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
        self.exp_avg = torch.zeros_like(self.flat_params)
        self.exp_avg_sq = torch.zeros_like(self.flat_params)
        self._flat_grads = torch.zeros_like(self.flat_params)
        self.steps = 0

    def _gather_grads(self) -> torch.Tensor:
        """The gradients as one buffer laid out like ``flat_params`` (zero where a parameter has no gradient)."""
        first = self._params[0].grad
        if first is not None and self.numel == sum(p.numel() for p in self._params):
            base = first.data_ptr()
            in_place = all(p.grad is not None and p.grad.is_contiguous() and p.grad.data_ptr() == base + 4 * o
                           for p, o in zip(self._params, self._offsets))
            if in_place and base % 16 == 0:
                # already one flat buffer in parameter order (FlatGradAllReducer): wrap it without copying
                return first.new_empty(0).set_(first.untyped_storage(), first.storage_offset(), (self.numel,), (1,))
        for p, o in zip(self._params, self._offsets):
            dst = self._flat_grads[o:o + p.numel()]
            if p.grad is None:
                dst.zero_()
            else:
                dst.copy_(p.grad.reshape(-1))
        return self._flat_grads

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        group = self.param_groups[0]
        self._check_layout()
        self.steps += 1
        grads = self._gather_grads()
        rc = _lib.load().opn_adam_step(self.numel, self.flat_params.data_ptr(), grads.data_ptr(), self.exp_avg.data_ptr(),
                                       self.exp_avg_sq.data_ptr(), float(group["lr"]), float(group["betas"][0]),
                                       float(group["betas"][1]), float(group["eps"]), float(group["weight_decay"]),
                                       self.steps, torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "opn_adam_step")
        return loss

    def _check_layout(self) -> None:
        """`model.to(...)` / `.float()` after construction re-allocates the parameters: the flat views would be stale."""
        base = self.flat_params.data_ptr()
        for p, o in zip(self._params, self._offsets):
            if p.data_ptr() != base + 4 * o:
                raise RuntimeError("FusedAdam: a parameter no longer lives in the flat buffer (the model was moved or cast "
                                   "after the optimiser was built); construct FusedAdam after model.to(device)")

    def state_dict(self):
        """The layout of ``torch.optim.Adam.state_dict()``: ``state[i] = {step, exp_avg, exp_avg_sq}`` per parameter (copies
        of the slices of the flat moment buffers) and ``param_groups[0]["params"] = [0..n)``, so checkpoint helpers and
        ``torch.optim.Adam.load_state_dict`` accept it and vice versa."""
        state = {}
        if self.steps > 0:
            for i, (p, o) in enumerate(zip(self._params, self._offsets)):
                n = p.numel()
                state[i] = {"step": torch.tensor(float(self.steps)),
                            "exp_avg": self.exp_avg[o:o + n].view_as(p).clone(),
                            "exp_avg_sq": self.exp_avg_sq[o:o + n].view_as(p).clone()}
        groups = []
        for g in self.param_groups:
            packed = {k: v for k, v in g.items() if k != "params"}
            packed["params"] = list(range(len(self._params)))
            groups.append(packed)
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, state_dict):
        if "state" not in state_dict or "param_groups" not in state_dict:
            raise ValueError("FusedAdam.load_state_dict expects the torch.optim layout {'state': ..., 'param_groups': ...}")
        groups = state_dict["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(self._params):
            raise ValueError("FusedAdam.load_state_dict: expected one parameter group with "
                             f"{len(self._params)} parameters")
        steps = set()
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        for key, entry in state_dict["state"].items():
            i = int(key)
            p, o = self._params[i], self._offsets[i]
            n = p.numel()
            if tuple(entry["exp_avg"].shape) != tuple(p.shape):
                raise ValueError(f"FusedAdam.load_state_dict: moment shape mismatch for parameter {i}")
            self.exp_avg[o:o + n].copy_(entry["exp_avg"].reshape(-1))
            self.exp_avg_sq[o:o + n].copy_(entry["exp_avg_sq"].reshape(-1))
            steps.add(int(float(entry["step"])))
        if len(steps) > 1:
            raise ValueError("FusedAdam.load_state_dict: parameters with different step counts are not supported "
                             "(one bias correction per launch)")
        self.steps = steps.pop() if steps else 0
        for g, saved in zip(self.param_groups, groups):
            g.update({k: v for k, v in saved.items() if k != "params"})
