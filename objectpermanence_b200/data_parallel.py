"""Batch-sharded data parallelism: one process per GPU, the flat gradient buffer all-reduced once per optimiser step
(NCCL over NVLink on the box; gloo in the CPU tests) -- on NCCL in two buckets where the backward pass says part of the
buffer is final early (OPNet: the weights of LSTM2), so that most of the transfer hides behind the remaining kernels.

The reference trains on a single device (baselines/training_main.py:144,162); this is the
one parallel strategy the B200 build adds.  Videos are independent through
every op of the LSTM models and every loss term is a mean over samples, so with equal shards
the average of the rank gradients equals the gradient of the global batch.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def init_process_group_from_env(backend: Optional[str] = None) -> int:
    """torchrun-style initialisation (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*).  Returns world size."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 1
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group(backend=backend)
    return dist.get_world_size()


def shard_bounds(global_batch: int, rank: int, world: int):
    """Contiguous, equal video shards (global_batch must divide evenly so rank means average exactly)."""
    if global_batch % world != 0:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world}")
    per = global_batch // world
    return rank * per, (rank + 1) * per


class FlatGradAllReducer:
    """Averages the gradients of `params` across ranks with a single collective.

    The flat fp32 buffer exists from construction and its slices are registered as the landing slots of the weight
    gradients (``ops.register_grad_slots``): the backward kernels write into it directly, so ``reduce()`` is the
    all-reduce alone (``ncclAvg`` on NCCL: no scaling pass either) plus a copy for any gradient that did not land in its
    slot (a transposed product, an accumulated gradient).  After ``reduce()`` every ``p.grad`` is a view into the buffer
    holding the rank-averaged gradient, so the optimiser sees identical values on every rank."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group=None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = group
        self.numel = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat: torch.Tensor = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self._views: List[torch.Tensor] = []
        off = 0
        for p in self.params:
            self._views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.collectives = 0
        self.copies = 0          # gradients that had to be copied into their slot (statistics)
        self._early = None       # (work handle, first element, end) of a bucket whose all-reduce is already in flight
        if dev.type == "cuda":
            from . import ops
            ops.register_grad_slots(self.params, self.flat)
            if os.environ.get("OPN_DP_EARLY_BUCKET", "1") not in ("0", ""):
                ops.set_early_grads_hook(self._start_early_bucket)

    def close(self) -> None:
        """Forget the landing slots (the buffer stays valid for whoever still holds views of it)."""
        if self.flat.device.type == "cuda":
            from . import ops
            ops.unregister_grad_slots(self.params)
            if ops._early_grads_hook == self._start_early_bucket:
                ops.set_early_grads_hook(None)

    def _start_early_bucket(self, grads) -> None:
        """Backward hook (ops.set_early_grads_hook): `grads` are final and lie back to back in the flat buffer (the weight
        gradients of OPNet's LSTM2, 75 % of the bytes, ~0.1 ms before the backward pass ends): their all-reduce starts now,
        on NCCL's stream, while the remaining kernels run; reduce() handles the rest of the buffer and waits for it."""
        if self._early is not None or not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        if dist.get_backend(self.group) != "nccl":
            return
        base, esz = self.flat.data_ptr(), self.flat.element_size()
        spans = sorted(((g.data_ptr() - base) // esz, (g.data_ptr() - base) // esz + g.numel()) for g in grads)
        if spans[0][0] < 0 or spans[-1][1] > self.numel or any(a[1] != b[0] for a, b in zip(spans, spans[1:])):
            return      # not (contiguous) slices of this buffer: the single all-reduce of reduce() covers them
        a, b = spans[0][0], spans[-1][1]
        work = dist.all_reduce(self.flat[a:b], op=dist.ReduceOp.AVG, group=self.group, async_op=True)
        self._early = (work, a, b)
        self.collectives += 1

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 -- interpreter shutdown
            pass

    @property
    def nbytes(self) -> int:
        return 4 * self.numel

    def reduce(self) -> torch.Tensor:
        for p, view in zip(self.params, self._views):
            g = p.grad
            if g is None:
                view.zero_()
            elif g.data_ptr() != view.data_ptr() or not g.is_contiguous():
                view.copy_(g)
                self.copies += 1
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if world > 1 and self._early is not None:
            # the bucket started from inside the backward pass is in flight: the two ends of the buffer now, then wait for it
            work, a, b = self._early
            self._early = None
            rest = [self.flat[lo:hi] for lo, hi in ((0, a), (b, self.numel)) if hi > lo]
            if len(rest) > 1:
                # one grouped NCCL launch for both ends (two separate small all-reduces would cost two launch latencies)
                try:
                    from torch.distributed.distributed_c10d import _coalescing_manager
                    with _coalescing_manager(group=self.group, device=self.flat.device, async_ops=False):
                        for t in rest:
                            dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group)
                    rest = []
                except Exception:  # noqa: BLE001 -- no grouped launch on this torch / backend: one all-reduce per end (AVG is idempotent)
                    pass
            for t in rest:
                dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group)
            self.collectives += 1
            work.wait()
        elif world > 1:
            if dist.get_backend(self.group) == "nccl":
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
            else:   # gloo has no AVG
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
                self.flat.mul_(1.0 / world)
            self.collectives += 1
        for p, view in zip(self.params, self._views):
            p.grad = view
        return self.flat


def broadcast_parameters(params: Iterable[torch.nn.Parameter], src: int = 0, group=None) -> None:
    """Make every rank start from rank `src`'s weights."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for p in params:
        dist.broadcast(p.data, src=src, group=group)
