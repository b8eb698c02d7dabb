"""Batch-sharded data parallelism: one process per GPU, ONE all-reduce of a flat gradient
buffer per optimiser step (NCCL over NVLink on the box; gloo in the CPU tests).

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

    After ``reduce()`` every ``p.grad`` is a view into one contiguous fp32 buffer holding the
    rank-averaged gradient, so the optimiser sees identical values on every rank."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group=None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = group
        self.numel = sum(p.numel() for p in self.params)
        self.flat: Optional[torch.Tensor] = None
        self.collectives = 0

    @property
    def nbytes(self) -> int:
        return 4 * self.numel

    def reduce(self) -> torch.Tensor:
        grads = [(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in self.params]
        if self.flat is None or self.flat.device != grads[0].device:
            self.flat = torch.empty(self.numel, dtype=torch.float32, device=grads[0].device)
        torch.cat(grads, out=self.flat)
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.mul_(1.0 / world)
            self.collectives += 1
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n
        return self.flat


def broadcast_parameters(params: Iterable[torch.nn.Parameter], src: int = 0, group=None) -> None:
    """Make every rank start from rank `src`'s weights."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for p in params:
        dist.broadcast(p.data, src=src, group=group)
