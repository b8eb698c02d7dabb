"""Input pipeline: the reference's datasets parse one pickle + one JSON per sample and pad with Python loops on
every access, every epoch (baselines/datasets.py:419-600, consumed at training_main.py:155-159,180-181).  The padded
tensors are tiny (108 kB per CATER video, ~1 GB for the whole train split), so they are read ONCE through the
reference's own dataset object, kept as contiguous tensors -- pinned on the host or resident on the GPU -- and
mini-batches are cut from them with one gather per tensor.

    cache = CachedDataset(reference_dataset, device=torch.device("cuda"))      # one pass over the files
    for (boxes, index_to_track), (labels, mask), names in cache.batches(32, shuffle=True, seed=epoch):
        ...                                                                    # same structure the DataLoader yields

The sample structure is the reference's:
``((boxes [T,15,F] f32, index_to_track [T] i64), (labels [T,4] f32, mask [T,4] bool | empty), video_name)``.
"""
from __future__ import annotations

from typing import Iterator, List, Optional, Sequence

import torch


class CachedDataset:
    def __init__(self, dataset: Sequence, device: Optional[torch.device] = None, pin: bool = True):
        boxes, index, labels, masks, self.names = [], [], [], [], []
        for i in range(len(dataset)):
            (b, idx), (lab, m), name = dataset[i]
            boxes.append(torch.as_tensor(b))
            index.append(torch.as_tensor(idx))
            labels.append(torch.as_tensor(lab))
            masks.append(torch.as_tensor(m))
            self.names.append(name)
        self.has_mask = len(masks) > 0 and all(m.numel() > 0 for m in masks)
        self.boxes = torch.stack(boxes).contiguous()
        self.index_to_track = torch.stack(index).contiguous()
        self.labels = torch.stack(labels).contiguous()
        self.mask = torch.stack(masks).contiguous() if self.has_mask else None
        self.device = torch.device(device) if device is not None else None
        if self.device is not None and self.device.type == "cuda":
            self.boxes, self.index_to_track, self.labels = [t.to(self.device) for t in (self.boxes, self.index_to_track, self.labels)]
            if self.mask is not None:
                self.mask = self.mask.to(self.device)
        elif pin and torch.cuda.is_available():
            self.boxes, self.index_to_track, self.labels = [t.pin_memory() for t in (self.boxes, self.index_to_track, self.labels)]
            if self.mask is not None:
                self.mask = self.mask.pin_memory()

    def __len__(self) -> int:
        return len(self.names)

    def __getitem__(self, i: int):
        mask = self.mask[i] if self.mask is not None else torch.empty(0)
        return (self.boxes[i], self.index_to_track[i]), (self.labels[i], mask), self.names[i]

    @property
    def nbytes(self) -> int:
        tensors = [self.boxes, self.index_to_track, self.labels] + ([self.mask] if self.mask is not None else [])
        return sum(t.numel() * t.element_size() for t in tensors)

    def batches(self, batch_size: int, shuffle: bool = False, seed: int = 0, drop_last: bool = False,
                rank: int = 0, world: int = 1) -> Iterator:
        """Mini-batches in the DataLoader's collated structure.  With world > 1 every rank takes a disjoint, equally
        sized slice of each global batch of ``batch_size * world`` samples (batch-sharded data parallelism)."""
        n = len(self)
        if shuffle:
            order = torch.randperm(n, generator=torch.Generator().manual_seed(seed))
        else:
            order = torch.arange(n)
        global_bs = batch_size * world
        for start in range(0, n, global_bs):
            idx = order[start:start + global_bs]
            if len(idx) < global_bs and (drop_last or world > 1):
                if drop_last or len(idx) < world:
                    break
                per = len(idx) // world
                idx = idx[rank * per:(rank + 1) * per]
            elif world > 1:
                idx = idx[rank * batch_size:(rank + 1) * batch_size]
            sel = idx.to(self.boxes.device)
            mask = self.mask.index_select(0, sel) if self.mask is not None else torch.empty(len(idx), 0)
            names: List[str] = [self.names[i] for i in idx.tolist()]
            yield ((self.boxes.index_select(0, sel), self.index_to_track.index_select(0, sel)),
                   (self.labels.index_select(0, sel), mask), names)
