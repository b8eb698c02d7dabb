"""CPU tests of the input-pipeline cache: it must hand out exactly what a torch DataLoader over the reference's
dataset objects yields (sample structure of baselines/datasets.py:449,508,567,600)."""
import numpy as np
import torch
from torch.utils import data

from objectpermanence_b200.data_cache import CachedDataset
from objectpermanence_b200.synthetic import make_batch


class _ReferenceShapedDataset(data.Dataset):
    """Stand-in with the reference's item structure; `with_mask=False` mimics the datasets that yield an empty mask."""

    def __init__(self, n, T=12, F=6, with_mask=True, seed=0):
        self.boxes, self.labels, self.mask = make_batch(n, T, F, seed=seed)
        self.with_mask = with_mask
        self.reads = 0

    def __len__(self):
        return len(self.boxes)

    def __getitem__(self, i):
        self.reads += 1
        mask = torch.from_numpy(self.mask[i]) if self.with_mask else torch.empty(0)
        index_to_track = torch.full((self.boxes.shape[1],), i % 15, dtype=torch.int64)
        return ((torch.from_numpy(self.boxes[i]), index_to_track), (torch.from_numpy(self.labels[i]), mask), f"video_{i:04d}")


def test_batches_equal_dataloader_collation():
    ds = _ReferenceShapedDataset(11)
    cache = CachedDataset(ds, pin=False)
    assert ds.reads == 11 and len(cache) == 11
    loader = data.DataLoader(ds, batch_size=4)
    for got, want in zip(cache.batches(4), loader):
        (gb, gi), (gl, gm), gn = got
        (wb, wi), (wl, wm), wn = want
        assert torch.equal(gb, wb) and torch.equal(gi, wi) and torch.equal(gl, wl) and torch.equal(gm, wm)
        assert list(gn) == list(wn)
    assert ds.reads == 22          # the loader re-read every sample, the cache none
    assert sum(1 for _ in cache.batches(4)) == 3 and sum(1 for _ in cache.batches(4, drop_last=True)) == 2
    (b0, _), _, name0 = cache[3]
    assert torch.equal(b0, torch.from_numpy(ds.boxes[3])) and name0 == "video_0003"


def test_empty_mask_datasets_and_shuffle_cover_every_sample_once():
    ds = _ReferenceShapedDataset(9, with_mask=False, seed=3)
    cache = CachedDataset(ds, pin=False)
    assert cache.mask is None
    seen = []
    for (_, _), (labels, mask), names in cache.batches(2, shuffle=True, seed=5):
        assert mask.numel() == 0 and mask.shape[0] == labels.shape[0]
        seen += names
    assert sorted(seen) == [f"video_{i:04d}" for i in range(9)]
    again = [n for *_, names in cache.batches(2, shuffle=True, seed=5) for n in names]
    other = [n for *_, names in cache.batches(2, shuffle=True, seed=6) for n in names]
    assert again == seen and other != seen


def test_rank_shards_are_disjoint_and_equal():
    cache = CachedDataset(_ReferenceShapedDataset(16, seed=1), pin=False)
    per_rank = [[n for *_, names in cache.batches(2, shuffle=True, seed=2, rank=r, world=4) for n in names] for r in range(4)]
    assert all(len(p) == 4 for p in per_rank)
    assert sorted(sum(per_rank, [])) == [f"video_{i:04d}" for i in range(16)]
    # every rank's k-th batch comes from the same global batch
    g0 = [set(p[:2]) for p in per_rank]
    order = [n for *_, names in cache.batches(8, shuffle=True, seed=2) for n in names][:8]
    assert set().union(*g0) == set(order)
