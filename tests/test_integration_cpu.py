"""CPU: the integration plumbing around the hot path, run against the UNMODIFIED reference staged in baseline/_ref
(oracle/stage_reference.py; skipped where it has not been staged).

* tools/run_reference_main.py drives the reference's own `main.py training` (baselines/training_main.py:120-252) on a
  fabricated 3-video CATER data set in the reference's pkl / json / tsv formats -- here with --stock (the reference's own
  modules on the CPU; the swapped-in B200 modules need a GPU: tests/test_gpu_integration.py);
* CachedDataset is fed from the reference's real dataset classes (baselines/datasets.py:419-600) and must hand out exactly
  what a DataLoader over them yields.
"""
import importlib
import json
import os
import sys

import pytest
import torch
from torch.utils import data

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tools"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle import stage_reference  # noqa: E402
import fabricate  # noqa: E402

pytestmark = pytest.mark.skipif(stage_reference.staged_root() is None, reason="reference not staged in baseline/_ref")


@pytest.fixture(scope="module")
def dataset_root(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("cater"))
    fabricate.fabricate(root, n_videos=3, seed=1)
    return root


def test_launcher_runs_the_reference_training_main_on_fabricated_files(dataset_root, tmp_path, capsys):
    import run_reference_main
    model_cfg = tmp_path / "model.json"
    model_cfg.write_text(json.dumps({"object_to_track_pred_dim": 15, "object_to_track_hidden_dim": 32, "videos_hidden_dim": 32}))
    train_cfg = tmp_path / "train.json"
    train_cfg.write_text(json.dumps(fabricate.training_config(dataset_root, "cpu", str(tmp_path / "ckpt"))))
    torch.manual_seed(0)
    run_reference_main.launch(["training", "--model_type", "opnet", "--model_config", str(model_cfg),
                               "--training_config", str(train_cfg)], stock=True)
    out = capsys.readouterr().out
    assert "Epoch 1 Training Set: Loss" in out and "Epoch 1 Dev Set: Loss" in out
    assert "Train Epoch: 1 [2/3" in out     # two mini-batches of the 3 fabricated videos, print_step = 1


@pytest.mark.parametrize("kind", ["train6", "infer5"])
def test_cached_dataset_equals_dataloader_over_the_reference_datasets(dataset_root, kind):
    from objectpermanence_b200.data_cache import CachedDataset
    stage_reference.import_reference()
    ref_datasets = importlib.import_module("baselines.datasets")
    samples, labels = os.path.join(dataset_root, "od_perception"), os.path.join(dataset_root, "labels")
    if kind == "train6":
        ds = ref_datasets.Cater6TracksForObjectsTrainingDataset(samples, labels, os.path.join(dataset_root, "containment_annotations.txt"))
    else:
        ds = ref_datasets.Cater5TracksForObjectsInferenceDataset(samples, labels)
    cache = CachedDataset(ds, pin=False)
    assert len(cache) == 3 and cache.boxes.shape == (3, 300, 15, 6 if kind == "train6" else 5)
    assert cache.has_mask == (kind == "train6")
    n = 0
    for got, want in zip(cache.batches(2), data.DataLoader(ds, batch_size=2)):
        (gb, gi), (gl, gm), gn = got
        (wb, wi), (wl, wm), wn = want
        assert torch.equal(gb, wb) and torch.equal(gi, wi) and torch.equal(gl, wl) and list(gn) == list(wn)
        if kind == "train6":
            assert gm.dtype == torch.bool and torch.equal(gm, wm) and gm.any()
        else:
            assert gm.numel() == 0 and wm.numel() == 0
        # the fabricated files exercise the reference's padding rules: snitch slot first, cone padding rows [0,0,0,0,0,1]
        if kind == "train6":
            assert ((gb[..., 4] == 0) & (gb[..., 5] == 1)).any() and (gb[:, :, 0, 4] == 0).any()
        n += len(gn)
    assert n == 3
