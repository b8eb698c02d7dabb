"""GPU: the reference's own `main.py training` and `main.py inference` (baselines/training_main.py:120-252,
baselines/inference_main.py:162-257), UNMODIFIED and staged in baseline/_ref, driven by tools/run_reference_main.py with
the learned models swapped for the B200 implementation -- the "main.py training|inference and the JSON files in configs/
run unchanged" clause of the north star -- on a fabricated 3-video data set in the reference's file formats.
The same command with --stock (the reference's own modules on the CPU) is the oracle: same seed, same files."""
import json
import os
import re
import sys

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tools"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle import stage_reference  # noqa: E402
import fabricate  # noqa: E402

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(stage_reference.staged_root() is None, reason="reference not staged in baseline/_ref")]

EPOCH_LINE = re.compile(r"Epoch 1 (Training|Dev) Set: Loss ([0-9.]+), Mean IoU ([0-9.naN]+), Mask Mean Iou ([0-9.naN]+)")


@pytest.fixture(scope="module")
def dataset_root(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("cater"))
    fabricate.fabricate(root, n_videos=3, seed=2, with_videos=True)
    return root


def _shipped_config(name):
    return os.path.join(stage_reference.staged_root(), "configs", name)


def test_reference_training_main_runs_on_the_b200_models_and_agrees_with_its_own_cpu_run(cuda_device, dataset_root, tmp_path, capsys):
    import run_reference_main
    results = {}
    for arm, device in (("stock", "cpu"), ("b200", "cuda:0")):
        train_cfg = tmp_path / f"train_{arm}.json"
        train_cfg.write_text(json.dumps(fabricate.training_config(dataset_root, device, str(tmp_path / f"ckpt_{arm}"))))
        torch.manual_seed(0)
        run_reference_main.launch(["training", "--model_type", "opnet", "--model_config", _shipped_config("opnet_model_config.json"),
                                   "--training_config", str(train_cfg)], stock=(arm == "stock"))
        out = capsys.readouterr().out
        lines = {m.group(1): (float(m.group(2)), float(m.group(3))) for m in EPOCH_LINE.finditer(out)}
        assert set(lines) == {"Training", "Dev"}, out
        results[arm] = lines
    for split in ("Training", "Dev"):
        (l0, i0), (l1, i1) = results["stock"][split], results["b200"][split]
        assert abs(l0 - l1) <= 2e-4, (split, l0, l1)          # printed with 4 decimals after two Adam steps
        assert abs(i0 - i1) <= 1e-3, (split, i0, i1)          # mean IoU equal to 3 decimals


def test_reference_inference_main_writes_the_same_boxes_with_the_b200_models(cuda_device, dataset_root, tmp_path, capsys):
    import run_reference_main
    from objectpermanence_b200 import inference
    from objectpermanence_b200.models_factory import ModelsFactory
    stage_reference.import_reference()
    with open(_shipped_config("opnet_model_config.json")) as f:
        model_cfg = json.load(f)
    torch.manual_seed(5)
    model = ModelsFactory.get_model("opnet", model_cfg)
    ckpt = str(tmp_path / "opnet.pth")
    torch.save(model.state_dict(), ckpt)
    boxes = {}
    for arm, device in (("stock", "cpu"), ("b200", "cuda:0")):
        out_dir = tmp_path / f"results_{arm}"
        out_dir.mkdir()
        cfg = tmp_path / f"inference_{arm}.json"
        cfg.write_text(json.dumps(fabricate.inference_config(dataset_root, device, ckpt)))
        if arm == "stock":
            # the reference hard-codes map_location "cuda:0" for its own weights (models_factory.py:77); the CPU oracle arm
            # loads them itself
            import baselines.models_factory as ref_factory
            original = torch.load
            torch.load = lambda f, *a, **k: original(f, map_location="cpu")
            try:
                run_reference_main.launch(["inference", "--model_type", "opnet", "--results_dir", str(out_dir),
                                           "--inference_config", str(cfg), "--model_config", _shipped_config("opnet_model_config.json")],
                                          stock=True)
            finally:
                torch.load = original
        else:
            run_reference_main.launch(["inference", "--model_type", "opnet", "--results_dir", str(out_dir),
                                       "--inference_config", str(cfg), "--model_config", _shipped_config("opnet_model_config.json")])
        capsys.readouterr()
        files = sorted(p for p in os.listdir(out_dir) if p.endswith("_bb.json"))
        assert len(files) == 3, os.listdir(out_dir)
        boxes[arm] = {p: np.array(json.load(open(out_dir / p))) for p in files}
    for name in boxes["stock"]:
        a, b = boxes["stock"][name], boxes["b200"][name]
        assert a.shape == b.shape == (300, 4)
        assert np.abs(a - b).max() <= 1 and (a == b).mean() >= 0.99     # int truncation may flip a pixel on 1e-7 differences
    # the package's own writer (no video loop) leaves the same files as the reference's
    import baselines.datasets as ref_datasets
    ds = ref_datasets.Cater6TracksForObjectsInferenceDataset(os.path.join(dataset_root, "od_perception"), os.path.join(dataset_root, "labels"))
    loader = torch.utils.data.DataLoader(ds, batch_size=2)
    model = ModelsFactory.get_model("opnet", model_cfg, ckpt)
    own_dir = tmp_path / "results_own"
    written = inference.run_inference("opnet", model, cuda_device, loader, str(own_dir))
    assert len(written) == 3
    for name, path in written.items():
        ref_file = tmp_path / "results_b200" / (name + "_bb.json")
        assert json.load(open(path)) == json.load(open(ref_file))
