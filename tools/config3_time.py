"""BASELINE config 3 (transformer_lstm [32,300]) step time, train and eval mode, timed like the bench readings."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from objectpermanence_b200.models_factory import ModelsFactory
from objectpermanence_b200.synthetic import make_batch
from objectpermanence_b200.training import TrainingStep
dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
for train in (True, False):
    torch.manual_seed(0)
    model = ModelsFactory.get_model("transformer_lstm", bench.TRANSFORMER_CFG).to(dev); model.train(train)
    step = TrainingStep(model, "transformer_lstm")
    b, l, _ = make_batch(32, 300, 5, seed=4321)
    b, l = torch.from_numpy(b).to(dev), torch.from_numpy(l).to(dev)
    ms = bench._event_time(lambda: step.forward_backward(b, l), 5, 3, flush)
    print(f"transformer_lstm [32,300] train={train}: {ms:.3f} ms/step = {32 / (ms * 1e-3):.0f} videos/s", flush=True)
