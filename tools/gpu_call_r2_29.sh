#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py tests/test_gpu_zz_regressions.py -m gpu -q -p no:cacheprovider -k "attention or transformer" --tb=line > gpurun_out/r02_29_tests.log 2>&1; tail -4 gpurun_out/r02_29_tests.log | cut -c1-250
timeout 300 python tools/attn_time.py 2>&1 | grep fused
timeout 300 python tools/profile_step.py --model transformer_lstm --steps 1 > /dev/null 2>&1
python - <<'PY'
import torch, time, sys
sys.path.insert(0, '.')
import bench
from objectpermanence_b200 import ops
dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
peaks, _ = bench.measured_peaks()
from objectpermanence_b200.models_factory import ModelsFactory
from objectpermanence_b200.synthetic import make_batch
from objectpermanence_b200.training import TrainingStep
for train in (True, False):
    torch.manual_seed(0)
    model = ModelsFactory.get_model("transformer_lstm", bench.TRANSFORMER_CFG).to(dev); model.train(train)
    step = TrainingStep(model, "transformer_lstm")
    b, l, _ = make_batch(32, 300, 5, seed=4321)
    b, l = torch.from_numpy(b).to(dev), torch.from_numpy(l).to(dev)
    ms = bench._event_time(lambda: step.forward_backward(b, l), 5, 3, flush)
    print(f"transformer_lstm [32,300] train={train}: {ms:.3f} ms/step = {32/(ms*1e-3):.0f} videos/s")
PY
