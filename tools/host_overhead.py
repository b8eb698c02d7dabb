"""Host time to ENQUEUE one training step (no synchronisation inside the loop) against the GPU time of the step, and a
cProfile of the enqueue path: with the GPU step under 2 ms the Python / ctypes side is what an end-to-end loop can be bound by."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from objectpermanence_b200.models_factory import ModelsFactory
from objectpermanence_b200.synthetic import make_batch
from objectpermanence_b200.training import TrainingStep
dev = torch.device("cuda:0")
torch.manual_seed(0)
cfg = {"object_to_track_pred_dim": 15, "object_to_track_hidden_dim": 256, "videos_hidden_dim": 512}
model = ModelsFactory.get_model("opnet", cfg).to(dev).train()
step = TrainingStep(model, "opnet")
b, l, _ = make_batch(32, 300, 6, seed=1)
b, l = torch.from_numpy(b).to(dev), torch.from_numpy(l).to(dev)
for _ in range(50): step.forward_backward(b, l)
torch.cuda.synchronize()
N = 200
t0 = time.perf_counter()
for _ in range(N): step.forward_backward(b, l)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3 * (t1 - t0) / N:.3f} ms per step; until the GPU drained {1e3 * (t2 - t0) / N:.3f} ms per step")
pr = cProfile.Profile()
pr.enable()
for _ in range(100): step.forward_backward(b, l)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
