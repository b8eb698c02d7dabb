"""A/B of host-side switches of the OPNet step in ONE process: ms per forward+loss+backward step (CUDA events, 256 MB L2
flush between steps) with the weight-gradient contractions in line (OPN_OPNET_WGRAD_OVERLAP=0) and on two streams (=1),
plus the largest gradient difference between the two (split-K atomics reorder sums, so not bit-equal)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectpermanence_b200.models_factory import ModelsFactory
from objectpermanence_b200.synthetic import make_batch
from objectpermanence_b200.training import TrainingStep

dev = torch.device("cuda:0")
cfg = {"object_to_track_pred_dim": 15, "object_to_track_hidden_dim": 256, "videos_hidden_dim": 512}
torch.manual_seed(0)
model = ModelsFactory.get_model("opnet", cfg).to(dev).train()
step = TrainingStep(model, "opnet")
b, l, _ = make_batch(32, 300, 6, seed=1234)
boxes, labels = torch.from_numpy(b).to(dev), torch.from_numpy(l).to(dev)
flush = torch.empty(64 * 1024 * 1024, device=dev)
STEPS = int(os.environ.get("STEPS", "30"))


def run(mode):
    os.environ["OPN_OPNET_WGRAD_OVERLAP"] = mode
    for _ in range(10):
        step.forward_backward(boxes, labels)
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(STEPS):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step.forward_backward(boxes, labels); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / STEPS, {k: v.grad.detach().clone() for k, v in model.named_parameters()}


for _ in range(200):
    step.forward_backward(boxes, labels)     # spin-up (clocks, lazy module loading)
torch.cuda.synchronize()
res = {}
for rep in range(2):
    for mode in ("0", "1"):
        ms, grads = run(mode)
        res[mode] = grads
        print(f"rep {rep} OPN_OPNET_WGRAD_OVERLAP={mode}: {ms:.4f} ms/step = {32 / ms * 1e3:.0f} videos/s", flush=True)
worst = max(((res["0"][k] - res["1"][k]).abs().max().item() / max(1e-12, res["0"][k].abs().max().item()), k) for k in res["0"])
print(f"largest relative gradient difference between the modes: {worst[0]:.2e} ({worst[1]})")
