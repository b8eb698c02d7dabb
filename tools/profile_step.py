"""A short run of the bench step for Nsight Compute (bench.py's 400-step spin-up is too long under a profiler):
5 warm-up steps, then `--steps` steps of zero_grad + forward + loss + backward at the headline config, or one LSTM layer
forward + backward at a large batch (--tc).  Usage (GPU box):
    ncu --metrics gpu__time_duration.sum --clock-control none -s <skip> -c <n> --csv --log-file out.csv python tools/profile_step.py
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectpermanence_b200 import _lib, ops
from objectpermanence_b200.models_factory import ModelsFactory
from objectpermanence_b200.synthetic import make_batch
from objectpermanence_b200.training import TrainingStep

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--model", default="opnet")
ap.add_argument("--tc", action="store_true", help="one H=512 LSTM layer at B=256 through the batch-wide tcgen05 kernels")
args = ap.parse_args()
dev = torch.device("cuda:0")
if args.tc:
    B, T, H = 256, 300, 512
    g = torch.Generator().manual_seed(0)
    x = torch.rand(B, T, 6, generator=g).to(dev)
    w_ih = ((torch.rand(4 * H, 6, generator=g) * 2 - 1) / H ** 0.5).to(dev).requires_grad_(True)
    w_hh = ((torch.rand(4 * H, H, generator=g) * 2 - 1) / H ** 0.5).to(dev).requires_grad_(True)
    dh = (torch.rand(B, T, H, generator=g) * 0.01).to(dev)
    for _ in range(2 + args.steps):
        ops.lstm_layer(x, w_ih, w_hh).backward(dh)
    torch.cuda.synchronize()
    print("launches", _lib.launch_count())
else:
    cfg = {"object_to_track_pred_dim": 15, "object_to_track_hidden_dim": 256, "videos_hidden_dim": 512}
    if args.model == "transformer_lstm":
        cfg = {"boxes_features_dim": 256, "num_attention_heads": 2, "num_attention_layers": 2, "num_lstm_layers": 2, "lstm_hidden_dim": 512}
    torch.manual_seed(0)
    model = ModelsFactory.get_model(args.model, cfg).to(dev).train()
    step = TrainingStep(model, args.model)
    b, l, _ = make_batch(args.batch, 300, 5 if args.model == "transformer_lstm" else 6, seed=1234)
    b, l = torch.from_numpy(b).to(dev), torch.from_numpy(l).to(dev)
    for _ in range(2 if args.model == "transformer_lstm" else 5):
        step.forward_backward(b, l)
    torch.cuda.synchronize()
    n0 = _lib.launch_count()
    for _ in range(args.steps):
        step.forward_backward(b, l)
    torch.cuda.synchronize()
    print("library launches per step", (_lib.launch_count() - n0) / args.steps)
