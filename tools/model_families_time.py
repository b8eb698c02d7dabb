"""fwd + loss + bwd time per step of every model family at the BASELINE shape [B=32, T=300] (shipped configs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectpermanence_b200 import ops, _lib
from objectpermanence_b200.models_factory import ModelsFactory
from objectpermanence_b200.synthetic import make_batch
dev = torch.device("cuda:0")
CONFIGS = {
    "opnet": ({"object_to_track_pred_dim": 15, "object_to_track_hidden_dim": 256, "videos_hidden_dim": 512}, 6),
    "opnet_lstm_mlp": ({"object_to_track_pred_dim": 15, "object_to_track_hidden_dim": 256, "videos_hidden_dim": 512}, 6),
    "baseline_lstm": ({"videos_hidden_dim": 512}, 5),
    "non_linear_lstm": ({"boxes_features_dim": 256, "videos_hidden_dim": 512}, 5),
    "transformer_lstm": ({"boxes_features_dim": 256, "num_attention_heads": 2, "num_attention_layers": 2, "num_lstm_layers": 2, "lstm_hidden_dim": 512}, 5),
}
B, T = 32, 300
for name, (cfg, F) in CONFIGS.items():
    torch.manual_seed(0)
    model = ModelsFactory.get_model(name, cfg).to(dev).train()
    b, l, m = make_batch(B, T, F, seed=1)
    boxes, labels = torch.from_numpy(b).to(dev), torch.from_numpy(l).to(dev)
    def step():
        for p in model.parameters(): p.grad = None
        out = model(boxes)
        y = out[0] if isinstance(out, tuple) else out
        loss = ops.training_loss(y, labels, None, False)
        loss[0].backward()
        return loss
    step(); step(); torch.cuda.synchronize()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): loss = step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name:18s} [B={B},T={T}] fwd+loss+bwd: {ms:7.2f} ms/step = {B / ms * 1e3:8.1f} videos/s, kernels/step {(_lib.launch_count() - n0) // 5}, "
          f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB", flush=True)
    del model
    torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()
