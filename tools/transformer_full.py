"""transformer_lstm at BASELINE config 3 shape [B=32,T=300,N=15,F=5]: fwd+loss+bwd timing and peak memory."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectpermanence_b200 import ops, _lib
from objectpermanence_b200.models_factory import ModelsFactory
from objectpermanence_b200.synthetic import make_batch
dev = torch.device("cuda:0")
cfg = {"boxes_features_dim": 256, "num_attention_heads": 2, "num_attention_layers": 2, "num_lstm_layers": 2, "lstm_hidden_dim": 512}
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.manual_seed(0)
model = ModelsFactory.get_model("transformer_lstm", cfg).to(dev).train()
b, l, m = make_batch(B, 300, 5, seed=1)
boxes, labels = torch.from_numpy(b).to(dev), torch.from_numpy(l).to(dev)
def step():
    for p in model.parameters(): p.grad = None
    y = model(boxes)
    loss = ops.training_loss(y, labels, None, False)
    loss[0].backward()
    return loss
for mode in (True, False):      # train mode applies the encoder's four dropout sites (p = 0.1); eval mode is the parity mode
    model.train(mode)
    step(); torch.cuda.synchronize()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): loss = step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"transformer_lstm [B={B},T=300] {'train (dropout 0.1)' if mode else 'eval (no dropout)'} fwd+loss+bwd: {ms:.2f} ms/step = "
          f"{B / ms * 1e3:.1f} videos/s, loss {loss[0].item():.5f}, kernels/step {(_lib.launch_count() - n0) // 3}, "
          f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB", flush=True)
