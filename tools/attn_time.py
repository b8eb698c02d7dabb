"""Fused (flash-style tcgen05) attention against the materialised form at config 3's shape: S = 9600, 2 heads of 128."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectpermanence_b200 import ops
dev = torch.device("cuda:0")
S, nhead, D = int(os.environ.get("ATTN_S", "9600")), 2, 256
g = torch.Generator().manual_seed(0)
qkv = (torch.rand(S, 3 * D, generator=g) * 2 - 1).to(dev).requires_grad_(True)
dctx = (torch.rand(S, D, generator=g) * 2 - 1).to(dev)
for mode in ("fused", "materialized"):
    os.environ["OPN_ATTENTION"] = mode
    for p_drop in (0.0, 0.1):
        def run():
            qkv.grad = None
            out = ops.self_attention(qkv, nhead, p_drop, 11, 0)
            return out
        for _ in range(3):
            out = run(); out.backward(dctx)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        n = 5
        tf = tb = 0.0
        for _ in range(n):
            ev[0].record(); out = run(); ev[1].record(); out.backward(dctx); ev[2].record(); torch.cuda.synchronize()
            tf += ev[0].elapsed_time(ev[1]); tb += ev[1].elapsed_time(ev[2])
        flops = 2 * 2 * nhead * S * S * 128            # QK^T + PV, forward
        print(f"S={S} {mode:13s} p_drop={p_drop}: fwd {tf / n:7.3f} ms ({flops / (tf / n * 1e-3) / 1e12:6.1f} TF/s useful)  "
              f"bwd {tb / n:7.3f} ms ({2.5 * flops / (tb / n * 1e-3) / 1e12:6.1f} TF/s useful)", flush=True)
