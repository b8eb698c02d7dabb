"""Per-role cycle breakdown of the fused attention forward (phases build; OPN_B200_LIB=.../libopnet_b200_phases.so)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectpermanence_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda:0")
S, nhead, D = 9600, 2, 256
qkv = (torch.rand(S, 3 * D, device=dev) * 2 - 1)
out = torch.empty(S, D, device=dev)
ws = torch.zeros(lib.opn_attention_workspace_bytes(S, D, nhead), dtype=torch.uint8, device=dev)
s = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    _lib.check(lib.opn_attention_fwd(S, D, nhead, qkv.data_ptr(), out.data_ptr(), ws.data_ptr(), ws.numel(), 0.0, 0, 0, s))
torch.cuda.synchronize()
w = ws[:4096].view(torch.int64).cpu()
n = S // 64
for role, name, labels in ((0, "softmax thread", ["loop", "wait s_full", "tmem ld + release", "max/exp/sum", "wait P planes free", "cvt + tmem st + arrive"]),
                           (1, "MMA thread", ["wait K/V tile", "wait S buffer", "issue S(j+1)", "wait P(j)", "issue PV(j)"])):
    ph = w[32 + 8 * role: 32 + 8 * role + len(labels)].tolist()
    print(f"{name:15s}: total {sum(ph) / n:7.0f} clk/tile | " + "  ".join(f"{l} {v / n:6.0f}" for l, v in zip(labels, ph)))

# ---- backward, key-tile kernel (roles 2, 3) ----
dctx = torch.rand(S, D, device=dev) * 2 - 1
dqkv = torch.empty(S, 3 * D, device=dev)
for _ in range(2):
    _lib.check(lib.opn_attention_bwd(S, D, nhead, out.data_ptr(), dctx.data_ptr(), dqkv.data_ptr(), ws.data_ptr(), ws.numel(), 0.0, 0, 0, s))
torch.cuda.synchronize()
w = ws[:4096].view(torch.int64).cpu()
for role, name, labels in ((2, "bwd_kv softmax thread", ["row stats + barriers", "wait S/dP", "tmem ld + exp + dS", "wait planes free", "cvt + tmem st + arrive"]),
                           (3, "bwd_kv MMA thread", ["wait dO tile", "issue dP^T(i)", "wait + issue S^T(i+1)", "wait planes", "issue dV, dK"])):
    ph = w[32 + 8 * role: 32 + 8 * role + len(labels)].tolist()
    print(f"{name:22s}: total {sum(ph) / n:7.0f} clk/tile | " + "  ".join(f"{l} {v / n:6.0f}" for l, v in zip(labels, ph)))
