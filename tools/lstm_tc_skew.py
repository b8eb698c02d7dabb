"""Wall-clock stamps of one forward step of every CTA of the tcgen05 recurrence (phases build): who is late, and how long the
hand-over takes after the LAST publish."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectpermanence_b200 import _lib
os.environ["OPN_LSTM_TC"] = "1"
lib = _lib.load(); dev = torch.device("cuda:0"); T = 300
s = torch.cuda.current_stream().cuda_stream
H, B = 512, 256
xp = torch.randn(B, T, 4 * H, device=dev) * 0.5
whh = (torch.rand(4 * H, H, device=dev) * 2 - 1) / (H ** 0.5)
hs = torch.empty(B, T, H, device=dev); gates = torch.empty(B, T, 4 * H, device=dev); cells = torch.empty(B, T, H, device=dev)
ws = torch.zeros(lib.opn_lstm_workspace_bytes(B, T, H), dtype=torch.uint8, device=dev)
for _ in range(2):
    _lib.check(lib.opn_lstm_fwd(B, T, H, xp.data_ptr(), whh.data_ptr(), hs.data_ptr(), gates.data_ptr(), cells.data_ptr(), ws.data_ptr(), ws.numel(), s))
torch.cuda.synchronize()
st = ws[4096:8192].view(torch.int64).cpu()[64:64 + 4 * 64].reshape(64, 4).double()
t0 = st[:, 0].min()
st = (st - t0)
print("per CTA (ns from the first acc_full of step 99): acc_full | published | counters ok | first tile")
for g in range(2):
    blk = st[32 * g:32 * g + 32]
    print(f"group {g}: acc_full {blk[:,0].min():6.0f}..{blk[:,0].max():6.0f}  published {blk[:,1].min():6.0f}..{blk[:,1].max():6.0f}  "
          f"counters {blk[:,2].min():6.0f}..{blk[:,2].max():6.0f}  first tile {blk[:,3].min():6.0f}..{blk[:,3].max():6.0f}")
print(st[:8])
