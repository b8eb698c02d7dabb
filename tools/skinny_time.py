"""Time the skinny contractions of the OPNet step (CUDA events, 20 repetitions each, 256 MB L2 flush in between)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectpermanence_b200 import ops
dev = torch.device("cuda:0")
flush = torch.empty(64 * 1024 * 1024, device=dev)
CASES = [("y = h2 Wo^T", False, True, 9600, 4, 512), ("dfb = dgates2 W_ih2", False, False, 9600, 6, 2048),
         ("dh2 = dy Wo", False, False, 9600, 512, 4), ("xproj2 = fb W_ih2^T", False, True, 9600, 2048, 6),
         ("dWo = dy^T h2", True, False, 4, 512, 9600), ("dW_ih2 = dgates2^T fb", True, False, 2048, 6, 9600),
         ("dWp^T = hs1^T dl", True, False, 256, 15, 9600)]
for name, ta, tb, M, N, K in CASES:
    A = torch.randn((K, M) if ta else (M, K), device=dev); B = torch.randn((N, K) if tb else (K, N), device=dev)
    C = torch.empty(M, N, device=dev)
    f = lambda: ops.sgemm(A, B, C, trans_a=ta, trans_b=tb, M=M, N=N, K=K, lda=A.shape[1], ldb=B.shape[1], ldc=N)
    f(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(20):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    mb = (A.numel() + B.numel() + C.numel()) * 4 / 1e6
    print(f"{name:24s} M={M:5d} N={N:5d} K={K:5d}: {tot / 20 * 1e3:7.1f} us  ({mb / (tot / 20 * 1e-3) / 1e3:6.0f} GB/s of {mb:.1f} MB)", flush=True)
